// total_rnn1 / total_rnn2 (lstmmodel.py:49,51: Bidirectional(LSTM(128)) on [read_rnn2 | signal] and Bidirectional(LSTM(64))
// on its BatchNormalised output) as ONE persistent tcgen05 kernel per layer: input projection AND recurrence fused, so the
// fp32 gate pre-activations (zin: 2 x 45 KB resp. 2 x 22.5 KB per window, written by a GEMM and read back by a recurrence
// kernel -- 70 % of the whole path's HBM traffic) never exist.
//
// Work split.  A CTA PAIR (tcgen05 cta_group::2) owns 256 windows (128 per CTA = the 128 TMEM lanes) x 256 gate columns
// (= 64 LSTM units x gates i,f,c,o, interleaved col = unit*4 + gate) of one direction.  One MMA covers M = 256 x N = 256
// and takes half of the B operand from each CTA, so [Wk ; Wr]^T for those 256 columns (K = KIN + UT = 320 for both layers,
// fp16 hi + lo) is RESIDENT: 160 KB per SM, loaded once per kernel.
//   * total_rnn2 (UT = 64):  one pair has all 256 gate columns.                       Cluster = 2 CTAs.
//   * total_rnn1 (UT = 128): two pairs split the 512 gate columns (units 0-63 / 64-127) of the SAME 256 windows and
//     exchange their halves of h_t every step.                                       Cluster = 4 CTAs.
//
// Per SM: 160 KB weights + 4 x 16 KB TMA ring of x_t tiles ([128 rows][64 K], lo then hi part of every K-chunk) + bias.
// There is no room for an h tile in shared memory, so h_{t-1} -- the A operand of the recurrent MMAs -- lives in TENSOR
// MEMORY: every epilogue thread owns one TMEM lane (= one window) and tcgen05.st's its new h (fp16 hi / lo pairs) next to
// the 256-column accumulator; the recurrent MMAs use the A-from-TMEM (.ts) form (layout pinned by tools/ts_probe).
//
// Pipeline.  The accumulator is two unit blocks of 128 columns (N = 128 MMAs); block 1 alternates between two TMEM regions, so
// only block 0 is ever waited for (`drained`: 32 registers per thread, right after the commit).  The epilogue (16 warps, one
// TMEM lane = one window per thread, 2 x 8 units) delivers h_t in QUARTERS (own block 0, sibling's block 0, own block 1,
// sibling's block 1: four mbarriers) and the MMA warp interleaves statically
//     proj kc0, kc1 | rec(q0) | proj kc2 | rec(q0') | [proj kc3] | rec(q1) | rec(q1') | commit
// so that only the 12 MMAs of the last quarter stand between the end of the epilogue and the next accumulator.  x_t is read
// exactly once (TMA ring, next step prefetched into L2 by the producer warp).
//
// total_rnn1's exchange: the thread that owns window w in pair p and the thread that owns w in pair 1-p (the "sibling" CTA)
// swap their 16 units of h_t in four halves of 4 units.  A warp stages a half (fp16 hi / lo, 2 x 256 B) in local shared memory
// and one lane hands it to the bulk-copy engine (cp.async.bulk.shared::cluster into a 16 KB landing zone of the sibling, bytes
// counted on the sibling WARP's mbarrier): no fences, no global round trip, nobody stalls on the ~20 B/clock DSMEM path, and the
// traffic is spread over the gate arithmetic.  The receiver copies its 32 B into TMEM and releases the zone with a remote
// arrive (`xfree`), which gates the sender's next phase.  Landing zone + staging cost two ring stages (2 x 16 KB instead of 4).
// History of the exchange (L2 round trip, st.async bursts, whole-block copies, a dedicated sender warp): DESIGN.md section 4.
#include <algorithm>

#include <cuda_fp8.h>

#include "nrv_cell.cuh"
#include "nrv_common.cuh"
#include "nrv_tc.cuh"

namespace nrv {

using namespace tc;

bool make_tmap_f16_k64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);   // nrv_gemm.cu
bool make_tmap_u8_k128(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);    // nrv_gemm.cu

constexpr int FP_EPI_WARPS = 16;                      // 4 per TMEM lane quarter; 16 units (64 gate columns) per thread
constexpr int FP_THREADS = 64 + 32 * FP_EPI_WARPS;    // warp 0: MMA issue (pair leader); warp 1: TMA producer; warps 2..17: epilogue
constexpr int FP_TILE = 128 * 64 * 2;                 // 16 KB: [128 rows][64 halves], K-major, 128-byte swizzle
constexpr int FP_XCH_BYTES = 16 * 1024;             // total_rnn1: landing zone of the sibling pair's h (one 8-unit block per thread, hi + lo)
constexpr uint32_t FP_H_COL = 384;                    // TMEM: accumulator block 0 [0, 128); block 1 [128, 256) on even steps, [256, 384) on
                                                      // odd steps; h_hi [384, 384 + UT/2); h_lo [.., 384 + UT)

template <int KIN, int UT>
struct FpCfg {
    static constexpr int KC = KIN / 64, RC = UT / 64, NP = UT / 64;
    static constexpr int W_BYTES = (KC + RC) * 2 * FP_TILE;
    static constexpr int STAGES = NP == 1 ? 4 : 2;                   // x ring depth (16 KB tiles)
    static constexpr int XCH = NP == 1 ? 0 : 2 * FP_XCH_BYTES;       // total_rnn1: landing zone + outgoing staging of the h exchange
    static constexpr size_t SMEM = (size_t)W_BYTES + STAGES * FP_TILE + XCH + 1024 /*bias*/ + 512 /*barriers*/ + 1024 /*alignment*/;
};

__device__ __forceinline__ void umma_commit_mask(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
#ifdef NRV_TRACE
// timeline of cluster 0 / direction 0 (debug builds: NRV_EXTRA_NVCC=-DNRV_TRACE): SM clock at key events of the MMA warp (role 0)
// and of the first epilogue warp (role 1), printed by the 4th launch of each instantiation
__device__ long long g_tr[2][2][128][8];
__device__ unsigned long long g_tr_cl[2][2][64][3];     // [inst][dir][cluster]: globaltimer at start / end, SM id
__device__ int g_tr_launch[2];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TR(role, step, ev) do { if (trace_on && (threadIdx.x & 31) == 0 && (step) < 128) g_tr[UT == 128][role][step][ev] = clock64() - tr_t0; } while (0)
#else
#define TR(role, step, ev) do { } while (0)
#endif
#ifdef NRV_HANG_DEBUG
// bounded spin: report which barrier / role / step is stuck and trap (debug builds only: NRV_EXTRA_NVCC=-DNRV_HANG_DEBUG)
template <bool CL>
__device__ __forceinline__ void fp_wait_dbg(uint64_t* bar, uint32_t parity, int tag, uint32_t g) {
    const uint32_t a = smem_u32(bar);
    for (long long it = 0;; ++it) {
        uint32_t ok;
        if (CL)
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                         : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        else
            asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}\n"
                         : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return;
        if (it > (1ll << 24)) {
            printf("HANG tag %d blk (%d,%d) rank %u warp %d lane %d g %u parity %u\n", tag, blockIdx.x, blockIdx.y, cluster_ctarank(),
                   threadIdx.x >> 5, threadIdx.x & 31, g, parity);
            __trap();
        }
    }
}
#define FP_WAIT(bar, parity, tag, g) fp_wait_dbg<false>(bar, parity, tag, g)
#else
#define FP_WAIT(bar, parity, tag, g) mbar_wait(bar, parity)
#endif
// 2-D tile -> L2 only (no shared-memory destination)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1)
                 : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy encodings as CUTLASS uses them: evict_first / evict_last, fraction 1)
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull, L2_EVICT_LAST = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_prefetch_2d_hint(const CUtensorMap* m, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile.L2::cache_hint [%0, {%1, %2}], %3;\n" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, uint32_t bar_cta, int c0, int c1,
                                                      uint64_t pol) {
    asm volatile(
        "{\n\t.reg .b32 remBar;\n\t"
        "mapa.shared::cluster.u32 remBar, %2, %5;\n\t"
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [remBar], "
        "%6;\n\t}\n" ::"r"(tc::smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(bar_cta), "l"(pol)
        : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(saddr), "r"(cta));
    return r;
}
// bulk copy (DMA) from this CTA's shared memory into another CTA's; the bytes are counted on that CTA's mbarrier (complete_tx)
__device__ __forceinline__ void bulk_copy_to_cluster(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(remote_dst),
                 "r"(local_src), "r"(bytes), "r"(remote_bar)
                 : "memory");
}

// Units [J0, J1) of one 32-column block (8 units x gates i,f,c,o) of the cell; the bias is read from shared memory unit by
// unit (keeps the 64 drained accumulator registers + 16 cell states close to the 96-register budget of a 576-thread CTA:
// warps are allocated in fours, so 20 x 32 x 96 registers).
// `zs` = scale of the accumulator (1 for the fp16 x 3 kernels; 2^-S when the operands carry power-of-two scales, see F8), `zs02` = 0.2 zs
template <int J0, int J1>
__device__ __forceinline__ void cell_units(const uint32_t (&v)[32], uint32_t sbias, float* c8, float* hv, float zs, float zs02) {
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        const float4 b = ld_shared_f4(sbias + j * 16);       // {0.2 b_i + 0.5, 0.2 b_f + 0.5, b_c, 0.2 b_o + 0.5}
        const float ig = __saturatef(fmaf(zs02, __uint_as_float(v[4 * j + 0]), b.x));
        const float fg = __saturatef(fmaf(zs02, __uint_as_float(v[4 * j + 1]), b.y));
#if defined(NRV_TANH_NR) && NRV_TANH_NR >= 1
        const float gg = tanh_fast_nr(fmaf(zs, __uint_as_float(v[4 * j + 2]), b.z));
#else
        const float gg = tanh_fast(fmaf(zs, __uint_as_float(v[4 * j + 2]), b.z));
#endif
        const float og = __saturatef(fmaf(zs02, __uint_as_float(v[4 * j + 3]), b.w));
        const float cn = fmaf(fg, c8[j], ig * gg);
        c8[j] = cn;
#if defined(NRV_TANH_NR) && NRV_TANH_NR >= 2
        hv[j] = og * tanh_fast_nr(cn);
#else
        hv[j] = og * tanh_fast(cn);
#endif
    }
}
// 4 units of h = hi + lo as fp16 pairs
__device__ __forceinline__ void pack_h4(const float* hv, uint2& phi, uint2& plo) {
    uint32_t ph[2], pl[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const __half2 hi = __floats2half2_rn(hv[2 * p], hv[2 * p + 1]);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(hv[2 * p] - hf.x, hv[2 * p + 1] - hf.y);
        ph[p] = half2_bits(hi); pl[p] = half2_bits(lo);
    }
    phi = make_uint2(ph[0], ph[1]);
    plo = make_uint2(pl[0], pl[1]);
}

// ---- 8-bit copies of activations for a product that runs its correction passes in e4m3 (F8 below) ----
// An activation x is handed over as its ordinary fp16 hi part and, per group of 4 units, 8 bytes
// {e4m3(x_lo 2^12) x 4, e4m3(x_hi) x 4}: as many bytes per row as the fp16 lo part they replace.
constexpr float F8_XLO = 4096.f;                 // 2^12: scale of the e4m3 copy of the lo part (|x_lo| <= 2^-12 |x|)
__device__ __forceinline__ uint32_t e4m3x2_f32(float a, float b) {
    return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ uint32_t e4m3x2_h2(__half2 v) {
    return (uint32_t)__nv_cvt_halfraw2_to_fp8x2(static_cast<__half2_raw>(v), __NV_SATFINITE, __NV_E4M3);
}
// 4 units of h: fp16 hi (phi), fp16 lo (plo) and the 8-bit copies (pq).  WANT_LO = false skips the fp16 lo part (nobody reads it
// when both the kernel's own recurrence and its consumer take the 8-bit copies)
template <bool WANT_LO>
__device__ __forceinline__ void pack_h4_f8(const float* hv, uint2& phi, uint2& plo, uint2& pq) {
    uint32_t ph[2], pl[2] = {0, 0}, l8[2], h8[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const __half2 hi = __floats2half2_rn(hv[2 * p], hv[2 * p + 1]);
        const float2 hf = __half22float2(hi);
        const float l0 = hv[2 * p] - hf.x, l1 = hv[2 * p + 1] - hf.y;
        ph[p] = half2_bits(hi);
        if (WANT_LO) pl[p] = half2_bits(__floats2half2_rn(l0, l1));
        l8[p] = e4m3x2_f32(l0 * F8_XLO, l1 * F8_XLO);
        h8[p] = e4m3x2_h2(hi);
    }
    phi = make_uint2(ph[0], ph[1]);
    plo = make_uint2(pl[0], pl[1]);
    pq = make_uint2(l8[0] | (l8[1] << 16), h8[0] | (h8[1] << 16));
}

// F8: the two correction passes of a product run as ONE e4m3 product (tcgen05 kind::f8f6f4: K = 32 bytes per instruction, twice the
// fp16 rate) on 8-bit copies of the operands.  The producing epilogue writes {e4m3(x_lo 2^12) x 4, e4m3(x_hi) x 4} per group of 4
// units in place of the fp16 lo part; pack_model lays the weights out the same way ({e4m3(W_hi 2^(S-12)) x 4, e4m3(W_lo 2^S) x 4}), so
// x_lo . W_hi + x_hi . W_lo is one K-contiguous e4m3 product over a 128-byte tile per 64 units: 2 instructions (1 fp16 + 1 e4m3) per
// K-step of 16 units instead of 3.  The fp16 hi parts of the ACTIVATIONS are the ordinary unscaled ones; the WEIGHTS carry power-of-two
// scales chosen per layer (S = 7 + floor(log2(448 / max|W|))) so that every pass accumulates 2^S z into the same fp32 accumulator:
//     fp16(x) . fp16(W 2^S)  +  e4m3(x_lo 2^12) . e4m3(W_hi 2^(S-12))  +  e4m3(x_hi) . e4m3(W_lo 2^S)
// and the epilogue folds 2^-S into the gate constants (free).
//   PF8: the PROJECTION runs this way (`wk_lo`, `tm_x_lo` hold the 8-bit copies; `wk_hi` the scaled fp16 weights);
//   RF8: the RECURRENCE does (`wr_lo` and the h_lo columns of TMEM hold the 8-bit copies, and so do the halves two pairs exchange);
//        with PF8 = false the projection stays fp16 x 3 on scaled weights (hi AND lo of W 2^S), so the layer's INPUT needs no copies;
//   OUT8: the CONSUMER is a PF8 layer: out_lo receives the 8-bit copies (with RF8 they are the registers the recurrence uses anyway).
// total_rnn2 = <PF8, RF8>, total_rnn1 = <RF8, OUT8>.  8-bit copies occupy the same bytes per row / tile / column as the fp16 lo parts
// they replace.  Precision: tests/precision_study.py and DESIGN.md section 4.
template <int KIN, int UT, bool PF8, bool RF8, bool OUT8>
__global__ void __launch_bounds__(FP_THREADS, 1)
lstm_fused_pair_kernel(const __half* __restrict__ wk_hi, const __half* __restrict__ wk_lo, const __half* __restrict__ wr_hi,
                       const __half* __restrict__ wr_lo, const float* __restrict__ bias,
                       const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                       __half* __restrict__ out_hi, __half* __restrict__ out_lo, int out_ld, int64_t nwp, int T, float acc_scale,
                       // total_rnn1 only: tile_base != nullptr and tile_base[tile] >= 0 -> the last K-chunk (CNN features) of that window
                       // tile is rows tile_base[tile] + t .. + 127 of the per-base feature table (tm_sf_hi / tm_sf_lo) instead of x
                       const __grid_constant__ CUtensorMap tm_sf_hi, const __grid_constant__ CUtensorMap tm_sf_lo,
                       const int32_t* __restrict__ tile_base,
                       // != 0: L2 eviction hints on the x tiles.  The forward and the backward cluster of a window tile read x_t at steps t
                       // and T-1-t: the first of the two readers asks L2 to keep the tile (evict_last), the second marks it dead
                       // (evict_first) -- total_rnn2's working set per round (213 MB) does not fit L2 without that
                       int l2hint) {
    using Cfg = FpCfg<KIN, UT>;
    constexpr int KC = Cfg::KC, NP = Cfg::NP, CS = 2 * NP, FP_STAGES = Cfg::STAGES;
    constexpr int NT = 4 * UT;                              // gate columns per direction
    constexpr uint32_t H_HI = FP_H_COL, H_LO = FP_H_COL + UT / 2;      // RF8: the H_LO columns hold the 8-bit copies of h
    constexpr int XLO_W = PF8 ? 128 : 64;                    // elements per 128-byte row of a lo tile (bytes / halves)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                    // [Wk chunk 0..KC-1 | Wr chunk 0..RC-1][hi | lo][128 rows][64]
    uint8_t* s_ring = smem + Cfg::W_BYTES;                  // [stage][128 rows][64]
    uint8_t* s_xch = s_ring + FP_STAGES * FP_TILE;          // NP == 2: [cg][hi | lo][128 rows] x 16 B, written by the sibling CTA (st.async)
    float* s_bias = reinterpret_cast<float*>(s_xch + Cfg::XCH);               // [256]: this pair's gate columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
    uint64_t* full = bars;                                  // [STAGES] pair leader's copy: 1 arrive + 32 KB tx (both CTAs' tiles)
    uint64_t* empty = bars + FP_STAGES;                     // [STAGES] both CTAs of the pair: multicast commit
    uint64_t* acc_ready = bars + 2 * FP_STAGES;             // both CTAs of the pair: multicast commit
    uint64_t* drained = acc_ready + 1;                      // pair leader's copy: 32 arrivals (16 epilogue warps x 2 CTAs)
    uint64_t* hq = acc_ready + 2;                           // [4] pair leader's copy, 32 arrivals each: quarter (b, who) of h_t is in
                                                            // TMEM in both CTAs; b = unit block 0/1, who = 0 own pair's units, 1 sibling's
    uint64_t* xfull = acc_ready + 6;                        // NP == 2: [16] per epilogue warp: 1 arrive.expect_tx + 1 KB from the sibling warp
    uint64_t* xfree = xfull + FP_EPI_WARPS;                 // NP == 2: [16] per epilogue warp: the sibling warp has read our block
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfree + FP_EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef NRV_TRACE
    const bool trace_on = blockIdx.x == 0 && blockIdx.y == 0;
    const long long tr_t0 = clock64();
    if (threadIdx.x == 0 && cluster_ctarank() == 0 && blockIdx.x / (2 * Cfg::NP) < 64) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_tr_cl[UT == 128][blockIdx.y][blockIdx.x / (2 * Cfg::NP)][0] = gtimer();
        g_tr_cl[UT == 128][blockIdx.y][blockIdx.x / (2 * Cfg::NP)][2] = smid;
    }
#endif
    const uint32_t rank = cluster_ctarank();
    const uint32_t p = rank >> 1, r = rank & 1, leader = rank & ~1u;
    const int dir = blockIdx.y;
    const int64_t ntw = nwp >> 7;
    const int64_t n_pairs = (ntw + 1) >> 1;
    const int64_t cl0 = blockIdx.x / CS, cl_stride = gridDim.x / CS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < FP_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_ready, 1);
        mbar_init(drained, 2 * FP_EPI_WARPS);
        for (int i = 0; i < 4; ++i) mbar_init(&hq[i], 2 * FP_EPI_WARPS);
        for (int i = 0; i < FP_EPI_WARPS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xfree[i], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo);
        if (tile_base) { tma_prefetch_desc(&tm_sf_hi); tma_prefetch_desc(&tm_sf_lo); }
    }
    if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
    if (threadIdx.x < 256) {       // hard_sigmoid(z + b) = sat(0.2 z + (0.2 b + 0.5)): the gates i, f, o keep the folded constant
        const float b = __ldg(bias + dir * NT + p * 256 + threadIdx.x);
        s_bias[threadIdx.x] = (threadIdx.x & 3) == 2 ? b : fmaf(0.2f, b, 0.5f);
    }
    {   // resident weights: this CTA's 128 gate columns, all of K, hi and lo.  Shared-memory row b*64 + j = gate column
        // b*128 + r*64 + j of the pair: an N = 128 MMA on unit block b takes rows [b*64, +64) from each CTA of the pair
        const size_t grow0 = (size_t)dir * NT + p * 256 + r * 64;
        auto gcol = [](int row) { return (row >> 6) * 128 + (row & 63); };
        constexpr int CK = KIN / 8, CR = UT / 8;            // 16-byte chunks per row
        // (PF8 / RF8: the lo parts are the interleaved 8-bit copies -- as many bytes per row as the fp16 lo part, same tiles)
        for (int i = threadIdx.x; i < 2 * 128 * CK; i += FP_THREADS) {
            const int c = i % CK, row = (i / CK) & 127, part = i / (CK * 128);
            const __half* src = (part ? wk_lo : wk_hi) + (grow0 + gcol(row)) * KIN;
            *reinterpret_cast<uint4*>(s_w + (size_t)((c >> 3) * 2 + part) * FP_TILE + sw128_offset(row, c & 7)) =
                __ldg(reinterpret_cast<const uint4*>(src) + c);
        }
        for (int i = threadIdx.x; i < 2 * 128 * CR; i += FP_THREADS) {
            const int c = i % CR, row = (i / CR) & 127, part = i / (CR * 128);
            const __half* src = (part ? wr_lo : wr_hi) + (grow0 + gcol(row)) * UT;
            *reinterpret_cast<uint4*>(s_w + (size_t)((KC + (c >> 3)) * 2 + part) * FP_TILE + sw128_offset(row, c & 7)) =
                __ldg(reinterpret_cast<const uint4*>(src) + c);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // barriers initialised, weights in place and TMEM allocated in every CTA of the cluster
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        // ===================== TMA producer (every CTA): x_t tiles of this CTA's 128 windows, in consumption order =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t tp = cl0; tp < n_pairs; tp += cl_stride) {
                const int64_t wtile = min(tp * 2 + (int64_t)r, ntw - 1);
                const int tb = tile_base ? __ldg(tile_base + wtile) : -1;      // >= 0: CNN features of this tile from the per-base table
                for (int s = 0; s < T; ++s) {
                    const int t = dir ? (T - 1 - s) : s;
                    const int grow = (int)(t * nwp + wtile * 128);
                    if (s + 1 < T) {                                       // next step's tiles -> L2 (the layer input comes from HBM: ring
                        const int gnext = grow + (dir ? -1 : 1) * (int)nwp;   // loads then see L2 latency, which 3-4 stages cover)
                        const int kx = tb >= 0 ? KC - 1 : KC;
                        if (l2hint) {
                            const uint64_t pn = (2 * (s + 1) + 1 < T) ? L2_EVICT_LAST : L2_EVICT_FIRST;
                            for (int i = 0; i < kx; ++i) { tma_prefetch_2d_hint(&tm_x_lo, i * XLO_W, gnext, pn); tma_prefetch_2d_hint(&tm_x_hi, i * 64, gnext, pn); }
                        } else {
                            for (int i = 0; i < kx; ++i) { tma_prefetch_2d(&tm_x_lo, i * XLO_W, gnext); tma_prefetch_2d(&tm_x_hi, i * 64, gnext); }
                        }
                        if (tb >= 0) { tma_prefetch_2d(&tm_sf_lo, 0, tb + t + (dir ? -1 : 1)); tma_prefetch_2d(&tm_sf_hi, 0, tb + t + (dir ? -1 : 1)); }
                    }
                    for (int i = 0; i < 2 * KC; ++i) {                     // (K-chunk, part): lo tile (F8: the 8-bit copies) first, then hi
                        FP_WAIT(&empty[stage], phase ^ 1, 1, (uint32_t)s);
                        if (r == 0) mbar_arrive_expect_tx(&full[stage], 2 * FP_TILE);
                        if (tb >= 0 && (i >> 1) == KC - 1)
                            tma_load_2d_pair(s_ring + (size_t)stage * FP_TILE, (i & 1) ? &tm_sf_hi : &tm_sf_lo, &full[stage], leader, 0, tb + t);
                        else if (l2hint)
                            tma_load_2d_pair_hint(s_ring + (size_t)stage * FP_TILE, (i & 1) ? &tm_x_hi : &tm_x_lo, &full[stage], leader,
                                                  (i >> 1) * ((i & 1) ? 64 : XLO_W), grow, (2 * s + 1 < T) ? L2_EVICT_LAST : L2_EVICT_FIRST);
                        else
                            tma_load_2d_pair(s_ring + (size_t)stage * FP_TILE, (i & 1) ? &tm_x_hi : &tm_x_lo, &full[stage], leader,
                                             (i >> 1) * ((i & 1) ? 64 : XLO_W), grow);
                        if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ===================== MMA issue (leader CTA of every pair) =====================
        if (r == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(256, 128);     // one MMA = 256 windows x one unit block (128 gate columns)
            constexpr uint32_t BB = 64 * 128;                           // byte offset of block 1's weight rows inside a tile
            const uint16_t mask = (uint16_t)(3u << leader);
            const uint32_t w_base = smem_u32(s_w), ring_base = smem_u32(s_ring);
            int stage = 0; uint32_t phase = 0;
            uint32_t g = 0;                                               // running step count over all tiles
            for (int64_t tp = cl0; tp < n_pairs; tp += cl_stride)
                for (int s = 0; s < T; ++s, ++g) {
                    if (g > 0) {                                          // the previous step's accumulator is in registers
                        FP_WAIT(drained, (g - 1) & 1, 2, g);
                        tc_fence_after();
                    }
                    const uint32_t d0 = tmem_base, d1 = tmem_base + 128 + (g & 1) * 128;      // block 1 alternates: never waits for a drain
                    // projection K-chunk kc: x_lo . W_hi, then x_hi . W_lo and x_hi . W_hi (two ring tiles), both unit blocks.
                    // F8: the lo tile is the 8-bit copies of the chunk's 64 units: 4 e4m3 MMAs (K = 32 bytes each) against the 8-bit
                    // weight tile are both correction passes; then the fp16 main pass on the hi tile
                    auto proj = [&](int kc) {
                        const uint32_t wb_hi = w_base + (uint32_t)((kc * 2 + 0) * FP_TILE), wb_lo = w_base + (uint32_t)((kc * 2 + 1) * FP_TILE);
                        FP_WAIT(&full[stage], phase, 3, g);                   // x_lo(kc)
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t xa = ring_base + (uint32_t)(stage * FP_TILE);
#pragma unroll
#ifdef NRV_FP_SKIP_XLO     // precision experiment only: drop the x_lo . W_hi pass of the projection (NP mask: 1 = total_rnn2, 2 = total_rnn1)
                            if (!((NRV_FP_SKIP_XLO) & NP))
#endif
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t a_lo = umma_desc_k_sw128(xa + k * 32);
                                if constexpr (PF8) {
                                    umma_f8_ss_pair(d0, a_lo, umma_desc_k_sw128(wb_lo + k * 32), idesc, (kc | k) != 0);
                                    umma_f8_ss_pair(d1, a_lo, umma_desc_k_sw128(wb_lo + BB + k * 32), idesc, (kc | k) != 0);
                                } else {
                                    umma_f16_ss_pair(d0, a_lo, umma_desc_k_sw128(wb_hi + k * 32), idesc, (kc | k) != 0);
                                    umma_f16_ss_pair(d1, a_lo, umma_desc_k_sw128(wb_hi + BB + k * 32), idesc, (kc | k) != 0);
                                }
                            }
                            umma_commit_mask(&empty[stage], mask);
                        }
                        __syncwarp();
                        if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
                        FP_WAIT(&full[stage], phase, 4, g);                   // x_hi(kc)
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t xa = ring_base + (uint32_t)(stage * FP_TILE);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t a_hi = umma_desc_k_sw128(xa + k * 32);
#ifdef NRV_FP_SKIP_XLO
                                const uint32_t acc0 = ((NRV_FP_SKIP_XLO) & NP) ? (uint32_t)((kc | k) != 0) : 1u;
#else
                                const uint32_t acc0 = 1u;
#endif
                                if constexpr (!PF8) umma_f16_ss_pair(d0, a_hi, umma_desc_k_sw128(wb_lo + k * 32), idesc, acc0);
                                umma_f16_ss_pair(d0, a_hi, umma_desc_k_sw128(wb_hi + k * 32), idesc, 1);
                                if constexpr (!PF8) umma_f16_ss_pair(d1, a_hi, umma_desc_k_sw128(wb_lo + BB + k * 32), idesc, acc0);
                                umma_f16_ss_pair(d1, a_hi, umma_desc_k_sw128(wb_hi + BB + k * 32), idesc, 1);
                            }
                            umma_commit_mask(&empty[stage], mask);
                        }
                        __syncwarp();
                        if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
                    };
                    // recurrent quarter (b, who): the 32 units b*32.. of pair `who ? 1-p : p` = K-steps k0, k0 + 1 of h_{t-1} (TMEM).
                    // F8: the 8 H_LO columns of a K-step are the 32 bytes of 8-bit copies of its 16 units = one e4m3 MMA (layout pinned by
                    // tools/ts_probe/ts_probe8.cu: column c = bytes 4c .. 4c + 3)
                    auto rec = [&](int b, int who) {
                        if (g > 0) {
                            FP_WAIT(&hq[b * 2 + who], (g - 1) & 1, 5, g);
                            tc_fence_after();
                        }
                        if (s > 0 && elect_one()) {
                            const int k0 = (int)(who ? 1 - p : p) * 4 + b * 2;
#pragma unroll
                            for (int k = k0; k < k0 + 2; ++k) {
                                const uint32_t wr = w_base + (uint32_t)(((KC + (k >> 2)) * 2) * FP_TILE) + (k & 3) * 32;
#pragma unroll
                                for (int blk = 0; blk < 2; ++blk) {
                                    const uint64_t b_hi = umma_desc_k_sw128(wr + blk * BB), b_lo = umma_desc_k_sw128(wr + FP_TILE + blk * BB);
                                    const uint32_t d = blk ? d1 : d0;
                                    if constexpr (RF8) {
                                        umma_f8_ts_pair(d, tmem_base + H_LO + k * 8, b_lo, idesc, 1);
                                    } else {
                                        umma_f16_ts_pair(d, tmem_base + H_LO + k * 8, b_hi, idesc, 1);
                                        umma_f16_ts_pair(d, tmem_base + H_HI + k * 8, b_lo, idesc, 1);
                                    }
                                    umma_f16_ts_pair(d, tmem_base + H_HI + k * 8, b_hi, idesc, 1);
                                }
                            }
                        }
                        __syncwarp();
                    };
                    // static interleave: the epilogue delivers h_t in quarters (own block 0, sibling's block 0, own block 1, sibling's
                    // block 1); each quarter's 6 recurrent MMAs are queued as soon as it lands, projection chunks fill the gaps
                    TR(0, g, 0);
                    proj(0); proj(1);
                    TR(0, g, 1);
                    rec(0, 0);
                    TR(0, g, 2);
                    proj(2);
                    if constexpr (NP == 2) rec(0, 1);
                    TR(0, g, 3);
                    if constexpr (KC > 3) proj(3);
                    TR(0, g, 4);
                    rec(1, 0);
                    TR(0, g, 5);
                    if constexpr (NP == 2) rec(1, 1);
                    if (elect_one()) umma_commit_mask(acc_ready, mask);
                    TR(0, g, 6);
                    __syncwarp();
                }
        }
    } else {
        // ===================== epilogue: warps 2..17; TMEM lane quarter = warp % 4; 64-column group = (warp - 2) / 4 =====================
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;                   // pair-local units b*32 + cg*8 .. +8 for block b = 0, 1
        const int row = q * 32 + lane;
        const uint32_t sb = smem_u32(s_bias) + (uint32_t)(cg * 32) * 4;   // block b at + b*512 B
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t own_col = p * 32 + cg * 4;         // this thread's units inside the h_hi / h_lo column ranges (block b at + b*16)
        const int ew = warp - 2;                           // epilogue warp index = index of its exchange barriers
        const uint32_t peer_col = (1 - p) * 32 + cg * 4;
        const uint32_t sib = NP == 2 ? (rank ^ 2u) : rank;   // sibling CTA: other pair, same windows
        const uint32_t xbar = mapa_u32(smem_u32(&xfull[ew]), sib);
        // landing zone / staging layout [cg][part][half][row] x 8 B: slot (part, half) of this lane at x2src + (part*2 + half) * 128 * 8;
        // a warp's half of a part is 256 contiguous bytes (one bulk copy)
        const uint32_t x2src = smem_u32(s_xch) + (uint32_t)((cg * 4) * 128 + row) * 8;
        const uint32_t x2out = x2src + FP_XCH_BYTES;
        const uint32_t x2out_w = smem_u32(s_xch) + FP_XCH_BYTES + (uint32_t)((cg * 4) * 128 + q * 32) * 8;
        const uint32_t x2dst_w = mapa_u32(smem_u32(s_xch) + (uint32_t)((cg * 4) * 128 + q * 32) * 8, sib);
        if (NP == 2 && lane == 0) mbar_arrive_expect_tx(&xfull[ew], 1024);                   // phase 0
        // receive exchange phase k: the sibling warp's 8 units x (hi, lo) of our rows -> TMEM columns of the other pair's units
        auto xch_recv = [&](uint32_t k) {
            FP_WAIT(&xfull[ew], k & 1, 7, k);
            if (lane == 0) mbar_arrive_expect_tx(&xfull[ew], 1024);                          // arm the next phase
            uint4 a, b2;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(x2src) : "memory");
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a.z), "=r"(a.w) : "r"(x2src + 128 * 8) : "memory");
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(b2.x), "=r"(b2.y) : "r"(x2src + 256 * 8) : "memory");
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(b2.z), "=r"(b2.w) : "r"(x2src + 384 * 8) : "memory");
            tmem_st_32x4(lane_addr + H_HI + peer_col + (k & 1) * 16, a);
            tmem_st_32x4(lane_addr + H_LO + peer_col + (k & 1) * 16, b2);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_remote(&xfree[ew], sib);
                mbar_arrive_remote(&hq[(k & 1) * 2 + 1], leader);
            }
        };
        const float zs02 = 0.2f * acc_scale;
        uint32_t g = 0;
        for (int64_t tp = cl0; tp < n_pairs; tp += cl_stride) {
            const int64_t wtile = min(tp * 2 + (int64_t)r, ntw - 1);     // odd tile count: the last peer repeats the last tile
            const int64_t w = wtile * 128 + row;
            float c[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) c[j] = 0.f;
            for (int s = 0; s < T; ++s, ++g) {
                const int t = dir ? (T - 1 - s) : s;
                FP_WAIT(acc_ready, g & 1, 6, g);
                tc_fence_after();
                if (warp == 2) TR(1, g, 0);
                uint32_t v0[32];
                tmem_ld_32x32(lane_addr + (uint32_t)(cg * 32), v0);
                tmem_ld_wait();
                tc_fence_before();                        // our tcgen05.ld of block 0 precede the next step's MMAs into it
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(drained, leader);
                if (warp == 2) TR(1, g, 1);
                const int64_t orow = ((int64_t)t * nwp + w) * out_ld + dir * UT;     // padded rows are written too (finite, never read as windows)
                __half* oh = out_hi + orow + p * 64 + cg * 8;
                __half* ol = out_lo + orow + p * 64 + cg * 8;
                // exchange in halves of 4 units, each handed to the bulk-copy engine as soon as it exists (2 x 256 B per warp): the DSMEM
                // traffic is spread over the arithmetic instead of arriving as a 16 KB burst at the end of a block
                // ph / pl: the operands of this kernel's own recurrence (TMEM, exchange): fp16 hi + fp16 lo, or with RF8 fp16 hi + 8-bit copies.
                // Global memory gets the hi part and the lo format the CONSUMER reads (OUT8: 8-bit copies, else fp16 lo): the same
                // registers when both agree, else gl
                uint32_t ph[4], pl[4];
                constexpr bool GL_OTHER = RF8 != OUT8;
                uint32_t gl[GL_OTHER ? 4 : 1];
                auto send_half = [&](int b, int half, const float* hv4, uint32_t wait_k) {      // wait_k != 0: first half of exchange phase wait_k
                    uint2 hi2, lo2;
                    if constexpr (RF8 || OUT8) {
                        uint2 l16, q2;
                        pack_h4_f8<GL_OTHER>(hv4, hi2, l16, q2);
                        if constexpr (RF8) {
                            lo2 = q2;
                            if constexpr (GL_OTHER) { gl[half * 2] = l16.x; gl[half * 2 + 1] = l16.y; }
                        } else {
                            lo2 = l16;
                            gl[half * 2] = q2.x; gl[half * 2 + 1] = q2.y;
                        }
                    } else {
                        pack_h4(hv4, hi2, lo2);
                    }
                    ph[half * 2] = hi2.x; ph[half * 2 + 1] = hi2.y; pl[half * 2] = lo2.x; pl[half * 2 + 1] = lo2.y;
                    if constexpr (NP == 2) {
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(x2out + (0 * 2 + half) * 128 * 8), "r"(hi2.x), "r"(hi2.y) : "memory");
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(x2out + (1 * 2 + half) * 128 * 8), "r"(lo2.x), "r"(lo2.y) : "memory");
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            if (wait_k) FP_WAIT(&xfree[ew], (wait_k - 1) & 1, 8, wait_k);
                            bulk_copy_to_cluster(x2dst_w + (0 * 2 + half) * 128 * 8, x2out_w + (0 * 2 + half) * 128 * 8, 256, xbar);
                            bulk_copy_to_cluster(x2dst_w + (1 * 2 + half) * 128 * 8, x2out_w + (1 * 2 + half) * 128 * 8, 256, xbar);
                            tma_store_commit();
                        }
                    }
                };
                auto publish2 = [&](int b) {
                    const uint4 phi = make_uint4(ph[0], ph[1], ph[2], ph[3]), plo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    tmem_st_32x4(lane_addr + H_HI + own_col + b * 16, phi);
                    tmem_st_32x4(lane_addr + H_LO + own_col + b * 16, plo);
                    *reinterpret_cast<uint4*>(oh + b * 32) = phi;
                    if constexpr (GL_OTHER) *reinterpret_cast<uint4*>(ol + b * 32) = make_uint4(gl[0], gl[1], gl[2], gl[3]);
                    else *reinterpret_cast<uint4*>(ol + b * 32) = plo;
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(&hq[b * 2], leader);
                };
                float hv[8];
                cell_units<0, 4>(v0, sb, &c[0], hv, acc_scale, zs02);
                send_half(0, 0, hv, 2 * g);                  // phase 2g: the sibling has consumed phase 2g-1 (wait_k = 0 for the very first)
                cell_units<4, 8>(v0, sb, &c[0], hv, acc_scale, zs02);
                send_half(0, 1, hv + 4, 0);
                uint32_t v1[32];
                tmem_ld_32x32(lane_addr + (uint32_t)(128 + (g & 1) * 128 + cg * 32), v1);
                publish2(0);
                if (warp == 2) TR(1, g, 2);
                tmem_ld_wait();
                cell_units<0, 4>(v1, sb + 512, &c[8], hv, acc_scale, zs02);
                if (warp == 2) TR(1, g, 3);
                if constexpr (NP == 2) xch_recv(2 * g);
                if (warp == 2) TR(1, g, 4);
                send_half(1, 0, hv, 2 * g + 1);              // phase 2g+1 after the sibling has consumed phase 2g
                cell_units<4, 8>(v1, sb + 512, &c[8], hv, acc_scale, zs02);
                send_half(1, 1, hv + 4, 0);
                publish2(1);
                if (warp == 2) TR(1, g, 5);
                if constexpr (NP == 2) xch_recv(2 * g + 1);
                if (warp == 2) TR(1, g, 6);
            }
        }
        if (NP == 2 && lane == 0) tma_store_wait_all();       // our last bulk copies are complete before this CTA retires
    }
    tc_fence_before();
    __syncthreads();
#ifdef NRV_TRACE
    if (threadIdx.x == 0 && cluster_ctarank() == 0 && blockIdx.x / (2 * Cfg::NP) < 64) g_tr_cl[UT == 128][blockIdx.y][blockIdx.x / (2 * Cfg::NP)][1] = gtimer();
    if (trace_on && threadIdx.x == 0 && atomicAdd(&g_tr_launch[UT == 128], 1) == 3) {
        // clusters of the PREVIOUS launch of this instantiation (all finished): start / end relative to the first start
        unsigned long long t0 = ~0ull;
        for (int d = 0; d < 2; ++d) for (int c = 0; c < (int)(gridDim.x / (2 * Cfg::NP)); ++c) if (g_tr_cl[UT == 128][d][c][0] < t0 && !(d == 0 && c == 0)) t0 = g_tr_cl[UT == 128][d][c][0];
        for (int d = 0; d < 2; ++d) for (int c = 0; c < (int)(gridDim.x / (2 * Cfg::NP)) && c < 64; ++c)
            printf("CL%d dir %d cluster %2d sm %3llu start %8lld end %8lld\n", UT, d, c, g_tr_cl[UT == 128][d][c][2],
                   (long long)(g_tr_cl[UT == 128][d][c][0] - t0), (long long)(g_tr_cl[UT == 128][d][c][1] - t0));
        for (int st = 0; st < 40; ++st) {
            printf("TR%d step %2d mma:", UT, st);
            for (int e = 0; e < 7; ++e) printf(" %7lld", g_tr[UT == 128][0][st][e]);
            printf("  epi:");
            for (int e = 0; e < 7; ++e) printf(" %7lld", g_tr[UT == 128][1][st][e]);
            printf("\n");
        }
    }
#endif
    cluster_sync_all();            // the leaders' MMAs read their peers' shared memory: nobody leaves early
    if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

template <int KIN, int UT, bool PF8, bool RF8, bool OUT8>
static int launch_fused_pair_t(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T,
                               int num_sms, cudaStream_t st) {
    using Cfg = FpCfg<KIN, UT>;
    constexpr int CS = 2 * Cfg::NP;
    constexpr bool F8W = PF8 || RF8;                  // the layer's weights carry the scale 2^S (LstmLayerDev::f8_*)
    CUtensorMap txh, txl;
    if (!make_tmap_f16_k64(&txh, x_hi, (int64_t)T * nwp, KIN, 128)) return -2;
    if (PF8) {   // x_lo = the producing layer's 8-bit copies: rows of 2 KIN bytes, 128-byte boxes (= the 64 units of an fp16 K-chunk)
        if (!make_tmap_u8_k128(&txl, x_lo, (int64_t)T * nwp, 2 * KIN, 128)) return -2;
    } else {
        if (!make_tmap_f16_k64(&txl, x_lo, (int64_t)T * nwp, KIN, 128)) return -2;
    }
    CUtensorMap tsh = txh, tsl = txl;                 // per-base CNN-feature table (total_rnn1 only)
    const bool sig_table = UT == 128 && io.tile_base && io.sf_hi && io.sf_lo && io.sf_rows > 0;
    if (sig_table && (!make_tmap_f16_k64(&tsh, io.sf_hi, io.sf_rows, 64, 128) || !make_tmap_f16_k64(&tsl, io.sf_lo, io.sf_rows, 64, 128))) return -2;
    auto kern = lstm_fused_pair_kernel<KIN, UT, PF8, RF8, OUT8>;
    static PerDevice per_dev;                     // co-resident clusters on the current device (per template instance)
    int& max_clusters = per_dev.cur();
    if (!max_clusters) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e != cudaSuccess) { fprintf(stderr, "[nrv] lstm_fused_pair<%d,%d>: smem attribute: %s\n", KIN, UT, cudaGetErrorString(e)); return -3; }
        cudaLaunchConfig_t qc = {};
        qc.gridDim = dim3(CS * 64, 2); qc.blockDim = dim3(FP_THREADS); qc.dynamicSmemBytes = Cfg::SMEM;
        cudaLaunchAttribute qa; qa.id = cudaLaunchAttributeClusterDimension; qa.val.clusterDim.x = CS; qa.val.clusterDim.y = 1; qa.val.clusterDim.z = 1;
        qc.attrs = &qa; qc.numAttrs = 1;
        int n = 0;
        const cudaError_t eo = cudaOccupancyMaxActiveClusters(&n, kern, &qc);
        if (eo != cudaSuccess || n < 2) {
            fprintf(stderr, "[nrv] lstm_fused_pair<%d,%d>: occupancy query: %s (n = %d)\n", KIN, UT, cudaGetErrorString(eo), n);
            cudaGetLastError(); n = num_sms / CS;
        }
        max_clusters = n;
        if (getenv("NRV_VERBOSE")) fprintf(stderr, "[nrv] lstm_fused_pair<%d,%d,%d,%d,%d>: cluster %d, %d co-resident clusters\n", KIN, UT, (int)PF8, (int)RF8, (int)OUT8, CS, n);
    }
    const int64_t n_pairs = ((nwp >> 7) + 1) / 2;
    const int per_dir = std::max(1, max_clusters / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CS * std::min<int64_t>(n_pairs, per_dir)), 2);
    cfg.blockDim = dim3(FP_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CS; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    // scaled fp16 weights whenever one of the products runs F8 (one accumulator scale); the lo parts are fp16 for an fp16 x 3 product, the
    // interleaved 8-bit copies for an F8 one
    const __half* wk_hi = F8W ? (const __half*)L.f8_wk_hi : (const __half*)L.pb_hi;
    const __half* wk_lo = PF8 ? reinterpret_cast<const __half*>(L.f8_wk8) : (F8W ? (const __half*)L.f8_wk_lo : (const __half*)L.pb_lo);
    const __half* wr_hi = F8W ? (const __half*)L.f8_wr_hi : (const __half*)L.rt_hi;
    const __half* wr_lo = RF8 ? reinterpret_cast<const __half*>(L.f8_wr8) : (F8W ? (const __half*)L.f8_wr_lo : (const __half*)L.rt_lo);
    if (!wk_hi || !wk_lo || !wr_hi || !wr_lo) return -1;
    // NRV_L2HINT=1: eviction hints on total_rnn2's input tiles (experiment switch, see the kernel parameter)
    static const int l2hint_env = getenv("NRV_L2HINT") ? atoi(getenv("NRV_L2HINT")) : 0;
    const int l2hint = (UT == 64) ? l2hint_env : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, wk_hi, wk_lo, wr_hi, wr_lo, (const float*)L.bias_tc, txh, txl, io.out_hi, io.out_lo,
                                             io.out_ld, nwp, T, F8W ? L.f8_acc_scale : 1.0f, tsh, tsl, sig_table ? io.tile_base : (const int32_t*)nullptr,
                                             l2hint);
    if (e != cudaSuccess) {
        fprintf(stderr, "[nrv] lstm_fused_pair<%d,%d>: launch (grid %u x 2, cluster %d, %zu B smem): %s\n", KIN, UT, cfg.gridDim.x, CS,
                (size_t)Cfg::SMEM, cudaGetErrorString(e));
        return -4;
    }
    return 1;
}

// total_rnn2: x = total_rnn1's output (K = 256), 64 units.  f8 != 0: both products with e4m3 correction passes; x_lo = the 8-bit copies
int launch_lstm_fused_pair64(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T, int num_sms,
                             cudaStream_t st, int f8) {
    if (nwp <= 0) return 0;
    if (L.u != 64 || L.in_a != 256 || !L.rt_hi || !L.pb_hi || !L.bias_tc || !io.out_hi || !io.out_lo || (nwp & 127) || (io.out_ld & 7)) return -1;
    if (f8) return launch_fused_pair_t<256, 64, true, true, false>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
    return launch_fused_pair_t<256, 64, false, false, false>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
}
// total_rnn1: x = [read_rnn11 | CNN features] (K = 192), 128 units, cluster of 4.  io.out_f8: the consumer (total_rnn2) reads 8-bit copies;
// io.rec_f8: the recurrence runs with e4m3 correction passes (the projection stays fp16 x 3: its inputs need no copies)
int launch_lstm_fused_pair128(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T, int num_sms,
                              cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 128 || !L.rt_hi || !L.pb_hi || !L.bias_tc || !io.out_hi || !io.out_lo || (nwp & 127) || (io.out_ld & 7)) return -1;
    if (io.rec_f8 && io.out_f8) return launch_fused_pair_t<192, 128, false, true, true>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
    if (io.rec_f8) return launch_fused_pair_t<192, 128, false, true, false>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
    if (io.out_f8) return launch_fused_pair_t<192, 128, false, false, true>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
    return launch_fused_pair_t<192, 128, false, false, false>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
}

}  // namespace nrv
