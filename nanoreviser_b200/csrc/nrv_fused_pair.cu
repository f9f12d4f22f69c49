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
// Pipeline (one accumulator, drained fast):
//     tensor core :  rec(s) | ............ proj(s+1): 3 x KIN/16 MMAs ............ | rec(s+1) | proj(s+2) ...
//     epilogue    :         | drain | gates, c, h -> TMEM + HBM (+ exchange)       |          | drain | ...
// 16 epilogue warps: each thread tcgen05.ld's its 64 accumulator columns into registers and signals `drained`; the
// projection MMAs of the next step (which need no h) start right then and run under the gate arithmetic; `h_ready`
// releases the recurrent MMAs.  x_t is read exactly once; the tensor pipe only idles during the drain.
//
// total_rnn1's exchange goes through the layer's own output: each thread stores its 16 units of h_t to the activation
// tensor (which the next layer needs anyway), fences, and arrives on the sibling CTA's `xfull` barrier; the sibling thread
// that owns the same window reads the peer's 16 units back (L2) and tcgen05.st's them into its TMEM.  No shared memory.
#include <algorithm>

#include "nrv_cell.cuh"
#include "nrv_common.cuh"
#include "nrv_tc.cuh"

namespace nrv {

using namespace tc;

bool make_tmap_f16_k64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);   // nrv_gemm.cu

constexpr int FP_EPI_WARPS = 16;                      // 4 per TMEM lane quarter; 16 units (64 gate columns) per thread
constexpr int FP_THREADS = 64 + 32 * FP_EPI_WARPS;    // warp 0: MMA issue (pair leader); warp 1: TMA producer; warps 2..17: epilogue
constexpr int FP_TILE = 128 * 64 * 2;                 // 16 KB: [128 rows][64 halves], K-major, 128-byte swizzle
constexpr int FP_STAGES = 4;
constexpr uint32_t FP_H_COL = 256;                    // TMEM: accumulator [0, 256); h_hi [256, 256 + UT/2); h_lo [.., 256 + UT)

template <int KIN, int UT>
struct FpCfg {
    static constexpr int KC = KIN / 64, RC = UT / 64, NP = UT / 64;
    static constexpr int W_BYTES = (KC + RC) * 2 * FP_TILE;
    static constexpr size_t SMEM = (size_t)W_BYTES + FP_STAGES * FP_TILE + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*alignment*/;
};

__device__ __forceinline__ void umma_commit_mask(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
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
#define FP_WAIT_CL(bar, parity, tag, g) fp_wait_dbg<true>(bar, parity, tag, g)
#else
#define FP_WAIT(bar, parity, tag, g) mbar_wait(bar, parity)
#define FP_WAIT_CL(bar, parity, tag, g) mbar_wait_cluster(bar, parity)
#endif
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// One 32-column block (8 units) of the cell with the bias read from shared memory unit by unit (keeps the 64 drained
// accumulator registers + 16 cell states close to the 96-register budget of a 576-thread CTA: warps are allocated in fours, so 20 x 32 x 96).
__device__ __forceinline__ void cell_block_bias(const uint32_t (&v)[32], uint32_t sbias, float* c8, uint4& phi, uint4& plo) {
    float hv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 b = ld_shared_f4(sbias + j * 16);
        const float zi = b.x + __uint_as_float(v[4 * j + 0]), zf = b.y + __uint_as_float(v[4 * j + 1]);
        const float zc = b.z + __uint_as_float(v[4 * j + 2]), zo = b.w + __uint_as_float(v[4 * j + 3]);
        const float ig = hsig(zi), fg = hsig(zf), gg = tanh_fast(zc), og = hsig(zo);
        const float cn = fmaf(fg, c8[j], ig * gg);
        c8[j] = cn;
        hv[j] = og * tanh_fast(cn);
    }
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 hi = __floats2half2_rn(hv[2 * p], hv[2 * p + 1]);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(hv[2 * p] - hf.x, hv[2 * p + 1] - hf.y);
        ph[p] = half2_bits(hi); pl[p] = half2_bits(lo);
    }
    phi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    plo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

template <int KIN, int UT>
__global__ void __launch_bounds__(FP_THREADS, 1)
lstm_fused_pair_kernel(const __half* __restrict__ wk_hi, const __half* __restrict__ wk_lo, const __half* __restrict__ wr_hi,
                       const __half* __restrict__ wr_lo, const float* __restrict__ bias,
                       const __grid_constant__ CUtensorMap tm_x_hi, const __grid_constant__ CUtensorMap tm_x_lo,
                       __half* __restrict__ out_hi, __half* __restrict__ out_lo, int out_ld, int64_t nwp, int T) {
    using Cfg = FpCfg<KIN, UT>;
    constexpr int KC = Cfg::KC, NP = Cfg::NP, CS = 2 * NP;
    constexpr int NT = 4 * UT;                              // gate columns per direction
    constexpr uint32_t H_HI = FP_H_COL, H_LO = FP_H_COL + UT / 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                    // [Wk chunk 0..KC-1 | Wr chunk 0..RC-1][hi | lo][128 rows][64]
    uint8_t* s_ring = smem + Cfg::W_BYTES;                  // [stage][128 rows][64]
    float* s_bias = reinterpret_cast<float*>(s_ring + FP_STAGES * FP_TILE);    // [256]: this pair's gate columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
    uint64_t* full = bars;                                  // [4] pair leader's copy: 1 arrive + 32 KB tx (both CTAs' tiles)
    uint64_t* empty = bars + FP_STAGES;                     // [4] both CTAs of the pair: multicast commit
    uint64_t* acc_ready = bars + 2 * FP_STAGES;             // both CTAs of the pair: multicast commit
    uint64_t* drained = acc_ready + 1;                      // pair leader's copy: 32 arrivals (16 epilogue warps x 2 CTAs)
    uint64_t* h_ready = acc_ready + 2;                      // pair leader's copy: 32 arrivals
    uint64_t* xfull = acc_ready + 3;                        // NP == 2: 16 arrivals from the sibling CTA (other pair, same windows)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        mbar_init(h_ready, 2 * FP_EPI_WARPS);
        mbar_init(xfull, FP_EPI_WARPS);
        fence_mbar_init();
        tma_prefetch_desc(&tm_x_hi); tma_prefetch_desc(&tm_x_lo);
    }
    if (warp == 0) tmem_alloc_pair(tmem_slot, 512);
    if (threadIdx.x < 256) s_bias[threadIdx.x] = __ldg(bias + dir * NT + p * 256 + threadIdx.x);
    {   // resident weights: this CTA's 128 gate columns (global rows dir*NT + p*256 + r*128 + row), all of K, hi and lo
        const size_t grow0 = (size_t)dir * NT + p * 256 + r * 128;
        constexpr int CK = KIN / 8, CR = UT / 8;            // 16-byte chunks per row
        for (int i = threadIdx.x; i < 2 * 128 * CK; i += FP_THREADS) {
            const int c = i % CK, row = (i / CK) & 127, part = i / (CK * 128);
            const __half* src = (part ? wk_lo : wk_hi) + (grow0 + row) * KIN;
            *reinterpret_cast<uint4*>(s_w + (size_t)((c >> 3) * 2 + part) * FP_TILE + sw128_offset(row, c & 7)) =
                __ldg(reinterpret_cast<const uint4*>(src) + c);
        }
        for (int i = threadIdx.x; i < 2 * 128 * CR; i += FP_THREADS) {
            const int c = i % CR, row = (i / CR) & 127, part = i / (CR * 128);
            const __half* src = (part ? wr_lo : wr_hi) + (grow0 + row) * UT;
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
                for (int s = 0; s < T; ++s) {
                    const int t = dir ? (T - 1 - s) : s;
                    const int grow = (int)(t * nwp + wtile * 128);
                    for (int i = 0; i < 2 * KC; ++i) {                     // (K-chunk, part): lo tile first, then hi
                        FP_WAIT(&empty[stage], phase ^ 1, 1, (uint32_t)s);
                        if (r == 0) mbar_arrive_expect_tx(&full[stage], 2 * FP_TILE);
                        tma_load_2d_pair(s_ring + (size_t)stage * FP_TILE, (i & 1) ? &tm_x_hi : &tm_x_lo, &full[stage], leader,
                                         (i >> 1) * 64, grow);
                        if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 0) {
        // ===================== MMA issue (leader CTA of every pair) =====================
        if (r == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(256, 256);
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
                    for (int kc = 0; kc < KC; ++kc) {
                        const uint32_t wb_hi = w_base + (uint32_t)((kc * 2 + 0) * FP_TILE), wb_lo = w_base + (uint32_t)((kc * 2 + 1) * FP_TILE);
                        FP_WAIT(&full[stage], phase, 3, g);                   // x_lo(kc)
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t xa = ring_base + (uint32_t)(stage * FP_TILE);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_f16_ss_pair(tmem_base, umma_desc_k_sw128(xa + k * 32), umma_desc_k_sw128(wb_hi + k * 32), idesc,
                                                 (kc | k) != 0);
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
                                umma_f16_ss_pair(tmem_base, a_hi, umma_desc_k_sw128(wb_lo + k * 32), idesc, 1);
                                umma_f16_ss_pair(tmem_base, a_hi, umma_desc_k_sw128(wb_hi + k * 32), idesc, 1);
                            }
                            umma_commit_mask(&empty[stage], mask);
                        }
                        __syncwarp();
                        if (++stage == FP_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (g > 0) {                                          // h of the previous step is in TMEM (both CTAs)
                        FP_WAIT(h_ready, (g - 1) & 1, 5, g);
                        tc_fence_after();
                    }
                    if (elect_one()) {
                        if (s > 0) {
#pragma unroll
                            for (int k = 0; k < UT / 16; ++k) {
                                const uint32_t wr = w_base + (uint32_t)(((KC + (k >> 2)) * 2) * FP_TILE) + (k & 3) * 32;
                                const uint64_t b_hi = umma_desc_k_sw128(wr), b_lo = umma_desc_k_sw128(wr + FP_TILE);
                                umma_f16_ts_pair(tmem_base, tmem_base + H_LO + k * 8, b_hi, idesc, 1);
                                umma_f16_ts_pair(tmem_base, tmem_base + H_HI + k * 8, b_lo, idesc, 1);
                                umma_f16_ts_pair(tmem_base, tmem_base + H_HI + k * 8, b_hi, idesc, 1);
                            }
                        }
                        umma_commit_mask(acc_ready, mask);
                    }
                    __syncwarp();
                }
        }
    } else {
        // ===================== epilogue: warps 2..17; TMEM lane quarter = warp % 4; 64-column group = (warp - 2) / 4 =====================
        const int q = warp & 3;
        const int cg = (warp - 2) >> 2;                   // units p*64 + cg*16 .. +16
        const int row = q * 32 + lane;
        const uint32_t sb = smem_u32(s_bias) + (uint32_t)(cg * 64) * 4;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t own_col = p * 32 + cg * 8;         // this thread's 16 units inside the h_hi / h_lo column ranges
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
                uint32_t v0[32], v1[32];
                tmem_ld_32x32(lane_addr + (uint32_t)(cg * 64), v0);
                tmem_ld_32x32(lane_addr + (uint32_t)(cg * 64 + 32), v1);
                tmem_ld_wait();
                tc_fence_before();                        // our tcgen05.ld precede the next step's MMAs
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(drained, leader);
                uint4 phi0, plo0, phi1, plo1;
                cell_block_bias(v0, sb, &c[0], phi0, plo0);
                cell_block_bias(v1, sb + 128, &c[8], phi1, plo1);
                tmem_st_32x4(lane_addr + H_HI + own_col, phi0);
                tmem_st_32x4(lane_addr + H_HI + own_col + 4, phi1);
                tmem_st_32x4(lane_addr + H_LO + own_col, plo0);
                tmem_st_32x4(lane_addr + H_LO + own_col + 4, plo1);
                const int64_t orow = ((int64_t)t * nwp + w) * out_ld + dir * UT;     // padded rows are written too (finite, never read as windows)
                {
                    __half* oh = out_hi + orow + p * 64 + cg * 16;
                    __half* ol = out_lo + orow + p * 64 + cg * 16;
                    *reinterpret_cast<uint4*>(oh) = phi0; *reinterpret_cast<uint4*>(oh + 8) = phi1;
                    *reinterpret_cast<uint4*>(ol) = plo0; *reinterpret_cast<uint4*>(ol + 8) = plo1;
                }
                if constexpr (NP == 2) {
                    // exchange through the layer output: publish our 16 units, fetch the sibling pair's 16 units of the same window
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(xfull, rank ^ 2u);
                    FP_WAIT_CL(xfull, g & 1, 7, g);
                    const __half* ph = out_hi + orow + (1 - p) * 64 + cg * 16;
                    const __half* pl = out_lo + orow + (1 - p) * 64 + cg * 16;
                    const uint4 a0 = ld_cg_u4(ph), a1 = ld_cg_u4(ph + 8), b0 = ld_cg_u4(pl), b1 = ld_cg_u4(pl + 8);
                    const uint32_t peer_col = (1 - p) * 32 + cg * 8;
                    tmem_st_32x4(lane_addr + H_HI + peer_col, a0);
                    tmem_st_32x4(lane_addr + H_HI + peer_col + 4, a1);
                    tmem_st_32x4(lane_addr + H_LO + peer_col, b0);
                    tmem_st_32x4(lane_addr + H_LO + peer_col + 4, b1);
                }
                tmem_st_wait();
                tc_fence_before();                        // our tcgen05.st of h precede the recurrent MMAs of the next step
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(h_ready, leader);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the leaders' MMAs read their peers' shared memory: nobody leaves early
    if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

template <int KIN, int UT>
static int launch_fused_pair_t(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T,
                               int num_sms, cudaStream_t st) {
    using Cfg = FpCfg<KIN, UT>;
    constexpr int CS = 2 * Cfg::NP;
    CUtensorMap txh, txl;
    if (!make_tmap_f16_k64(&txh, x_hi, (int64_t)T * nwp, KIN, 128) || !make_tmap_f16_k64(&txl, x_lo, (int64_t)T * nwp, KIN, 128)) return -2;
    auto kern = lstm_fused_pair_kernel<KIN, UT>;
    static int max_clusters = 0;                  // co-resident clusters on this device (per template instance)
    if (!max_clusters) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
        if (e == cudaSuccess && CS > 2) e = cudaSuccess;
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
        if (getenv("NRV_VERBOSE")) fprintf(stderr, "[nrv] lstm_fused_pair<%d,%d>: cluster %d, %d co-resident clusters\n", KIN, UT, CS, n);
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
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const __half*)L.pb_hi, (const __half*)L.pb_lo, (const __half*)L.rt_hi,
                                             (const __half*)L.rt_lo, (const float*)L.bias_tc, txh, txl, io.out_hi, io.out_lo, io.out_ld, nwp, T);
    if (e != cudaSuccess) {
        fprintf(stderr, "[nrv] lstm_fused_pair<%d,%d>: launch (grid %u x 2, cluster %d, %zu B smem): %s\n", KIN, UT, cfg.gridDim.x, CS,
                (size_t)Cfg::SMEM, cudaGetErrorString(e));
        return -4;
    }
    return 1;
}

// total_rnn2: x = total_rnn1's output (K = 256), 64 units
int launch_lstm_fused_pair64(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T, int num_sms,
                             cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 64 || L.in_a != 256 || !L.rt_hi || !L.pb_hi || !L.bias_tc || !io.out_hi || !io.out_lo || (nwp & 127) || (io.out_ld & 7)) return -1;
    return launch_fused_pair_t<256, 64>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
}
// total_rnn1: x = [read_rnn11 | CNN features] (K = 192), 128 units, cluster of 4
int launch_lstm_fused_pair128(const LstmLayerDev& L, const __half* x_hi, const __half* x_lo, const LstmIo& io, int64_t nwp, int T, int num_sms,
                              cudaStream_t st) {
    if (nwp <= 0) return 0;
    if (L.u != 128 || !L.rt_hi || !L.pb_hi || !L.bias_tc || !io.out_hi || !io.out_lo || (nwp & 127) || (io.out_ld & 7)) return -1;
    return launch_fused_pair_t<192, 128>(L, x_hi, x_lo, io, nwp, T, num_sms, st);
}

}  // namespace nrv
