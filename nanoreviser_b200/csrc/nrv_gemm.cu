// Time-batched input-projection GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA):
//
//     C[M][N] (fp32) = A[M][K] . B[N][K]^T (+ bias[N])
//
// A (activations) and B (weights, stored transposed = K-major) are both given as split-precision fp16
// pairs (hi, lo); three tcgen05.mma.kind::f16 passes  hi*hi + lo*hi + hi*lo  accumulate in fp32 in TMEM,
// which reproduces the fp32 contraction to ~2^-22 relative -- single-pass fp16/bf16/tf32 cannot hold the
// 1e-3 logit tolerance (measured again in DESIGN.md section 4).
//
// Used for the input projections of the Bi-LSTM layers total_rnn1 / total_rnn2 (lstmmodel.py:49,51):
// rows are (window, timestep) pairs, columns are the interleaved gate pre-activations of both directions.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: per 64-wide K chunk, 4 bulk-tensor loads (A_hi, A_lo, B_hi, B_lo) into a
//               2-stage shared-memory ring (128-byte swizzle), completion on an mbarrier
//   warp 1      allocates TMEM (2 x 256 fp32 columns = double-buffered 128x256 accumulator) and issues
//               the MMAs from one elected lane; tcgen05.commit releases ring slots / publishes accumulators
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> (+bias) -> global, one TMEM
//               lane quarter per warp; overlaps the next tile's MMAs through the second accumulator
#include <string.h>

#include <algorithm>

#include "nrv_common.cuh"
#include "nrv_tc.cuh"

namespace nrv {

using namespace tc;

constexpr int G_TM = 128, G_KC = 64;
constexpr int G_A_BYTES = G_TM * G_KC * 2;            // 16 KB (one of hi / lo)
constexpr int G_EPI_WARPS = 8;                        // 2 per TMEM lane quarter: column halves of the tile
constexpr int G_THREADS = 64 + 32 * G_EPI_WARPS;

constexpr int G_A2_BYTES = 4 * G_A_BYTES;            // fused head: relu(Dense128) tile [128][128] as fp16 (hi, lo), 2 K-chunks each
constexpr int G_W2_BYTES = 4 * 32 * 128;              // fused head: W2^T [32][128] as fp16 (hi, lo), 2 K-chunks each (4 x 4 KB)

template <int TN, bool FUSE2>
struct GemmCfg {
    static constexpr int B_BYTES = TN * G_KC * 2;                      // 32 KB (TN = 256) / 16 KB (TN = 128)
    static constexpr int STAGE_BYTES = 2 * G_A_BYTES + 2 * B_BYTES;    // 96 KB / 64 KB
    static constexpr int STAGES = (TN == 256 || FUSE2) ? 2 : 3;
    static constexpr int TMEM_COLS = FUSE2 ? 512 : 2 * TN;             // double-buffered accumulator (+ 32 columns for the fused head)
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + (FUSE2 ? G_A2_BYTES + G_W2_BYTES : 0) + 1024 /*alignment slack*/ +
                                   256 /*barriers*/ + 768 /*FUSE2 biases*/;
};

struct GemmOut {
    float* c;
    const float* bias;      // [N] or nullptr
    int mode;               // 0: C[r][N] row-major.  1: zin tiles C[dir][r / 128][col / 4][128][4] with
                            //    dir = column / n_per_dir (M must be a multiple of 128)
    int T;
    int64_t nw;
    int n_per_dir;
    int relu;               // apply max(x, 0) after the bias (dense heads)
    // mode 2 (dense heads, TN = N = 128, FUSE2 kernel): the epilogue applies bias + relu, re-splits the 128-wide row into an
    // fp16 (hi, lo) A tile in shared memory and the MMA warp runs the NEXT dense layer back to back on the tensor core:
    // c[r][32] = relu(relu(acc + bias) . W2 + b2) -- the 128-wide intermediate never leaves the SM
    const __half* w2t_hi;   // W2^T [32][128]
    const __half* w2t_lo;
    const float* b2;
    float inv_in_scale;     // mode 2: 1 / (power-of-two scale of the A operand), applied to the accumulator before the bias
};

template <int G_TN, bool FUSE2>
__global__ void __launch_bounds__(G_THREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  GemmOut out, int64_t M, int N, int K) {
    using Cfg = GemmCfg<G_TN, FUSE2>;
    constexpr int G_STAGES = Cfg::STAGES, G_STAGE_BYTES = Cfg::STAGE_BYTES, G_B_BYTES = Cfg::B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // swizzled tiles need 1024-byte alignment: ring stages first, then the fused head's tiles, barriers last
    uint8_t* s_a2 = smem + G_STAGES * G_STAGE_BYTES;                 // [hi|lo][kc][128 rows][64]   (FUSE2 only)
    uint8_t* s_w2 = s_a2 + (FUSE2 ? G_A2_BYTES : 0);                 // [hi|lo][kc][32 rows][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w2 + (FUSE2 ? G_W2_BYTES : 0));
    uint64_t* full = bars;                  // [G_STAGES]
    uint64_t* empty = bars + G_STAGES;      // [G_STAGES]
    uint64_t* tfull = bars + 2 * G_STAGES;  // [2]
    uint64_t* tempty = tfull + 2;           // [2]
    uint64_t* a2_ready = tempty + 2;        // fused head: epilogue -> MMA warp ("A tile of the second dense is in smem")
    uint64_t* d2_ready = a2_ready + 1;      // fused head: MMA warp -> epilogue ("second accumulator complete")
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_ready + 1);
    float* s_bias = reinterpret_cast<float*>(bars + 16);             // FUSE2: bias of the first dense [128] | b2 [32] (16-byte aligned)
    constexpr uint32_t D2_COL = 2 * G_TN;                            // TMEM columns of the fused head's accumulator

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = N / G_TN;
    const int64_t n_tiles_m = (M + G_TM - 1) / G_TM;
    const int64_t n_tiles = n_tiles_m * n_tiles_n;
    const int n_chunks = K / G_KC;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo); tma_prefetch_desc(&tm_b_hi); tma_prefetch_desc(&tm_b_lo);
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], G_EPI_WARPS); }
        mbar_init(a2_ready, G_EPI_WARPS); mbar_init(d2_ready, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    if (FUSE2) {   // W2^T [32 rows][128 halves] -> two K-major swizzled [32][64] tiles per part
        for (int i = threadIdx.x; i < 2 * 32 * 16; i += G_THREADS) {
            const int part = i >> 9, row = (i >> 4) & 31, c16 = i & 15;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(part ? out.w2t_lo : out.w2t_hi) + row * 16 + c16);
            *reinterpret_cast<uint4*>(s_w2 + (part * 2 + (c16 >> 3)) * 4096 + sw128_offset(row, c16 & 7)) = v;
        }
        if (threadIdx.x < 128) s_bias[threadIdx.x] = __ldg(out.bias + threadIdx.x);
        if (threadIdx.x < 32) s_bias[128 + threadIdx.x] = __ldg(out.b2 + threadIdx.x);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (int)((tile / n_tiles_n) * G_TM), n0 = (int)((tile % n_tiles_n) * G_TN);
                for (int kc = 0; kc < n_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = smem + stage * G_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], G_STAGE_BYTES);
                    tma_load_2d(st, &tm_a_hi, &full[stage], kc * G_KC, m0);
                    tma_load_2d(st + G_A_BYTES, &tm_a_lo, &full[stage], kc * G_KC, m0);
                    tma_load_2d(st + 2 * G_A_BYTES, &tm_b_hi, &full[stage], kc * G_KC, n0);
                    tma_load_2d(st + 2 * G_A_BYTES + G_B_BYTES, &tm_b_lo, &full[stage], kc * G_KC, n0);
                    if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16_f32(G_TM, G_TN);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t a2_phase = 0;
        bool pending_head = false;
        // second dense of the fused head for the tile whose epilogue is running: D2[128][32] = A2[128][128] . W2^T
        auto issue_head = [&]() {
            mbar_wait(a2_ready, a2_phase);
            a2_phase ^= 1;
            tc_fence_after();
            if (elect_one()) {
                constexpr uint32_t idesc2 = umma_idesc_f16_f32(G_TM, 32);
                const uint32_t a2 = smem_u32(s_a2), w2 = smem_u32(s_w2);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int kc = k >> 2, kk = k & 3;
                    const uint64_t ah = umma_desc_k_sw128(a2 + (0 * 2 + kc) * G_A_BYTES + kk * 32);
                    const uint64_t al = umma_desc_k_sw128(a2 + (1 * 2 + kc) * G_A_BYTES + kk * 32);
                    const uint64_t bh = umma_desc_k_sw128(w2 + (0 * 2 + kc) * 4096 + kk * 32);
                    const uint64_t bl = umma_desc_k_sw128(w2 + (1 * 2 + kc) * 4096 + kk * 32);
                    umma_f16_ss(tmem_base + D2_COL, al, bh, idesc2, k != 0);
                    umma_f16_ss(tmem_base + D2_COL, ah, bl, idesc2, 1);
                    umma_f16_ss(tmem_base + D2_COL, ah, bh, idesc2, 1);
                }
                umma_commit(d2_ready);
            }
            __syncwarp();
        };
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tempty[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G_TN);
            for (int kc = 0; kc < n_chunks; ++kc) {
                mbar_wait(&full[stage], phase);              // TMA bytes have landed
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t st = smem_u32(smem + stage * G_STAGE_BYTES);
                    const uint64_t a_hi = umma_desc_k_sw128(st), a_lo = umma_desc_k_sw128(st + G_A_BYTES);
                    const uint64_t b_hi = umma_desc_k_sw128(st + 2 * G_A_BYTES);
                    const uint64_t b_lo = umma_desc_k_sw128(st + 2 * G_A_BYTES + G_B_BYTES);
#pragma unroll
                    for (int k = 0; k < G_KC / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);       // +32 bytes along K inside the swizzle atom
                        umma_f16_ss(d_tmem, a_lo + adv, b_hi + adv, idesc, (kc | k) != 0);
                        umma_f16_ss(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_f16_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    umma_commit(&empty[stage]);                             // frees the ring slot when the MMAs retire
                    if (kc == n_chunks - 1) umma_commit(&tfull[acc]);       // accumulator complete
                }
                __syncwarp();
                if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            if (FUSE2) {
                if (pending_head) issue_head();              // previous tile's second dense (its epilogue ran meanwhile)
                pending_head = true;
            }
        }
        if (FUSE2 && pending_head) issue_head();
    } else {
        // ===================== epilogue (warps 2..9): TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====================
        const int q = warp & 3;
        const int ch = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        uint32_t head_phase = 0;
        const uint32_t sb = smem_u32(s_bias);
        const uint32_t sa2 = smem_u32(s_a2);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t m0 = (tile / n_tiles_n) * G_TM;
            const int n0 = (int)((tile % n_tiles_n) * G_TN);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int64_t r = m0 + q * 32 + lane;             // this thread's output row
            float* dst = nullptr;
            int64_t cstride = 4;                              // floats between consecutive column quads
            if (out.mode == 0 || out.mode == 2) {
                if (r < M) dst = out.c + r * (int64_t)N + n0;   // row-major: a quad is 4 consecutive floats
            } else {
                // zin tiles [dir][m_tile][col/4][128 rows][4]: lanes (rows) are contiguous for a fixed column quad
                const int dir = n0 / out.n_per_dir;
                const int nloc = n0 - dir * out.n_per_dir;
                const int64_t m_tile = m0 / G_TM;
                dst = out.c + ((int64_t)dir * n_tiles_m + m_tile) * ((int64_t)out.n_per_dir * G_TM) + (int64_t)(nloc >> 2) * (G_TM * 4) +
                      (q * 32 + lane) * 4;
                cstride = G_TM * 4;
            }
            if (FUSE2) {
                // (1) relu(acc + bias1) -> fp16 (hi, lo) A tile of the second dense, K index = column of this layer
#pragma unroll
                for (int cbi = 0; cbi < 2; ++cbi) {
                    const int cb = ch * 2 + cbi;
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G_TN + cb * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int g8 = 0; g8 < 4; ++g8) {
                        uint32_t ph[4], pl[4];
#pragma unroll
                        for (int e = 0; e < 4; e += 2) {
                            const int j = g8 * 8 + e * 2;
                            const float4 bq = ld_shared_f4(sb + (uint32_t)(cb * 32 + j) * 4);
                            const float is = out.inv_in_scale;
                            const float x0 = fmaxf(fmaf(__uint_as_float(v[j]), is, bq.x), 0.f), x1 = fmaxf(fmaf(__uint_as_float(v[j + 1]), is, bq.y), 0.f);
                            const float x2 = fmaxf(fmaf(__uint_as_float(v[j + 2]), is, bq.z), 0.f), x3 = fmaxf(fmaf(__uint_as_float(v[j + 3]), is, bq.w), 0.f);
                            const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
                            ph[e] = *reinterpret_cast<const uint32_t*>(&h01); ph[e + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                            pl[e] = *reinterpret_cast<const uint32_t*>(&l01); pl[e + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                        }
                        const uint32_t off = ((cb >> 1) * G_A_BYTES) + sw128_offset(q * 32 + lane, (cb & 1) * 4 + g8);
                        st_shared_v4(sa2 + off, make_uint4(ph[0], ph[1], ph[2], ph[3]));
                        st_shared_v4(sa2 + 2 * G_A_BYTES + off, make_uint4(pl[0], pl[1], pl[2], pl[3]));
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(a2_ready); mbar_arrive(&tempty[acc]); }   // main accumulator is drained
                // (2) second accumulator -> + b2, relu -> out[r][32]: 16 columns per warp
                mbar_wait(d2_ready, head_phase);
                head_phase ^= 1;
                tc_fence_after();
                uint32_t v[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(tmem_base + ((uint32_t)(q * 32) << 16) + D2_COL + (uint32_t)(ch * 16))
                    : "memory");
                tmem_ld_wait();
                if (r < M) {
                    float4* o = reinterpret_cast<float4*>(out.c + r * 32 + ch * 16);
#pragma unroll
                    for (int o4 = 0; o4 < 4; ++o4) {
                        const float4 b = ld_shared_f4(sb + (uint32_t)(128 + ch * 16 + o4 * 4) * 4);
                        o[o4] = make_float4(fmaxf(__uint_as_float(v[4 * o4]) + b.x, 0.f), fmaxf(__uint_as_float(v[4 * o4 + 1]) + b.y, 0.f),
                                            fmaxf(__uint_as_float(v[4 * o4 + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[4 * o4 + 3]) + b.w, 0.f));
                    }
                }
                tc_fence_before();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
#pragma unroll 1
            for (int cbi = 0; cbi < G_TN / 64; ++cbi) {
                const int cb = ch * (G_TN / 64) + cbi;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G_TN + cb * 32), v);
                tmem_ld_wait();
                if (dst) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                               __uint_as_float(v[j + 3]));
                        if (out.bias) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(out.bias + n0 + cb * 32 + j));
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        if (out.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        *reinterpret_cast<float4*>(dst + (int64_t)(cb * 8 + (j >> 2)) * cstride) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ============================================================================================================
// CTA-PAIR projection GEMM (tcgen05 cta_group::2, cluster of 2): the kernel used for the Bi-LSTM input projections.
//   * one tcgen05.mma covers M = 256 rows (128 per CTA) x N = 256 columns and reads half of the B operand from each
//     CTA's shared memory, so each SM keeps only HALF of the 256-column weight tile -- all of K, hi and lo -- RESIDENT
//     (K = 192: 96 KB, K = 256: 128 KB).  A cluster owns one column tile for its whole life and walks over row pairs;
//     the clusters of the 2-4 column tiles walk in lockstep, so an A tile is fetched from HBM once and hits L2 after.
//   * only the A operand streams: 32 KB chunks ([128 rows][64] hi + lo) through a 3-6 stage TMA ring per CTA, both
//     CTAs' loads credited to the LEADER's "full" barrier; tcgen05.commit.multicast frees the slot in both CTAs.
//   * the single-CTA kernel above is load-latency-bound (2 x 96 KB stages, A and B streamed: tensor pipe 52 %);
//     here the bytes per MMA halve twice (B resident, A per CTA) and the ring is deep enough to cover HBM latency.
//   * epilogue per CTA: its 128 rows of the double-buffered accumulator -> (+bias) -> zin tiles, coalesced.
// ============================================================================================================
constexpr int GP_THREADS = 320;                       // warp 0: TMA producer, warp 1: MMA issue, warps 2..9: epilogue
constexpr int GP_A_STAGE = 2 * G_A_BYTES;             // 32 KB: A_hi + A_lo chunk

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GP_THREADS, 1)
gemm_f16x3_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                       const __half* __restrict__ b_hi, const __half* __restrict__ b_lo, GemmOut out, int64_t M, int N, int K,
                       int n_stages) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int n_chunks = K / G_KC;
    uint8_t* s_b = smem;                                              // [part][kc][128 rows][64]  resident
    uint8_t* s_a = smem + (size_t)2 * n_chunks * G_A_BYTES;           // [stage][hi|lo][128 rows][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_a + (size_t)n_stages * GP_A_STAGE);
    uint64_t* full = bars;                  // [n_stages] leader's copy is used: 1 arrive (leader producer) + 64 KB of tx
    uint64_t* empty = bars + 8;             // [n_stages] both CTAs: multicast commit
    uint64_t* tfull = bars + 16;            // [2] both CTAs: multicast commit
    uint64_t* tempty = bars + 18;           // [2] leader's copy is used: 16 arrivals (8 epilogue warps x 2 CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
    float* s_bias = reinterpret_cast<float*>(bars + 22);             // [256] bias of this cluster's column tile (1 KB)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_tiles_n = N / 256;
    const int64_t n_tiles_m = M / G_TM;
    const int64_t n_mpairs = (n_tiles_m + 1) / 2;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_tile = cluster_id % n_tiles_n;
    const int wi = cluster_id / n_tiles_n;                                 // worker index among the clusters of this column tile
    const int n_workers = (n_clusters - n_tile + n_tiles_n - 1) / n_tiles_n;
    const int n0 = n_tile * 256;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_a_hi); tma_prefetch_desc(&tm_a_lo);
        for (int s = 0; s < n_stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 16); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    if (threadIdx.x < 256) s_bias[threadIdx.x] = out.bias ? __ldg(out.bias + n0 + threadIdx.x) : 0.f;
    {   // resident B: this CTA's 128 rows of the column tile, all of K, hi and lo -> swizzled K-major [128][64] tiles
        const int per_row = K / 8;                                         // 16-byte chunks per row
        for (int i = threadIdx.x; i < 2 * 128 * per_row; i += GP_THREADS) {
            const int part = i / (128 * per_row), rem = i - part * (128 * per_row);
            const int row = rem / per_row, c = rem - row * per_row;
            const __half* src = (part ? b_lo : b_hi) + ((size_t)(n0 + rank * 128 + row)) * K;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + c);
            *reinterpret_cast<uint4*>(s_b + ((size_t)(part * n_chunks + (c >> 3))) * G_A_BYTES + sw128_offset(row, c & 7)) = v;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Cross-CTA signalling uses plain mbarrier arrives / waits: everything that is published lives in shared memory or TMEM and
    // is ordered by the async proxy (TMA complete_tx, tcgen05.commit) or tcgen05.fence -- cluster-scope acquire/release would
    // cost a MEMBAR.GPU + ERRBAR per arrive and an L1 invalidation (CCTL.IVALL) per wait (ncu: 20 % of the epilogue's samples).
    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own 128 rows of the A operand =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int64_t mp = wi; mp < n_mpairs; mp += n_workers) {
                const int m0 = (int)((mp * 2 + rank) * G_TM);             // rows beyond M are zero-filled by TMA
                for (int kc = 0; kc < n_chunks; ++kc) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* st = s_a + (size_t)stage * GP_A_STAGE;
                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * GP_A_STAGE);    // both CTAs' bytes
                    tma_load_2d_pair(st, &tm_a_hi, &full[stage], 0, kc * G_KC, m0);
                    tma_load_2d_pair(st + G_A_BYTES, &tm_a_lo, &full[stage], 0, kc * G_KC, m0);
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(256, 256);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            const uint32_t b_base = smem_u32(s_b);
            for (int64_t mp = wi; mp < n_mpairs; mp += n_workers) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);                  // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
                for (int kc = 0; kc < n_chunks; ++kc) {
                    mbar_wait(&full[stage], phase);                      // both CTAs' A chunks have landed
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t st = smem_u32(s_a + (size_t)stage * GP_A_STAGE);
#pragma unroll
                        for (int k = 0; k < G_KC / 16; ++k) {
                            const uint64_t a_hi = umma_desc_k_sw128(st + k * 32), a_lo = umma_desc_k_sw128(st + G_A_BYTES + k * 32);
                            const uint64_t bh = umma_desc_k_sw128(b_base + (uint32_t)((0 * n_chunks + kc) * G_A_BYTES) + k * 32);
                            const uint64_t bl = umma_desc_k_sw128(b_base + (uint32_t)((1 * n_chunks + kc) * G_A_BYTES) + k * 32);
                            umma_f16_ss_pair(d_tmem, a_lo, bh, idesc, (kc | k) != 0);
                            umma_f16_ss_pair(d_tmem, a_hi, bl, idesc, 1);
                            umma_f16_ss_pair(d_tmem, a_hi, bh, idesc, 1);
                        }
                        umma_commit_pair(&empty[stage]);                  // frees the ring slot in both CTAs
                        if (kc == n_chunks - 1) umma_commit_pair(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs): own 128 rows -> zin tile =====================
        // TMEM lane quarter = warp % 4; column half (4 blocks of 32 columns) = (warp - 2) / 4
        const int q = warp & 3;
        const int ch = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        const int dir = n0 / out.n_per_dir;
        const int nloc = n0 - dir * out.n_per_dir;
        const uint32_t sb = smem_u32(s_bias);
        for (int64_t mp = wi; mp < n_mpairs; mp += n_workers) {
            const int64_t m_tile = mp * 2 + rank;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            float* dst = nullptr;
            if (m_tile < n_tiles_m)
                dst = out.c + ((int64_t)dir * n_tiles_m + m_tile) * ((int64_t)out.n_per_dir * G_TM) + (int64_t)(nloc >> 2) * (G_TM * 4) +
                      (q * 32 + lane) * 4;
#pragma unroll 2
            for (int cbi = 0; cbi < 4; ++cbi) {
                const int cb = ch * 4 + cbi;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + cb * 32), v);
                tmem_ld_wait();
                if (dst) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bq = ld_shared_f4(sb + (uint32_t)(cb * 32 + j) * 4);
                        const float4 o = make_float4(__uint_as_float(v[j]) + bq.x, __uint_as_float(v[j + 1]) + bq.y,
                                                     __uint_as_float(v[j + 2]) + bq.z, __uint_as_float(v[j + 3]) + bq.w);
                        *reinterpret_cast<float4*>(dst + (int64_t)(cb * 8 + (j >> 2)) * (G_TM * 4)) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&tempty[acc], 0);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

// ---- fp32 -> (hi, lo) fp16 split ------------------------------------------------------------------------
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __half h, l;
    split_f16(x[i], h, l);
    hi[i] = h; lo[i] = l;
}

int launch_split_f16(const float* x, __half* hi, __half* lo, int64_t n, cudaStream_t st) {
    if (n <= 0) return 0;
    split_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, hi, lo, n);
    return 1;
}

// ---- host side: tensor maps + launch --------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// fp16 row-major [rows][K] tensor, box = [box_rows][64] (128-byte rows), 128-byte swizzle
bool make_tmap_f16_k64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)G_KC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// fp16 operand of 32 columns out of rows of K halves: box = 64 bytes x box_rows, 64-byte swizzle (nrv_rec_tc.cu, read_rnn11's x tiles)
bool make_tmap_f16_k32_sw64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// fp16 operand tail of 16 columns (K = 16: one UMMA K-step): box = 32 bytes x box_rows, 32-byte swizzle (nrv_cnn.cu: columns
// 384..399 of the Dense(400 -> 64) weights)
bool make_tmap_f16_k16_sw32(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {16u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 8-bit operand (e4m3 copies of activations, nrv_fused_pair.cu F8): rows of K bytes, box = 128 bytes x box_rows, 128-byte swizzle --
// the same 16 KB tile geometry as a K = 64 fp16 tile, so the UMMA descriptors and k-step offsets (32 B) are shared
bool make_tmap_u8_k128(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int TN, bool FUSE2>
static int launch_gemm_tn(const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo, int64_t M, int N, int K,
                          const GemmOut& o, int num_sms, cudaStream_t st) {
    using Cfg = GemmCfg<TN, FUSE2>;
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (!make_tmap_f16_k64(&ta_hi, a_hi, M, K, G_TM) || !make_tmap_f16_k64(&ta_lo, a_lo, M, K, G_TM) ||
        !make_tmap_f16_k64(&tb_hi, b_hi, N, K, TN) || !make_tmap_f16_k64(&tb_lo, b_lo, N, K, TN))
        return -2;
    auto kern = gemm_f16x3_kernel<TN, FUSE2>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    const int64_t n_tiles = ((M + G_TM - 1) / G_TM) * (N / TN);
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, num_sms > 0 ? num_sms : 148);
    kern<<<grid, G_THREADS, Cfg::SMEM, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, o, M, N, K);
    return 1;
}

int launch_gemm_f16x3(const __half* a_hi, const __half* a_lo, const __half* b_hi, const __half* b_lo, int64_t M, int N, int K,
                      float* c, const float* bias, int mode, int T, int64_t nw, int n_per_dir, int relu, int num_sms,
                      cudaStream_t st, const __half* w2t_hi, const __half* w2t_lo, const float* b2, float in_scale) {
    if (M <= 0) return 0;
    if (K % G_KC != 0 || K <= 0 || N <= 0) return -1;
    GemmOut o{c, bias, mode, T, nw, n_per_dir, relu, w2t_hi, w2t_lo, b2, 1.f / in_scale};
    // CTA-pair kernel (B resident, 8 epilogue warps) unless NRV_GEMM=single: proj2 36.7 -> 30.8 ms, proj3 21.3 -> 17.9 ms per step
    const bool use_pair = !(getenv("NRV_GEMM") && !strcmp(getenv("NRV_GEMM"), "single"));
    if (mode == 1 && use_pair && N % 256 == 0 && n_per_dir % 256 == 0 && M % G_TM == 0 && K <= 256) {
        CUtensorMap ta_hi, ta_lo;
        if (!make_tmap_f16_k64(&ta_hi, a_hi, M, K, G_TM) || !make_tmap_f16_k64(&ta_lo, a_lo, M, K, G_TM)) return -2;
        const int n_chunks = K / G_KC;
        const size_t b_bytes = (size_t)2 * n_chunks * G_A_BYTES;
        int n_stages = (int)((232448 - 2304 - b_bytes) / GP_A_STAGE);
        if (n_stages > 6) n_stages = 6;
        if (n_stages < 2) return -1;
        const size_t smem = b_bytes + (size_t)n_stages * GP_A_STAGE + 1024 + 256 + 1024;
        cudaFuncSetAttribute(gemm_f16x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int sms = num_sms > 0 ? num_sms : 148;
        const int n_tiles_n = N / 256;
        const int64_t n_mpairs = (M / G_TM + 1) / 2;
        int64_t n_clusters = std::min<int64_t>(sms / 2, n_mpairs * n_tiles_n);
        if (n_clusters < n_tiles_n) n_clusters = n_tiles_n;                // every column tile needs a worker
        gemm_f16x3_pair_kernel<<<(unsigned)(2 * n_clusters), GP_THREADS, smem, st>>>(ta_hi, ta_lo, b_hi, b_lo, o, M, N, K, n_stages);
        return 1;
    }
    if (mode == 2) {
        if (N != 128 || !w2t_hi || !w2t_lo || !b2 || !bias) return -1;
        return launch_gemm_tn<128, true>(a_hi, a_lo, b_hi, b_lo, M, N, K, o, num_sms, st);
    }
    if (mode == 1 && (M % G_TM != 0)) return -1;
    if (N % 256 == 0) {
        if (mode == 1 && (n_per_dir % 256 != 0)) return -1;
        return launch_gemm_tn<256, false>(a_hi, a_lo, b_hi, b_lo, M, N, K, o, num_sms, st);
    }
    if (N % 128 == 0) {
        if (mode == 1 && (n_per_dir % 128 != 0)) return -1;
        return launch_gemm_tn<128, false>(a_hi, a_lo, b_hi, b_lo, M, N, K, o, num_sms, st);
    }
    return -1;
}

}  // namespace nrv
