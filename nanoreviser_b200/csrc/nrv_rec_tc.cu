// K3 recurrence on the 5th-generation tensor cores (tcgen05 + TMEM), the "persistent recurrence kernel that
// keeps the recurrent weights resident in shared memory" of BASELINE.json's north_star.
//
// The input projection x_t.Wk (+bias, with the preceding BatchNormalization folded in) has already been
// computed for all timesteps by the time-batched GEMM (nrv_gemm.cu) into zin[dir][t][w][4u].  What is left
// per step is  z = zin[t] + h_{t-1} . Wr ;  gates ;  c, h  -- sequential in t, independent across windows.
//
// u = 64 variant (read_rnn11, total_rnn2; lstmmodel.py:46,51).  One CTA = one direction x TWO tiles of 128
// windows that ping-pong:
//   * Wr^T [256 gate columns][64] as an fp16 (hi, lo) pair stays in shared memory for the whole kernel
//     (64 KB, K-major, 128-byte swizzle -- the B operand of every MMA);
//   * h_{t-1} of each tile lives in shared memory as an fp16 (hi, lo) pair in the same swizzled K-major
//     layout (the A operand), written by that tile's epilogue warps;
//   * the accumulator of each tile is a 128-lane x 256-column fp32 block of TMEM;
//   * warp 0 issues, per tile and step, 12 tcgen05.mma (4 K-steps x {lo*hi, hi*lo, hi*hi}) from one lane and
//     commits to an mbarrier; warps 1-4 / 5-8 are the epilogue of tile A / B: tcgen05.ld -> + zin ->
//     hard_sigmoid / tanh -> c (registers, never leaves the thread) -> h -> shared memory + global.
// While the epilogue of tile A runs on the CUDA cores, the tensor core works on tile B, and vice versa.
#include "nrv_common.cuh"
#include "nrv_tc.cuh"

namespace nrv {

using namespace tc;

constexpr int RT_THREADS = 288;                 // 1 MMA warp + 2 x 4 epilogue warps
constexpr int RT_W_BYTES = 256 * 64 * 2;        // 32 KB: Wr^T hi (or lo)
constexpr int RT_H_BYTES = 128 * 64 * 2;        // 16 KB: h tile hi (or lo)
constexpr size_t RT_SMEM = 2 * RT_W_BYTES + 4 * RT_H_BYTES + 1024 + 128;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tanh(|x|) = 1 - 2 / (exp(2|x|) + 1), sign restored afterwards: with e >= 1 the reciprocal is <= 0.5, which
// bounds the absolute error of MUFU.EX2 + MUFU.RCP to ~2e-7 everywhere; saturates cleanly at +-1.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = ex2_approx(fabsf(x) * 2.8853900817779268f);
    const float t = fmaf(-2.f, rcp_approx(e + 1.f), 1.f);
    return copysignf(t, x);
}
__device__ __forceinline__ float hsig(float x) { return __saturatef(fmaf(0.2f, x, 0.5f)); }

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <int OMODE>
__global__ void __launch_bounds__(RT_THREADS, 1)
lstm_rec_tc64_kernel(const __half* __restrict__ wr_hi, const __half* __restrict__ wr_lo,
                     const float* __restrict__ zin, float* __restrict__ act_out, __half* __restrict__ out_hi,
                     __half* __restrict__ out_lo, int out_ld, int64_t nw, int T) {
    constexpr int U = 64, N = 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_whi = smem;
    uint8_t* s_wlo = smem + RT_W_BYTES;
    uint8_t* s_h = smem + 2 * RT_W_BYTES;               // [tile][hi|lo][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * RT_W_BYTES + 4 * RT_H_BYTES);
    uint64_t* h_ready = bars;                           // [2] count 4 (one arrive per epilogue warp)
    uint64_t* acc_ready = bars + 2;                     // [2] count 1 (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int64_t w0[2] = {(int64_t)blockIdx.x * 256, (int64_t)blockIdx.x * 256 + 128};
    const bool valid[2] = {w0[0] < nw, w0[1] < nw};

    if (threadIdx.x == 0) {
        mbar_init(&h_ready[0], 4); mbar_init(&h_ready[1], 4);
        mbar_init(&acc_ready[0], 1); mbar_init(&acc_ready[1], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    // resident recurrent weights: [256 rows][64 halves] -> K-major 128-byte-swizzled tiles
    {
        const uint4* gh = reinterpret_cast<const uint4*>(wr_hi + (size_t)dir * N * U);
        const uint4* gl = reinterpret_cast<const uint4*>(wr_lo + (size_t)dir * N * U);
        for (int i = threadIdx.x; i < N * 8; i += RT_THREADS) {
            const int row = i >> 3, c = i & 7;
            const uint32_t off = sw128_offset(row, c);
            *reinterpret_cast<uint4*>(s_whi + off) = __ldg(gh + i);
            *reinterpret_cast<uint4*>(s_wlo + off) = __ldg(gl + i);
        }
    }
    fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16_f32(128, N);
        const uint64_t b_hi = umma_desc_k_sw128(smem_u32(s_whi)), b_lo = umma_desc_k_sw128(smem_u32(s_wlo));
        for (int s = 1; s < T; ++s) {
            for (int X = 0; X < 2; ++X) {
                if (!valid[X]) continue;
                mbar_wait(&h_ready[X], (uint32_t)((s - 1) & 1));      // h_{s-1} of tile X is in shared memory
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t a_hi = umma_desc_k_sw128(smem_u32(s_h + (X * 2 + 0) * RT_H_BYTES));
                    const uint64_t a_lo = umma_desc_k_sw128(smem_u32(s_h + (X * 2 + 1) * RT_H_BYTES));
                    const uint32_t d = tmem_base + (uint32_t)(X * N);
#pragma unroll
                    for (int k = 0; k < U / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);            // +32 B along K (16-byte units)
                        umma_f16_ss(d, a_lo + adv, b_hi + adv, idesc, k != 0);
                        umma_f16_ss(d, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_f16_ss(d, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    umma_commit(&acc_ready[X]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: warps 1-4 -> tile 0, warps 5-8 -> tile 1 =====================
        const int X = (warp - 1) >> 2;
        const int q = warp & 3;                          // TMEM lane quarter accessible to this warp
        const int row = q * 32 + lane;
        const int64_t w = w0[X] + row;
        const bool live = w < nw;
        if (valid[X]) {
            uint8_t* hs_hi = s_h + (X * 2 + 0) * RT_H_BYTES;
            uint8_t* hs_lo = s_h + (X * 2 + 1) * RT_H_BYTES;
            float c[U];
#pragma unroll
            for (int j = 0; j < U; ++j) c[j] = 0.f;
            for (int s = 0; s < T; ++s) {
                const int t = dir ? (T - 1 - s) : s;
                const float* zrow = zin + (((int64_t)dir * T + t) * nw + (live ? w : 0)) * N;
                if (s > 0) {
                    mbar_wait(&acc_ready[X], (uint32_t)((s - 1) & 1));
                    tc_fence_after();
                }
#pragma unroll
                for (int cb = 0; cb < N / 32; ++cb) {
                    uint32_t v[32];
                    if (s > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(X * N + cb * 32), v);
                    float4 z[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        z[j] = live ? __ldg(reinterpret_cast<const float4*>(zrow + cb * 32) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s > 0) {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            z[j].x += __uint_as_float(v[4 * j + 0]); z[j].y += __uint_as_float(v[4 * j + 1]);
                            z[j].z += __uint_as_float(v[4 * j + 2]); z[j].w += __uint_as_float(v[4 * j + 3]);
                        }
                    }
                    float h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {                     // Keras 2.2.4 LSTM cell, gates i,f,c,o
                        const float ig = hsig(z[j].x), fg = hsig(z[j].y), gg = tanh_fast(z[j].z), og = hsig(z[j].w);
                        const float cn = fmaf(fg, c[cb * 8 + j], ig * gg);
                        c[cb * 8 + j] = cn;
                        h[j] = og * tanh_fast(cn);
                    }
                    __half hh[8], hl[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) split_f16(h[j], hh[j], hl[j]);
                    const uint4 phi = make_uint4(pack_half2(hh[0], hh[1]), pack_half2(hh[2], hh[3]), pack_half2(hh[4], hh[5]),
                                                 pack_half2(hh[6], hh[7]));
                    const uint4 plo = make_uint4(pack_half2(hl[0], hl[1]), pack_half2(hl[2], hl[3]), pack_half2(hl[4], hl[5]),
                                                 pack_half2(hl[6], hl[7]));
                    if (s + 1 < T) {                                   // A operand of the next step
                        const uint32_t off = sw128_offset(row, cb);
                        *reinterpret_cast<uint4*>(hs_hi + off) = phi;
                        *reinterpret_cast<uint4*>(hs_lo + off) = plo;
                    }
                    if (live) {
                        if (OMODE == 0) {
                            float* o = act_out + (w * T + t) * (2 * U) + dir * U + cb * 8;
                            *reinterpret_cast<float4*>(o) = make_float4(h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<float4*>(o + 4) = make_float4(h[4], h[5], h[6], h[7]);
                        } else {
                            const int64_t off = (w * T + t) * out_ld + dir * U + cb * 8;
                            *reinterpret_cast<uint4*>(out_hi + off) = phi;
                            *reinterpret_cast<uint4*>(out_lo + off) = plo;
                        }
                    }
                }
                if (s + 1 < T) {
                    tc_fence_before();           // our tcgen05.ld of this step precede the next MMA's writes
                    fence_proxy_async_smem();    // our h writes are visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_ready[X]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

int launch_lstm_rec_tc64(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st) {
    if (n_win <= 0) return 0;
    if (L.u != 64 || !L.rt_hi) return -1;
    dim3 grid((unsigned)((n_win + 255) / 256), 2);
    if (io.out_hi) {
        auto kern = lstm_rec_tc64_kernel<1>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM);
        kern<<<grid, RT_THREADS, RT_SMEM, st>>>(L.rt_hi, L.rt_lo, io.zin, nullptr, io.out_hi, io.out_lo, io.out_ld, n_win, T);
    } else {
        auto kern = lstm_rec_tc64_kernel<0>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RT_SMEM);
        kern<<<grid, RT_THREADS, RT_SMEM, st>>>(L.rt_hi, L.rt_lo, io.zin, io.act_out, nullptr, nullptr, 0, n_win, T);
    }
    return 1;
}

// ============================================================================================================
// u = 128 variant (total_rnn1, lstmmodel.py:49): N = 512 gate columns, K = 128.
// Wr^T as an fp16 (hi, lo) pair is 2 x 128 KB -- more than one SM's shared memory -- so the hi half stays
// resident (128 KB, used by the lo*hi and hi*hi passes) and the lo half (used only by the hi*lo pass) is
// streamed from L2 every step through a 2-stage TMA ring of [128 columns][64 K] boxes (16 KB), issued by a
// dedicated producer warp that never waits on the recurrence (weights do not depend on h).
// One CTA = one direction x one tile of 128 windows; the accumulator fills all 512 TMEM columns, split in two
// 256-column halves (units 0-63 / 64-127) with their own "accumulator ready" barriers, so that the epilogue of
// half 0 overlaps the MMAs of half 1.  Eight epilogue warps: two per TMEM lane quarter, one per column half.
// ============================================================================================================
constexpr int R2_THREADS = 320;                       // MMA warp, TMA warp, 8 epilogue warps
constexpr int R2_WHI_BYTES = 2 * 512 * 64 * 2;        // 128 KB: two K-chunks of [512 rows][64]
constexpr int R2_H_BYTES = 128 * 64 * 2;              // 16 KB: one K-chunk of h (hi or lo)
constexpr int R2_BOX_BYTES = 128 * 64 * 2;            // 16 KB: streamed W_lo box
constexpr int R2_STAGES = 2;
constexpr size_t R2_SMEM = (size_t)R2_WHI_BYTES + 4 * R2_H_BYTES + R2_STAGES * R2_BOX_BYTES + 1024 + 128;

__global__ void __launch_bounds__(R2_THREADS, 1)
lstm_rec_tc128_kernel(const __half* __restrict__ wr_hi, const __grid_constant__ CUtensorMap tm_wlo,
                      const float* __restrict__ zin, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                      int out_ld, int64_t nw, int T) {
    constexpr int U = 128, N = 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_whi = smem;                                   // [kc][512 rows][64]
    uint8_t* s_h = smem + R2_WHI_BYTES;                      // [hi|lo][kc][128 rows][64]
    uint8_t* s_ring = s_h + 4 * R2_H_BYTES;                  // [stage][128 rows][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + R2_STAGES * R2_BOX_BYTES);
    uint64_t* h_ready = bars;                                // count 8
    uint64_t* acc_ready = bars + 1;                          // [2] count 1
    uint64_t* full = bars + 3;                               // [R2_STAGES]
    uint64_t* empty = bars + 3 + R2_STAGES;                  // [R2_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 + 2 * R2_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int64_t w0 = (int64_t)blockIdx.x * 128;

    if (threadIdx.x == 0) {
        mbar_init(h_ready, 8);
        mbar_init(&acc_ready[0], 1); mbar_init(&acc_ready[1], 1);
        for (int s = 0; s < R2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&tm_wlo);
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    {
        const uint4* gh = reinterpret_cast<const uint4*>(wr_hi + (size_t)dir * N * U);   // [512][128] halves: 16 chunks per row
        for (int i = threadIdx.x; i < N * 16; i += R2_THREADS) {
            const int row = i >> 4, c16 = i & 15;
            const int kc = c16 >> 3, c = c16 & 7;
            *reinterpret_cast<uint4*>(s_whi + kc * (512 * 128) + sw128_offset(row, c)) = __ldg(gh + i);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        // ===================== TMA producer: W_lo boxes, (step, half, n-quarter-in-half, k-chunk) order ==========
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int s = 1; s < T; ++s)
                for (int p = 0; p < 8; ++p) {
                    const int nq = p >> 1, kc = p & 1;           // rows nq*128.., K chunk kc
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full[stage], R2_BOX_BYTES);
                    tma_load_2d(s_ring + stage * R2_BOX_BYTES, &tm_wlo, &full[stage], kc * 64, dir * N + nq * 128);
                    if (++stage == R2_STAGES) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 0) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc256 = umma_idesc_f16_f32(128, 256), idesc128 = umma_idesc_f16_f32(128, 128);
        int stage = 0; uint32_t phase = 0;
        const uint32_t a_base = smem_u32(s_h), w_base = smem_u32(s_whi);
        for (int s = 1; s < T; ++s) {
            mbar_wait(h_ready, (uint32_t)((s - 1) & 1));
            tc_fence_after();
            for (int hf = 0; hf < 2; ++hf) {
                const uint32_t d = tmem_base + (uint32_t)(hf * 256);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {                 // resident W_hi: lo*hi and hi*hi
                        const int kc = k >> 2, kk = k & 3;
                        const uint64_t a_hi = umma_desc_k_sw128(a_base + (0 * 2 + kc) * R2_H_BYTES + kk * 32);
                        const uint64_t a_lo = umma_desc_k_sw128(a_base + (1 * 2 + kc) * R2_H_BYTES + kk * 32);
                        const uint64_t b_hi = umma_desc_k_sw128(w_base + kc * (512 * 128) + hf * (256 * 128) + kk * 32);
                        umma_f16_ss(d, a_lo, b_hi, idesc256, k != 0);
                        umma_f16_ss(d, a_hi, b_hi, idesc256, 1);
                    }
                }
                __syncwarp();
                for (int p = 0; p < 4; ++p) {                     // streamed W_lo: hi*lo, four [128 x 64] boxes per half
                    const int nq = p >> 1, kc = p & 1;
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t bst = smem_u32(s_ring + stage * R2_BOX_BYTES);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint64_t a_hi = umma_desc_k_sw128(a_base + (0 * 2 + kc) * R2_H_BYTES + kk * 32);
                            umma_f16_ss(d + (uint32_t)(nq * 128), a_hi, umma_desc_k_sw128(bst + kk * 32), idesc128, 1);
                        }
                        umma_commit(&empty[stage]);
                        if (p == 3) umma_commit(&acc_ready[hf]);
                    }
                    __syncwarp();
                    if (++stage == R2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== epilogue: warps 2..9; lane quarter = warp % 4, column half = (warp - 2) / 4 ===========
        const int q = warp & 3;
        const int hf = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int64_t w = w0 + row;
        const bool live = w < nw;
        float c[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) c[j] = 0.f;
        for (int s = 0; s < T; ++s) {
            const int t = dir ? (T - 1 - s) : s;
            const float* zrow = zin + (((int64_t)dir * T + t) * nw + (live ? w : 0)) * N + hf * 256;
            if (s > 0) {
                mbar_wait(&acc_ready[hf], (uint32_t)((s - 1) & 1));
                tc_fence_after();
            }
#pragma unroll
            for (int cb = 0; cb < 8; ++cb) {
                uint32_t v[32];
                if (s > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 256 + cb * 32), v);
                float4 z[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    z[j] = live ? __ldg(reinterpret_cast<const float4*>(zrow + cb * 32) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (s > 0) {
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        z[j].x += __uint_as_float(v[4 * j + 0]); z[j].y += __uint_as_float(v[4 * j + 1]);
                        z[j].z += __uint_as_float(v[4 * j + 2]); z[j].w += __uint_as_float(v[4 * j + 3]);
                    }
                }
                float h[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float ig = hsig(z[j].x), fg = hsig(z[j].y), gg = tanh_fast(z[j].z), og = hsig(z[j].w);
                    const float cn = fmaf(fg, c[cb * 8 + j], ig * gg);
                    c[cb * 8 + j] = cn;
                    h[j] = og * tanh_fast(cn);
                }
                __half hh[8], hl[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) split_f16(h[j], hh[j], hl[j]);
                const uint4 phi = make_uint4(pack_half2(hh[0], hh[1]), pack_half2(hh[2], hh[3]), pack_half2(hh[4], hh[5]),
                                             pack_half2(hh[6], hh[7]));
                const uint4 plo = make_uint4(pack_half2(hl[0], hl[1]), pack_half2(hl[2], hl[3]), pack_half2(hl[4], hl[5]),
                                             pack_half2(hl[6], hl[7]));
                if (s + 1 < T && hf == 1) {
                    // units 64 + cb*8 .. +8  ->  K chunk 1 of the h tile, 16-byte chunk cb.  Safe to overwrite now:
                    // acc_ready[1] fires after ALL MMAs of this step, i.e. nothing reads h_{s-1} any more.
                    const uint32_t off = sw128_offset(row, cb);
                    *reinterpret_cast<uint4*>(s_h + (0 * 2 + 1) * R2_H_BYTES + off) = phi;
                    *reinterpret_cast<uint4*>(s_h + (1 * 2 + 1) * R2_H_BYTES + off) = plo;
                }
                if (live) {
                    const int64_t off = (w * T + t) * out_ld + dir * U + hf * 64 + cb * 8;
                    *reinterpret_cast<uint4*>(out_hi + off) = phi;
                    *reinterpret_cast<uint4*>(out_lo + off) = plo;
                }
            }
            if (s + 1 < T && hf == 0) {
                // Half 0 ran while the MMAs of half 1 were still reading h_{s-1}: publish its part of h_s only after
                // they have retired, from the copy it has just written to global memory (own writes, program order).
                if (s > 0) mbar_wait(&acc_ready[1], (uint32_t)((s - 1) & 1));
                if (live) {
                    const int64_t goff = (w * T + t) * out_ld + dir * U;
#pragma unroll
                    for (int cb = 0; cb < 8; ++cb) {
                        const uint32_t off = sw128_offset(row, cb);
                        *reinterpret_cast<uint4*>(s_h + (0 * 2 + 0) * R2_H_BYTES + off) = *reinterpret_cast<const uint4*>(out_hi + goff + cb * 8);
                        *reinterpret_cast<uint4*>(s_h + (1 * 2 + 0) * R2_H_BYTES + off) = *reinterpret_cast<const uint4*>(out_lo + goff + cb * 8);
                    }
                }
            }
            if (s + 1 < T) {
                tc_fence_before();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(h_ready);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

bool make_tmap_f16_k64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);   // nrv_gemm.cu

int launch_lstm_rec_tc128(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st) {
    if (n_win <= 0) return 0;
    if (L.u != 128 || !L.rt_hi || !io.out_hi) return -1;
    CUtensorMap tm;
    if (!make_tmap_f16_k64(&tm, L.rt_lo, 2 * 512, 128, 128)) return -2;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(lstm_rec_tc128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R2_SMEM);
        attr = true;
    }
    dim3 grid((unsigned)((n_win + 127) / 128), 2);
    lstm_rec_tc128_kernel<<<grid, R2_THREADS, R2_SMEM, st>>>(L.rt_hi, tm, io.zin, io.out_hi, io.out_lo, io.out_ld, n_win, T);
    return 1;
}

}  // namespace nrv
