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
// tanh(x) = 1 - 2 / (exp(2x) + 1); absolute error ~2e-7 (MUFU.EX2 + MUFU.RCP), saturates cleanly at +-1
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = ex2_approx(x * 2.8853900817779268f);
    return fmaf(-2.f, rcp_approx(e + 1.f), 1.f);
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

}  // namespace nrv
