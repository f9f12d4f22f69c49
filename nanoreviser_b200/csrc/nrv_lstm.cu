// K3 (fp32 SIMT variant): one Keras-2.2.4 Bidirectional(LSTM(u, return_sequences=True)) layer
// (lstmmodel.py:44,46,49,51) over a chunk of windows, fused with the BatchNormalization that
// follows it (lstmmodel.py:45,47,50).
//
// Per CTA: a tile of TM windows, one direction (blockIdx.y).  For every timestep
//     z[TM][4u] = [x_t | h_{t-1}] . [Wk ; Wr] + b                 (one GEMM, K = in + u)
//     i,f,o = hard_sigmoid(z) ; g = tanh(z_c) ; c = f*c + i*g ; h = o*tanh(c)
// with gate columns interleaved (col = unit*4 + gate) so that a thread's 8x8 register tile holds all
// four gates of its two units and the cell state never leaves registers.  The x_t|h tile lives in
// shared memory k-major; weights stream from L2 through a double-buffered cp.async ring.
// Windows overlap in the read, but every window restarts from a zero state, so nothing is shared
// between windows except the per-base inputs (features / CNN output), which are indexed, not copied.
#include <cuda_fp16.h>

#include "nrv_common.cuh"

namespace nrv {

template <int IN_A, int IN_B, int U, int TM>
struct LstmTile {
    static constexpr int IN = IN_A + IN_B;
    static constexpr int K = IN + U;
    static constexpr int KC = 16;
    static constexpr int KP = (K + KC - 1) / KC * KC;
    static constexpr int NCH = KP / KC;
    static constexpr int N = 4 * U;
    static constexpr int WARPS_N = N / 64;
    static constexpr int WARPS_M = TM / 32;
    static constexpr int NT = 32 * WARPS_M * WARPS_N;
    static constexpr int A_LD = TM + 4;
    static constexpr size_t SMEM = (size_t)(KP * A_LD + 2 * KC * N) * sizeof(float);
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float hard_sigmoid(float x) { return __saturatef(fmaf(0.2f, x, 0.5f)); }

// ZMODE 0: the input projection is part of the per-step GEMM (K = in + u), accumulators start from the bias.
// ZMODE 1: the input projection was computed by the tensor-core GEMM (nrv_gemm.cu): accumulators start from
//          zin[dir][t][w][4u] (+ zsig[base][2][4u], the per-base part of total_rnn1's input), K = u only.
// OMODE 0: fp32 output with the following BatchNormalization folded in.  OMODE 1: raw h as an fp16 (hi, lo)
//          pair -- the A operand of the next layer's tensor-core projection (its BN is folded into that GEMM).
template <int IN_A, int IN_B, int U, int TM, int ZMODE, int OMODE>
__global__ void __launch_bounds__((LstmTile<IN_A, IN_B, U, TM>::NT), 1)
lstm_layer_kernel(const float* __restrict__ act_in, const float* __restrict__ base_in,
                  const int32_t* __restrict__ win_base, const float* __restrict__ wcat0,
                  const float* __restrict__ wcat1, const float* __restrict__ bias0,
                  const float* __restrict__ bias1, const float* __restrict__ bn_scale,
                  const float* __restrict__ bn_shift, float* __restrict__ act_out,
                  const float* __restrict__ zin, const float* __restrict__ zsig,
                  __half* __restrict__ out_hi, __half* __restrict__ out_lo, int out_ld, int64_t out_nwp,
                  int64_t n_win, int T) {
    using C = LstmTile<IN_A, IN_B, U, TM>;
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                       // [KP][A_LD]  rows [0,IN) = x_t, [IN,K) = h_{t-1}
    float* Bs = smem + C::KP * C::A_LD;     // [2][KC][N]

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp / C::WARPS_N, wn = warp % C::WARPS_N;
    const int ly = lane >> 3, lx = lane & 7;
    const int dir = blockIdx.y;
    const int64_t w0 = (int64_t)blockIdx.x * TM;
    const float* __restrict__ wcat = dir ? wcat1 : wcat0;
    const float* __restrict__ bias = dir ? bias1 : bias0;

    const int m0 = wm * 32 + ly * 8;                 // first of this thread's 8 rows
    const int uA = wn * 16 + lx, uB = uA + 8;        // this thread's two units
    const int colA = wn * 64 + lx * 4, colB = colA + 32;

    for (int i = tid; i < C::KP * C::A_LD; i += C::NT) As[i] = 0.f;
    float c_state[8][2];
#pragma unroll
    for (int r = 0; r < 8; ++r) c_state[r][0] = c_state[r][1] = 0.f;
    const float4 bA = *reinterpret_cast<const float4*>(bias + uA * 4);
    const float4 bB = *reinterpret_cast<const float4*>(bias + uB * 4);
    float bnsA = 1.f, bnsB = 1.f, bntA = 0.f, bntB = 0.f;
    if (bn_scale) {
        bnsA = bn_scale[dir * U + uA]; bnsB = bn_scale[dir * U + uB];
        bntA = bn_shift[dir * U + uA]; bntB = bn_shift[dir * U + uB];
    }
    __syncthreads();

    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        // ---- x_t -> As rows [0, IN) -------------------------------------------------------
        if (IN_A > 0) {
            constexpr int Q = IN_A / 4;
            for (int i = tid; i < TM * Q; i += C::NT) {
                const int m = i % TM, q = i / TM;
                const int64_t w = w0 + m;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w < n_win) v = __ldg(reinterpret_cast<const float4*>(act_in + (w * T + t) * IN_A) + q);
                As[(q * 4 + 0) * C::A_LD + m] = v.x;
                As[(q * 4 + 1) * C::A_LD + m] = v.y;
                As[(q * 4 + 2) * C::A_LD + m] = v.z;
                As[(q * 4 + 3) * C::A_LD + m] = v.w;
            }
        }
        if (IN_B > 0) {
            constexpr int Q = IN_B / 2;
            for (int i = tid; i < TM * Q; i += C::NT) {
                const int m = i % TM, q = i / TM;
                const int64_t w = w0 + m;
                float2 v = make_float2(0.f, 0.f);
                if (w < n_win) {
                    const int64_t b = (int64_t)win_base[w] + t;
                    v = __ldg(reinterpret_cast<const float2*>(base_in + b * IN_B) + q);
                }
                As[(IN_A + q * 2 + 0) * C::A_LD + m] = v.x;
                As[(IN_A + q * 2 + 1) * C::A_LD + m] = v.y;
            }
        }
        // ---- GEMM over K in chunks of KC, weights double-buffered ----------------------------
        float acc[8][8];
        if (ZMODE == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                acc[r][0] = bA.x; acc[r][1] = bA.y; acc[r][2] = bA.z; acc[r][3] = bA.w;
                acc[r][4] = bB.x; acc[r][5] = bB.y; acc[r][6] = bB.z; acc[r][7] = bB.w;
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int64_t w = w0 + m0 + r;
                float4 zA = make_float4(0.f, 0.f, 0.f, 0.f), zB = zA;
                if (w < n_win) {
                    const float* zp = zin + (((int64_t)dir * T + t) * n_win + w) * C::N;
                    zA = __ldg(reinterpret_cast<const float4*>(zp + colA));
                    zB = __ldg(reinterpret_cast<const float4*>(zp + colB));
                    if (zsig) {
                        const float* sp = zsig + ((int64_t)win_base[w] + t) * (2 * C::N) + dir * C::N;
                        const float4 sA = __ldg(reinterpret_cast<const float4*>(sp + colA));
                        const float4 sB = __ldg(reinterpret_cast<const float4*>(sp + colB));
                        zA.x += sA.x; zA.y += sA.y; zA.z += sA.z; zA.w += sA.w;
                        zB.x += sB.x; zB.y += sB.y; zB.z += sB.z; zB.w += sB.w;
                    }
                }
                acc[r][0] = zA.x; acc[r][1] = zA.y; acc[r][2] = zA.z; acc[r][3] = zA.w;
                acc[r][4] = zB.x; acc[r][5] = zB.y; acc[r][6] = zB.z; acc[r][7] = zB.w;
            }
        }
        constexpr int CHUNK_V4 = C::KC * C::N / 4;
        for (int i = tid; i < CHUNK_V4; i += C::NT) cp_async16(Bs + i * 4, wcat + i * 4);
        cp_async_commit();
        for (int kc = 0; kc < C::NCH; ++kc) {
            if (kc + 1 < C::NCH) {
                float* dst = Bs + ((kc + 1) & 1) * C::KC * C::N;
                const float* src = wcat + (size_t)(kc + 1) * C::KC * C::N;
                for (int i = tid; i < CHUNK_V4; i += C::NT) cp_async16(dst + i * 4, src + i * 4);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float* Bc = Bs + (kc & 1) * C::KC * C::N;
            const float* Ac = As + kc * C::KC * C::A_LD;
#pragma unroll
            for (int k = 0; k < C::KC; ++k) {
                const float4 a0 = *reinterpret_cast<const float4*>(Ac + k * C::A_LD + m0);
                const float4 a1 = *reinterpret_cast<const float4*>(Ac + k * C::A_LD + m0 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(Bc + k * C::N + colA);
                const float4 b1 = *reinterpret_cast<const float4*>(Bc + k * C::N + colB);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[r][j] = fmaf(a[r], b[j], acc[r][j]);
            }
            __syncthreads();
        }
        // ---- LSTM cell (Keras 2.2.4: recurrent_activation = hard_sigmoid, activation = tanh) ----
        float hA[8], hB[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            {
                const float ig = hard_sigmoid(acc[r][0]), fg = hard_sigmoid(acc[r][1]);
                const float gg = tanhf(acc[r][2]), og = hard_sigmoid(acc[r][3]);
                const float c = fmaf(fg, c_state[r][0], ig * gg);
                c_state[r][0] = c;
                hA[r] = og * tanhf(c);
            }
            {
                const float ig = hard_sigmoid(acc[r][4]), fg = hard_sigmoid(acc[r][5]);
                const float gg = tanhf(acc[r][6]), og = hard_sigmoid(acc[r][7]);
                const float c = fmaf(fg, c_state[r][1], ig * gg);
                c_state[r][1] = c;
                hB[r] = og * tanhf(c);
            }
        }
        // h_t -> As h rows (all threads are past the last read of As: trailing __syncthreads above)
        *reinterpret_cast<float4*>(As + (C::IN + uA) * C::A_LD + m0) = make_float4(hA[0], hA[1], hA[2], hA[3]);
        *reinterpret_cast<float4*>(As + (C::IN + uA) * C::A_LD + m0 + 4) = make_float4(hA[4], hA[5], hA[6], hA[7]);
        *reinterpret_cast<float4*>(As + (C::IN + uB) * C::A_LD + m0) = make_float4(hB[0], hB[1], hB[2], hB[3]);
        *reinterpret_cast<float4*>(As + (C::IN + uB) * C::A_LD + m0 + 4) = make_float4(hB[4], hB[5], hB[6], hB[7]);
        // output (input time order; Bidirectional concat = [fwd | bwd]) with the following BN folded in
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t w = w0 + m0 + r;
            if (w < n_win) {
                if (OMODE == 0) {
                    const int64_t off = (w * T + t) * (2 * U) + dir * U;
                    act_out[off + uA] = fmaf(hA[r], bnsA, bntA);
                    act_out[off + uB] = fmaf(hB[r], bnsB, bntB);
                } else {
                    // OMODE 1: raw h.  OMODE 2: BatchNormalization applied in fp32 first (read_rnn1: its BN has
                    // zero-variance channels with a x31.6 gain and large offsets -- folding it into the next GEMM
                    // would cancel catastrophically in split-fp16), then split.
                    // row(t, w) = t*nwp + w when a padded time-major layout is requested, else (w, t)
                    const int64_t off = (out_nwp ? ((int64_t)t * out_nwp + w) : (w * T + t)) * out_ld + dir * U;
                    const float yA = (OMODE == 2) ? fmaf(hA[r], bnsA, bntA) : hA[r];
                    const float yB = (OMODE == 2) ? fmaf(hB[r], bnsB, bntB) : hB[r];
                    const __half h1 = __float2half_rn(yA), h2 = __float2half_rn(yB);
                    out_hi[off + uA] = h1; out_lo[off + uA] = __float2half_rn(yA - __half2float(h1));
                    out_hi[off + uB] = h2; out_lo[off + uB] = __float2half_rn(yB - __half2float(h2));
                }
            }
        }
        // the first __syncthreads of the next step's K loop orders these smem writes before any read
    }
}

template <int IN_A, int IN_B, int U, int TM, int ZMODE, int OMODE>
static int launch_one(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st) {
    using C = LstmTile<IN_A, IN_B, U, TM>;
    auto kern = lstm_layer_kernel<IN_A, IN_B, U, TM, ZMODE, OMODE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    dim3 grid((unsigned)((n_win + TM - 1) / TM), 2);
    float* const* w = ZMODE ? L.wrec : L.wcat;
    kern<<<grid, C::NT, C::SMEM, st>>>(io.act_in, io.base_in, io.win_base, w[0], w[1], L.bias[0], L.bias[1],
                                       L.bn_scale, L.bn_shift, io.act_out, io.zin, io.zsig, io.out_hi, io.out_lo, io.out_ld,
                                       io.out_nwp, n_win, T);
    return 1;
}

// variant 0: fused fp32 path (projection inside the recurrence GEMM), fp32+BN output
// variant 1: fused fp32 path, raw fp16 (hi, lo) output             (feeds a tensor-core projection)
// variant 2: recurrence only, pre-activations from the tensor-core projection, fp16 (hi, lo) output
// variant 3: recurrence only, fp32 output (last layer; no BN follows)
int launch_lstm_layer(int layer, int variant, const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T,
                      cudaStream_t st) {
    if (n_win <= 0) return 0;
    switch (layer * 4 + variant) {
        case 0 * 4 + 0: return launch_one<0, 6, 16, 128, 0, 0>(L, io, n_win, T, st);
        case 0 * 4 + 1: return launch_one<0, 6, 16, 128, 0, 2>(L, io, n_win, T, st);
        case 1 * 4 + 0: return launch_one<32, 0, 64, 128, 0, 0>(L, io, n_win, T, st);
        case 1 * 4 + 1: return launch_one<32, 0, 64, 128, 0, 1>(L, io, n_win, T, st);
        case 2 * 4 + 0: return launch_one<128, 64, 128, 64, 0, 0>(L, io, n_win, T, st);
        case 2 * 4 + 2: return launch_one<0, 0, 128, 64, 1, 1>(L, io, n_win, T, st);
        case 3 * 4 + 0: return launch_one<256, 0, 64, 128, 0, 0>(L, io, n_win, T, st);
        case 3 * 4 + 3: return launch_one<0, 0, 64, 128, 1, 0>(L, io, n_win, T, st);
    }
    return -1;
}

}  // namespace nrv
