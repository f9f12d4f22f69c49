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

// ============================================================================================================
// read_rnn1 (lstmmodel.py:44: Bidirectional(LSTM(16)) on the 6 feature columns) for the tensor-core path.
// The layer is tiny (K = 6 + 16, N = 64): the generic tile kernel above pads K to 32, re-streams the weights every
// step and scatters 2-byte stores.  Here ONE THREAD owns one (window, direction): x_t comes straight from the
// per-base feature table (indexed by first base + t, never materialised per window), h and c stay in registers
// for all T steps, the 22 x 64 weights sit in shared memory and are read as warp-wide broadcasts (LDS.128), and
// each step ends in two 32-byte stores of BatchNormalization(h) as an fp16 (hi, lo) pair into row (t, w) of the
// padded time-major operand of read_rnn11 (BN is applied here, not folded: see nrv_api.cu pack_model).
// ============================================================================================================
constexpr int R1_THREADS = 128;

__device__ __forceinline__ float r1_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float r1_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// tanh(|x|) = 1 - 2 / (exp(2|x|) + 1): abs error ~2e-7 (same form as the tensor-core recurrences)
__device__ __forceinline__ float r1_tanh(float x) {
    const float e = r1_ex2(fabsf(x) * 2.8853900817779268f);
    return copysignf(fmaf(-2.f, r1_rcp(e + 1.f), 1.f), x);
}

constexpr int R1_WPT = 2;                              // windows per thread: every weight broadcast (LDS.128) feeds 4 x R1_WPT FMAs

__global__ void __launch_bounds__(R1_THREADS, 3)
read_rnn1_kernel(const float* __restrict__ x, const int32_t* __restrict__ win_base, const float* __restrict__ wcat0,
                 const float* __restrict__ wcat1, const float* __restrict__ bias0, const float* __restrict__ bias1,
                 const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, __half* __restrict__ out_hi,
                 __half* __restrict__ out_lo, int out_ld, int64_t nwp, int64_t n_win, int T) {
    constexpr int U = 16, IN = 6, K = IN + U, N = 4 * U, P = R1_WPT;
    __shared__ __align__(16) float s_w[K * N];        // [k][unit*4 + gate]
    __shared__ __align__(16) float s_b[N];
    __shared__ float s_bn[2 * U];                      // scale | shift of this direction
    const int dir = blockIdx.y;
    const float* wcat = dir ? wcat1 : wcat0;
    const float* bias = dir ? bias1 : bias0;
    for (int i = threadIdx.x; i < K * N; i += R1_THREADS) s_w[i] = wcat[i];
    if (threadIdx.x < N) s_b[threadIdx.x] = bias[threadIdx.x];
    if (threadIdx.x < U) {
        s_bn[threadIdx.x] = bn_scale[dir * U + threadIdx.x];
        s_bn[U + threadIdx.x] = bn_shift[dir * U + threadIdx.x];
    }
    __syncthreads();
    int64_t w[P], b0[P];
    bool ok[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        w[p] = ((int64_t)blockIdx.x * P + p) * R1_THREADS + threadIdx.x;
        ok[p] = w[p] < n_win;
        b0[p] = ok[p] ? win_base[w[p]] : 0;
    }
    if (!ok[0]) return;
    float in[P][K];                                    // x_t (6) | h_{t-1} (16)
    float c[P][U];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int j = 0; j < U; ++j) { in[p][IN + j] = 0.f; c[p][j] = 0.f; }
    for (int s = 0; s < T; ++s) {
        const int t = dir ? (T - 1 - s) : s;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const float2* xp = reinterpret_cast<const float2*>(x + (b0[p] + t) * IN);
            const float2 x0 = __ldg(xp), x1 = __ldg(xp + 1), x2 = __ldg(xp + 2);
            in[p][0] = x0.x; in[p][1] = x0.y; in[p][2] = x1.x; in[p][3] = x1.y; in[p][4] = x2.x; in[p][5] = x2.y;
        }
        float hn[P][U];
#pragma unroll
        for (int g = 0; g < 8; ++g) {                  // 2 units = 8 gate columns at a time
            float z[P][8];
            {
                const float4 ba = *reinterpret_cast<const float4*>(s_b + g * 8), bb = *reinterpret_cast<const float4*>(s_b + g * 8 + 4);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    z[p][0] = ba.x; z[p][1] = ba.y; z[p][2] = ba.z; z[p][3] = ba.w;
                    z[p][4] = bb.x; z[p][5] = bb.y; z[p][6] = bb.z; z[p][7] = bb.w;
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4 wa = *reinterpret_cast<const float4*>(s_w + k * N + g * 8);
                const float4 wb = *reinterpret_cast<const float4*>(s_w + k * N + g * 8 + 4);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const float a = in[p][k];
                    z[p][0] = fmaf(a, wa.x, z[p][0]); z[p][1] = fmaf(a, wa.y, z[p][1]);
                    z[p][2] = fmaf(a, wa.z, z[p][2]); z[p][3] = fmaf(a, wa.w, z[p][3]);
                    z[p][4] = fmaf(a, wb.x, z[p][4]); z[p][5] = fmaf(a, wb.y, z[p][5]);
                    z[p][6] = fmaf(a, wb.z, z[p][6]); z[p][7] = fmaf(a, wb.w, z[p][7]);
                }
            }
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = g * 2 + q;
                    const float ig = hard_sigmoid(z[p][q * 4 + 0]), fg = hard_sigmoid(z[p][q * 4 + 1]);
                    const float gg = r1_tanh(z[p][q * 4 + 2]), og = hard_sigmoid(z[p][q * 4 + 3]);
                    const float cn = fmaf(fg, c[p][j], ig * gg);
                    c[p][j] = cn;
                    hn[p][j] = og * r1_tanh(cn);
                }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            uint32_t ph[8], pl[8];
#pragma unroll
            for (int p2 = 0; p2 < 8; ++p2) {
                in[p][IN + 2 * p2] = hn[p][2 * p2]; in[p][IN + 2 * p2 + 1] = hn[p][2 * p2 + 1];
                const float y0 = fmaf(hn[p][2 * p2], s_bn[2 * p2], s_bn[U + 2 * p2]);
                const float y1 = fmaf(hn[p][2 * p2 + 1], s_bn[2 * p2 + 1], s_bn[U + 2 * p2 + 1]);
                const __half2 hi = __floats2half2_rn(y0, y1);
                const float2 hf = __half22float2(hi);
                const __half2 lo = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
                ph[p2] = *reinterpret_cast<const uint32_t*>(&hi); pl[p2] = *reinterpret_cast<const uint32_t*>(&lo);
            }
            if (ok[p]) {
                const int64_t off = ((int64_t)t * nwp + w[p]) * out_ld + dir * U;
                uint4* oh = reinterpret_cast<uint4*>(out_hi + off);
                uint4* ol = reinterpret_cast<uint4*>(out_lo + off);
                oh[0] = make_uint4(ph[0], ph[1], ph[2], ph[3]); oh[1] = make_uint4(ph[4], ph[5], ph[6], ph[7]);
                ol[0] = make_uint4(pl[0], pl[1], pl[2], pl[3]); ol[1] = make_uint4(pl[4], pl[5], pl[6], pl[7]);
            }
        }
    }
}

int launch_read_rnn1(const LstmLayerDev& L, const LstmIo& io, int64_t n_win, int T, cudaStream_t st) {
    if (n_win <= 0) return 0;
    if (L.u != 16 || L.in_b != 6 || L.in_a != 0 || !io.out_hi || !io.out_lo || !io.out_nwp || (io.out_ld & 7) || !L.bn_scale) return -1;
    dim3 grid((unsigned)((n_win + R1_THREADS * R1_WPT - 1) / (R1_THREADS * R1_WPT)), 2);
    read_rnn1_kernel<<<grid, R1_THREADS, 0, st>>>(io.base_in, io.win_base, L.wcat[0], L.wcat[1], L.bias[0], L.bias[1], L.bn_scale,
                                                   L.bn_shift, io.out_hi, io.out_lo, io.out_ld, io.out_nwp, n_win, T);
    return 1;
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
