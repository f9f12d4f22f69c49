// Dense heads of the predict model (lstmmodel.py:56-63 / :108-115), fused per tile of windows:
//   per timestep Dense(128,relu) -> Dense(32,relu) -> main_out Dense(6,relu); Flatten (t*6+k);
//   feature Dense(16,relu); final_out Dense(n_class, softmax); argmax (first maximum).
#include <algorithm>

#include "nrv_common.cuh"

namespace nrv {

constexpr int HD_WIN = 8;                 // windows per CTA
constexpr int HD_THREADS = 256;
constexpr int HD_LD = 132;                // 128 + 4 (float4 aligned, conflict-free row groups)

template <int T>
struct HeadsSmem {
    float in[HD_WIN * T][HD_LD];
    float d1[HD_WIN * T][HD_LD];
    float d2[HD_WIN * T][36];
    float d3[HD_WIN][T * 6 + 2];
    float ft[HD_WIN][16];
};

template <int T, bool D1_DONE>
__global__ void __launch_bounds__(HD_THREADS)
heads_kernel(HeadsDev H, const float* __restrict__ act_in, int64_t n_win, int64_t in_nwp, float* __restrict__ probs,
             uint8_t* __restrict__ labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeadsSmem<T>& s = *reinterpret_cast<HeadsSmem<T>*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t w0 = (int64_t)blockIdx.x * HD_WIN;
    constexpr int ROWS = HD_WIN * T;
    // ---- load [ROWS][128] ------------------------------------------------------------------
    for (int i = tid; i < ROWS * 32; i += HD_THREADS) {
        const int row = i >> 5, q = i & 31;
        const int64_t w = w0 + row / T;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (w < n_win) {
            const int64_t grow = in_nwp ? ((int64_t)(row % T) * in_nwp + w) : (w0 * T + row);
            v = __ldg(reinterpret_cast<const float4*>(act_in + grow * 128) + q);
        }
        if (D1_DONE) *reinterpret_cast<float4*>(&s.d1[row][q * 4]) = v;
        else *reinterpret_cast<float4*>(&s.in[row][q * 4]) = v;
    }
    __syncthreads();
    const int g = tid >> 5;          // window inside the tile (one warp per window)
    const int tx = tid & 31;
    // ---- Dense(128 -> 128, relu): thread = T rows x 4 cols -------------------------------------
    if (!D1_DONE) {
        float acc[T][4];
        const float4 bb = __ldg(reinterpret_cast<const float4*>(H.d1b) + tx);
#pragma unroll
        for (int r = 0; r < T; ++r) { acc[r][0] = bb.x; acc[r][1] = bb.y; acc[r][2] = bb.z; acc[r][3] = bb.w; }
        const float4* Wp = reinterpret_cast<const float4*>(H.d1k) + tx;
        for (int k = 0; k < 128; k += 4) {
            const float4 wv0 = __ldg(Wp + (k + 0) * 32), wv1 = __ldg(Wp + (k + 1) * 32);
            const float4 wv2 = __ldg(Wp + (k + 2) * 32), wv3 = __ldg(Wp + (k + 3) * 32);
#pragma unroll
            for (int r = 0; r < T; ++r) {
                const float4 a = *reinterpret_cast<const float4*>(&s.in[g * T + r][k]);
                acc[r][0] = fmaf(a.x, wv0.x, acc[r][0]); acc[r][1] = fmaf(a.x, wv0.y, acc[r][1]);
                acc[r][2] = fmaf(a.x, wv0.z, acc[r][2]); acc[r][3] = fmaf(a.x, wv0.w, acc[r][3]);
                acc[r][0] = fmaf(a.y, wv1.x, acc[r][0]); acc[r][1] = fmaf(a.y, wv1.y, acc[r][1]);
                acc[r][2] = fmaf(a.y, wv1.z, acc[r][2]); acc[r][3] = fmaf(a.y, wv1.w, acc[r][3]);
                acc[r][0] = fmaf(a.z, wv2.x, acc[r][0]); acc[r][1] = fmaf(a.z, wv2.y, acc[r][1]);
                acc[r][2] = fmaf(a.z, wv2.z, acc[r][2]); acc[r][3] = fmaf(a.z, wv2.w, acc[r][3]);
                acc[r][0] = fmaf(a.w, wv3.x, acc[r][0]); acc[r][1] = fmaf(a.w, wv3.y, acc[r][1]);
                acc[r][2] = fmaf(a.w, wv3.z, acc[r][2]); acc[r][3] = fmaf(a.w, wv3.w, acc[r][3]);
            }
        }
#pragma unroll
        for (int r = 0; r < T; ++r)
            *reinterpret_cast<float4*>(&s.d1[g * T + r][tx * 4]) =
                make_float4(fmaxf(acc[r][0], 0.f), fmaxf(acc[r][1], 0.f), fmaxf(acc[r][2], 0.f), fmaxf(acc[r][3], 0.f));
    }
    __syncwarp();
    // ---- Dense(128 -> 32, relu): thread = T rows x 1 col ---------------------------------------
    {
        float acc[T];
        const float bb = __ldg(H.d2b + tx);
#pragma unroll
        for (int r = 0; r < T; ++r) acc[r] = bb;
        for (int k = 0; k < 128; k += 4) {
            const float w0v = __ldg(H.d2k + (k + 0) * 32 + tx), w1v = __ldg(H.d2k + (k + 1) * 32 + tx);
            const float w2v = __ldg(H.d2k + (k + 2) * 32 + tx), w3v = __ldg(H.d2k + (k + 3) * 32 + tx);
#pragma unroll
            for (int r = 0; r < T; ++r) {
                const float4 a = *reinterpret_cast<const float4*>(&s.d1[g * T + r][k]);
                acc[r] = fmaf(a.x, w0v, acc[r]); acc[r] = fmaf(a.y, w1v, acc[r]);
                acc[r] = fmaf(a.z, w2v, acc[r]); acc[r] = fmaf(a.w, w3v, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < T; ++r) s.d2[g * T + r][tx] = fmaxf(acc[r], 0.f);
    }
    __syncwarp();
    // ---- main_out Dense(32 -> 6, relu) + Flatten (index t*6 + k) --------------------------------
    for (int o = tx; o < T * 6; o += 32) {
        const int t = o / 6, k = o - t * 6;
        float a = __ldg(H.mb + k);
#pragma unroll 8
        for (int j = 0; j < 32; ++j) a = fmaf(s.d2[g * T + t][j], __ldg(H.mk + j * 6 + k), a);
        s.d3[g][o] = fmaxf(a, 0.f);
    }
    __syncwarp();
    // ---- feature Dense(T*6 -> 16, relu) ---------------------------------------------------------
    if (tx < 16) {
        float a = __ldg(H.fb + tx);
        for (int j = 0; j < T * 6; ++j) a = fmaf(s.d3[g][j], __ldg(H.fk + j * 16 + tx), a);
        s.ft[g][tx] = fmaxf(a, 0.f);
    }
    __syncwarp();
    // ---- final_out Dense(16 -> n_class) + softmax + argmax --------------------------------------
    const int nc = H.n_class;
    float logit = -INFINITY;
    if (tx < nc) {
        float a = __ldg(H.ob + tx);
#pragma unroll
        for (int j = 0; j < 16; ++j) a = fmaf(s.ft[g][j], __ldg(H.ok + j * nc + tx), a);
        logit = a;
    }
    float mx = logit;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float e = (tx < nc) ? expf(logit - mx) : 0.f;
    float sum = e;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float p = e / sum;
    // argmax with numpy semantics (first maximum)
    float bp = (tx < nc) ? p : -1.f;
    int bi = tx;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        const float op = __shfl_xor_sync(0xffffffffu, bp, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (op > bp || (op == bp && oi < bi)) { bp = op; bi = oi; }
    }
    const int64_t w = w0 + g;
    if (w < n_win) {
        if (probs && tx < nc) probs[w * nc + tx] = p;
        if (labels && tx == 0) labels[w] = (uint8_t)bi;
    }
}

// ---- tail of the heads when relu(Dense(128)) already ran on the tensor cores (nrv_gemm.cu) ---------------------
// Persistent CTAs, all small weights resident in shared memory, one warp per window.
template <int T>
struct TailSmem {
    float d1[HD_WIN * T][HD_LD];
    float d2[HD_WIN * T][36];
    float d3[HD_WIN][T * 6 + 2];
    float ft[HD_WIN][16];
    float w2[128][32];
    float b2[32];
    float mk[32][6];
    float mb[8];
    float fk[T * 6][16];
    float fb[16];
    float ok[16][8];
    float ob[8];
};

template <int T, bool D2_DONE>
__global__ void __launch_bounds__(HD_THREADS, 2)
heads_tail_kernel(HeadsDev H, const float* __restrict__ d1_in, int64_t n_win, int64_t in_nwp, float* __restrict__ probs,
                  uint8_t* __restrict__ labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TailSmem<T>& s = *reinterpret_cast<TailSmem<T>*>(smem_raw);
    const int tid = threadIdx.x;
    const int nc = H.n_class;
    if (!D2_DONE)
        for (int i = tid; i < 128 * 32; i += HD_THREADS) (&s.w2[0][0])[i] = __ldg(H.d2k + i);
    for (int i = tid; i < 32 * 6; i += HD_THREADS) (&s.mk[0][0])[i] = __ldg(H.mk + i);
    for (int i = tid; i < T * 6 * 16; i += HD_THREADS) (&s.fk[0][0])[i] = __ldg(H.fk + i);
    if (tid < 32) s.b2[tid] = __ldg(H.d2b + tid);
    if (tid < 6) s.mb[tid] = __ldg(H.mb + tid);
    if (tid < 16) s.fb[tid] = __ldg(H.fb + tid);
    if (tid < 16 * 8) s.ok[tid >> 3][tid & 7] = ((tid & 7) < nc) ? __ldg(H.ok + (tid >> 3) * nc + (tid & 7)) : 0.f;
    if (tid < 8) s.ob[tid] = (tid < nc) ? __ldg(H.ob + tid) : 0.f;
    const int g = tid >> 5, tx = tid & 31;
    constexpr int ROWS = HD_WIN * T;
    const int64_t n_groups = (n_win + HD_WIN - 1) / HD_WIN;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t w0 = grp * HD_WIN;
        __syncthreads();                                   // previous group's tile fully consumed / weights loaded
        if (D2_DONE) {
            // input is relu(Dense(32)) [rows][32]: 8 float4 per row
            for (int i = tid; i < ROWS * 8; i += HD_THREADS) {
                const int row = i >> 3, q = i & 7;
                const int64_t w = w0 + row / T;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w < n_win) {
                    const int64_t grow = in_nwp ? ((int64_t)(row % T) * in_nwp + w) : (w * T + row % T);
                    v = __ldg(reinterpret_cast<const float4*>(d1_in + grow * 32) + q);
                }
                *reinterpret_cast<float4*>(&s.d2[row][q * 4]) = v;
            }
            __syncthreads();
        } else {
            for (int i = tid; i < ROWS * 32; i += HD_THREADS) {
                const int row = i >> 5, q = i & 31;
                const int64_t w = w0 + row / T;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (w < n_win) {
                    const int64_t grow = in_nwp ? ((int64_t)(row % T) * in_nwp + w) : (w * T + row % T);
                    v = __ldg(reinterpret_cast<const float4*>(d1_in + grow * 128) + q);
                }
                *reinterpret_cast<float4*>(&s.d1[row][q * 4]) = v;
            }
            __syncthreads();
            // ---- Dense(128 -> 32, relu): lane = output column, T rows per lane ----
            float acc[T];
#pragma unroll
            for (int r = 0; r < T; ++r) acc[r] = s.b2[tx];
#pragma unroll 4
            for (int k = 0; k < 128; k += 4) {
                const float w0v = s.w2[k][tx], w1v = s.w2[k + 1][tx], w2v = s.w2[k + 2][tx], w3v = s.w2[k + 3][tx];
#pragma unroll
                for (int r = 0; r < T; ++r) {
                    const float4 a = *reinterpret_cast<const float4*>(&s.d1[g * T + r][k]);
                    acc[r] = fmaf(a.x, w0v, acc[r]); acc[r] = fmaf(a.y, w1v, acc[r]);
                    acc[r] = fmaf(a.z, w2v, acc[r]); acc[r] = fmaf(a.w, w3v, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < T; ++r) s.d2[g * T + r][tx] = fmaxf(acc[r], 0.f);
        }
        __syncwarp();
        // ---- main_out Dense(32 -> 6, relu) + Flatten (index t*6 + k) ----
        for (int o = tx; o < T * 6; o += 32) {
            const int t = o / 6, k = o - t * 6;
            float a0 = s.mb[k], a1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                a0 = fmaf(s.d2[g * T + t][j], s.mk[j][k], a0);
                a1 = fmaf(s.d2[g * T + t][j + 1], s.mk[j + 1][k], a1);
            }
            s.d3[g][o] = fmaxf(a0 + a1, 0.f);
        }
        __syncwarp();
        // ---- feature Dense(T*6 -> 16, relu): two half-sums per output, combined by shuffle ----
        {
            const int o = tx & 15, half = tx >> 4;
            constexpr int HALF = (T * 6 + 1) / 2;
            float a = 0.f;
            for (int j = half * HALF; j < min((half + 1) * HALF, T * 6); ++j) a = fmaf(s.d3[g][j], s.fk[j][o], a);
            a += __shfl_xor_sync(0xffffffffu, a, 16);
            if (tx < 16) s.ft[g][tx] = fmaxf(a + s.fb[tx], 0.f);
        }
        __syncwarp();
        // ---- final_out Dense(16 -> n_class) + softmax + argmax (first maximum) ----
        float logit = -INFINITY;
        if (tx < nc) {
            float a = s.ob[tx];
#pragma unroll
            for (int j = 0; j < 16; ++j) a = fmaf(s.ft[g][j], s.ok[j][tx], a);
            logit = a;
        }
        float mx = logit;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float e = (tx < nc) ? expf(logit - mx) : 0.f;
        float sum = e;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float pr = e / sum;
        float bp = (tx < nc) ? pr : -1.f;
        int bi = tx;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const float op = __shfl_xor_sync(0xffffffffu, bp, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (op > bp || (op == bp && oi < bi)) { bp = op; bi = oi; }
        }
        const int64_t w = w0 + g;
        if (w < n_win) {
            if (probs && tx < nc) probs[w * nc + tx] = pr;
            if (labels && tx == 0) labels[w] = (uint8_t)bi;
        }
    }
}

// ---- tail of the heads when relu(Dense(128)) AND relu(Dense(32)) already ran on the tensor cores ---------------------
// One THREAD per window: main_out Dense(32 -> 6, relu) for the T timesteps, Flatten (t*6 + k), feature Dense(16, relu),
// final_out Dense(n_class) + softmax + argmax, all in registers; the small weights are warp-wide broadcasts from shared
// memory.  (The warp-per-window kernel above spent 6 ms per step on 0.3 ms worth of FMAs: 66 outputs over 32 lanes,
// two shared-memory reads per FMA.)  Input rows are time-major: row(t, w) = t*nwp + w, 32 floats each.
constexpr int HT_THREADS = 128;

template <int T>
__global__ void __launch_bounds__(HT_THREADS)
heads_tail_thread_kernel(HeadsDev H, const float* __restrict__ d2_in, int64_t n_win, int64_t in_nwp, float* __restrict__ probs,
                         uint8_t* __restrict__ labels) {
    __shared__ __align__(16) float s_mk[32][8];        // main_out kernel, 6 columns padded to 8
    __shared__ __align__(16) float s_fk[T * 6][16];
    __shared__ __align__(16) float s_ok[16][8];
    __shared__ __align__(16) float s_mb[8], s_fb[16], s_ob[8];
    const int tid = threadIdx.x;
    const int nc = H.n_class;
    for (int i = tid; i < 32 * 8; i += HT_THREADS) s_mk[i >> 3][i & 7] = ((i & 7) < 6) ? __ldg(H.mk + (i >> 3) * 6 + (i & 7)) : 0.f;
    for (int i = tid; i < T * 6 * 16; i += HT_THREADS) (&s_fk[0][0])[i] = __ldg(H.fk + i);
    for (int i = tid; i < 16 * 8; i += HT_THREADS) s_ok[i >> 3][i & 7] = ((i & 7) < nc) ? __ldg(H.ok + (i >> 3) * nc + (i & 7)) : 0.f;
    if (tid < 8) { s_mb[tid] = (tid < 6) ? __ldg(H.mb + tid) : 0.f; s_ob[tid] = (tid < nc) ? __ldg(H.ob + tid) : 0.f; }
    if (tid < 16) s_fb[tid] = __ldg(H.fb + tid);
    __syncthreads();
    const int64_t w = (int64_t)blockIdx.x * HT_THREADS + tid;
    if (w >= n_win) return;
    float ft[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) ft[o] = s_fb[o];
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
        const float4* rowp = reinterpret_cast<const float4*>(d2_in + ((int64_t)t * in_nwp + w) * 32);
        float4 r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = __ldg(rowp + q);
        float m[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) m[k] = s_mb[k];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float a[4] = {r[q].x, r[q].y, r[q].z, r[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 wa = *reinterpret_cast<const float4*>(&s_mk[q * 4 + e][0]);
                const float2 wb = *reinterpret_cast<const float2*>(&s_mk[q * 4 + e][4]);
                m[0] = fmaf(a[e], wa.x, m[0]); m[1] = fmaf(a[e], wa.y, m[1]); m[2] = fmaf(a[e], wa.z, m[2]);
                m[3] = fmaf(a[e], wa.w, m[3]); m[4] = fmaf(a[e], wb.x, m[4]); m[5] = fmaf(a[e], wb.y, m[5]);
            }
        }
        // feature Dense: flattened index t*6 + k
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float a = fmaxf(m[k], 0.f);
#pragma unroll
            for (int o4 = 0; o4 < 4; ++o4) {
                const float4 fw = *reinterpret_cast<const float4*>(&s_fk[t * 6 + k][o4 * 4]);
                ft[o4 * 4 + 0] = fmaf(a, fw.x, ft[o4 * 4 + 0]); ft[o4 * 4 + 1] = fmaf(a, fw.y, ft[o4 * 4 + 1]);
                ft[o4 * 4 + 2] = fmaf(a, fw.z, ft[o4 * 4 + 2]); ft[o4 * 4 + 3] = fmaf(a, fw.w, ft[o4 * 4 + 3]);
            }
        }
    }
    float logit[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) logit[k] = s_ob[k];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float a = fmaxf(ft[j], 0.f);
        const float4 wa = *reinterpret_cast<const float4*>(&s_ok[j][0]), wb = *reinterpret_cast<const float4*>(&s_ok[j][4]);
        logit[0] = fmaf(a, wa.x, logit[0]); logit[1] = fmaf(a, wa.y, logit[1]); logit[2] = fmaf(a, wa.z, logit[2]);
        logit[3] = fmaf(a, wa.w, logit[3]); logit[4] = fmaf(a, wb.x, logit[4]); logit[5] = fmaf(a, wb.y, logit[5]);
        logit[6] = fmaf(a, wb.z, logit[6]); logit[7] = fmaf(a, wb.w, logit[7]);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < nc) mx = fmaxf(mx, logit[k]);
    float e[8], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = (k < nc) ? expf(logit[k] - mx) : 0.f; sum += e[k]; }
    float bp = -1.f;
    int bi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float pr = e[k] / sum;
        if (k < nc) {
            if (probs) probs[w * nc + k] = pr;
            if (pr > bp) { bp = pr; bi = k; }          // first maximum (numpy argmax semantics)
        }
    }
    if (labels) labels[w] = (uint8_t)bi;
}

template <int T>
static int launch_heads_tail_thread_t(const HeadsDev& H, const float* d2, int64_t n_win, int64_t in_nwp, float* probs,
                                      uint8_t* labels, cudaStream_t st) {
    const unsigned grid = (unsigned)((n_win + HT_THREADS - 1) / HT_THREADS);
    heads_tail_thread_kernel<T><<<grid, HT_THREADS, 0, st>>>(H, d2, n_win, in_nwp, probs, labels);
    return 1;
}

template <int T, bool D2_DONE>
static int launch_heads_tail_t(const HeadsDev& H, const float* d1, int64_t n_win, int64_t in_nwp, float* probs,
                               uint8_t* labels, cudaStream_t st) {
    auto kern = heads_tail_kernel<T, D2_DONE>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem<T>));
    const int64_t n_groups = (n_win + HD_WIN - 1) / HD_WIN;
    const unsigned grid = (unsigned)std::min<int64_t>(n_groups, 2 * 148);
    kern<<<grid, HD_THREADS, sizeof(TailSmem<T>), st>>>(H, d1, n_win, in_nwp, probs, labels);
    return 1;
}

template <int T, bool D1>
static int launch_heads_t(const HeadsDev& H, const float* act_in, int64_t n_win, int64_t in_nwp, float* probs,
                          uint8_t* labels, cudaStream_t st) {
    auto kern = heads_kernel<T, D1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HeadsSmem<T>));
    const unsigned grid = (unsigned)((n_win + HD_WIN - 1) / HD_WIN);
    kern<<<grid, HD_THREADS, sizeof(HeadsSmem<T>), st>>>(H, act_in, n_win, in_nwp, probs, labels);
    return 1;
}

int launch_heads(const HeadsDev& H, const float* act_in, int64_t n_win, int T, float* probs, uint8_t* labels,
                 int stage, int64_t in_nwp, cudaStream_t st) {
    if (n_win <= 0) return 0;
#define NRV_HEADS_CASE(TT)                                                                        \
    case TT:                                                                                      \
        return stage == 2   ? (in_nwp ? launch_heads_tail_thread_t<TT>(H, act_in, n_win, in_nwp, probs, labels, st)     \
                                      : launch_heads_tail_t<TT, true>(H, act_in, n_win, in_nwp, probs, labels, st)) \
               : stage == 1 ? launch_heads_tail_t<TT, false>(H, act_in, n_win, in_nwp, probs, labels, st)  \
                            : launch_heads_t<TT, false>(H, act_in, n_win, in_nwp, probs, labels, st);
    switch (T) {   // W is read from the weights (feature.kernel.shape[0] / 6); the shipped files have 11
        NRV_HEADS_CASE(5)
        NRV_HEADS_CASE(7)
        NRV_HEADS_CASE(9)
        NRV_HEADS_CASE(11)
        NRV_HEADS_CASE(13)
    }
#undef NRV_HEADS_CASE
    return -1;
}

}  // namespace nrv
