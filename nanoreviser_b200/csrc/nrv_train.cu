// Training operators (include/nrv_train.h): forward + backward of the layers of nanorevutils/lstmmodel.py:32-133 and
// nanorevcnn.py:17-38 in training mode, the two losses of the train model (lstmmodel.py:65-74) and the Adam update.
// One kernel per operator (the GEMM on the tensor cores as 3 x TF32, the rest fp32 SIMT); the orchestration (which tensors, which
// order, the CUDA graph around a step) is nanoreviser_b200/train.py.  A training step at the reference's batch size (512 windows) is a
// chain of ~740 small dependent launches; these kernels are written for being right first: tests/test_train_gpu.py holds every
// gradient against an fp64 autograd graph of the same network.
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "../../include/nrv_train.h"

namespace {

thread_local std::string g_err;

int fail(const char* what, cudaError_t e) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return -1;
}
int bad(const char* what) {
    g_err = what;
    return -2;
}
#define LAUNCH_CHECK(what)                                 \
    do {                                                   \
        const cudaError_t e_ = cudaGetLastError();         \
        if (e_ != cudaSuccess) return fail(what, e_);      \
    } while (0)

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline unsigned blocks_for(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

// ---------------------------------------------------------------------------------------------------------------------------
// GEMM: 64 x 64 tile of C per CTA, K in steps of 16.  The products run on the tensor cores as 3 x TF32 (mma.sync m16n8k8):
// every fp32 operand is split into hi = tf32(x) and lo = tf32(x - hi) and a.b ~ a_hi.b_hi + a_lo.b_hi + a_hi.b_lo with fp32
// accumulation, which keeps fp32-level accuracy (the gradient tests hold every tensor to 1e-3 of fp64 autograd, measured 1e-6 .. 3e-5).
// 8 warps as 2 (M) x 4 (N): a warp owns a 32 x 16 sub-tile = 2 x 2 MMA tiles.
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// SPLITK: blockIdx.z owns the K range [z * kchunk, (z + 1) * kchunk) and ADDS alpha * partial into C with atomics (C holds beta * C
// already): the weight-gradient products X^T dZ have K = all rows of the batch (thousands) and a C of a few tiles only
template <bool TA, bool TB, bool SPLITK>
__global__ void __launch_bounds__(256)
gemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float beta,
            float* __restrict__ C, int ldc, int kchunk) {
    constexpr int LD = 72;                               // row stride of the staged tiles: fragment loads hit 32 distinct banks
    __shared__ float As[16][LD], Bs[16][LD];             // As[k][m], Bs[k][n]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp >> 2) * 32, wn = (warp & 3) * 16;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[2][2][4] = {};
    const int kbeg = SPLITK ? blockIdx.z * kchunk : 0;
    const int kend = SPLITK ? min(K, kbeg + kchunk) : K;
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;
            {   // A tile -> As[k][m]
                const int kk = TA ? idx >> 6 : idx & 15, mm = TA ? idx & 63 : idx >> 4;
                const int gm = m0 + mm, gk = k0 + kk;
                float v = 0.f;
                if (gm < M && gk < kend) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
                As[kk][mm] = v;
            }
            {   // B tile -> Bs[k][n]
                const int kk = TB ? idx & 15 : idx >> 6, nn = TB ? idx >> 4 : idx & 63;
                const int gn = n0 + nn, gk = k0 + kk;
                float v = 0.f;
                if (gn < N && gk < kend) v = TB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
                Bs[kk][nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 16; ks += 8) {
            uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {             // A fragment (16 x 8, row): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)
                const int r = wm + mi * 16 + g;
                tf32_split(As[ks + t][r], ah[mi][0], al[mi][0]);
                tf32_split(As[ks + t][r + 8], ah[mi][1], al[mi][1]);
                tf32_split(As[ks + t + 4][r], ah[mi][2], al[mi][2]);
                tf32_split(As[ks + t + 4][r + 8], ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {             // B fragment (8 x 8, col): b0 (k = t, n = g) b1 (k = t + 4, n = g)
                const int c = wn + ni * 8 + g;
                tf32_split(Bs[ks + t][c], bh[ni][0], bl[ni][0]);
                tf32_split(Bs[ks + t + 4][c], bh[ni][1], bl[ni][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
                }
        }
        __syncthreads();
    }
    // C fragment: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int gm = m0 + wm + mi * 16 + g + (e >> 1) * 8, gn = n0 + wn + ni * 8 + 2 * t + (e & 1);
                if (gm >= M || gn >= N) continue;
                float* c = C + (size_t)gm * ldc + gn;
                const float v = acc[mi][ni][e];
                if (SPLITK) atomicAdd(c, alpha * v);
                else *c = beta == 0.f ? alpha * v : fmaf(alpha, v, beta * *c);
            }
}

// C[M,N] (ldc) *= beta (beta == 0: set to zero, also over NaNs)
__global__ void scale2d_kernel(float* __restrict__ C, int M, int N, int ldc, float beta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    const int r = (int)(i / N), c = (int)(i - (int64_t)r * N);
    float* p = C + (size_t)r * ldc + c;
    *p = beta == 0.f ? 0.f : beta * *p;
}

__global__ void bias_act_kernel(float* __restrict__ Y, int M, int N, int ldy, const float* __restrict__ bias, int relu) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)M * N) return;
    const int r = (int)(i / N), c = (int)(i - (int64_t)r * N);
    float v = Y[(size_t)r * ldy + c] + bias[c];
    if (relu) v = fmaxf(v, 0.f);
    Y[(size_t)r * ldy + c] = v;
}

__global__ void relu_bwd_kernel(float* __restrict__ dY, const float* __restrict__ Y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(Y[i] > 0.f)) dY[i] = 0.f;
}

// column sums: a CTA = 32 columns x 8 row-lanes over the rows [blockIdx.y * rchunk, +rchunk); double partial sums, one float atomic
// per column and CTA into out (which holds beta * out already)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ X, int M, int N, int ldx, float* __restrict__ out, int rchunk) {
    __shared__ double part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int r0 = blockIdx.y * rchunk, r1 = min(M, r0 + rchunk);
    double s = 0.0;
    if (c < N)
        for (int r = r0 + rl; r < r1; r += 8) s += (double)X[(size_t)r * ldx + c];
    part[rl][threadIdx.x & 31] = s;
    __syncthreads();
    if (rl == 0 && c < N) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(&out[c], (float)t);
    }
}

__global__ void copy2d_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, int rows, int cols, int acc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float v = src[(size_t)r * lds + c];
    float* d = dst + (size_t)r * ldd + c;
    *d = acc ? *d + v : v;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Conv1D k = 3 'same' + relu, channels last
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void conv1d_fwd_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ b,
                                  float* __restrict__ Y, int64_t rows, int L, int cin, int cout) {
    __shared__ float w[3 * 8 * 8 + 8];
    for (int i = threadIdx.x; i < 3 * cin * cout; i += blockDim.x) w[i] = W[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) w[192 + i] = b[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // (sequence, position)
    if (i >= rows) return;
    const int p = (int)(i % L);
    float acc[8];
#pragma unroll
    for (int co = 0; co < 8; ++co) acc[co] = co < cout ? w[192 + co] : 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int q = p + k - 1;
        if (q < 0 || q >= L) continue;
        const float* x = X + (i + k - 1) * cin;
        for (int ci = 0; ci < cin; ++ci) {
            const float xv = x[ci];
#pragma unroll
            for (int co = 0; co < 8; ++co)
                if (co < cout) acc[co] = fmaf(xv, w[(k * cin + ci) * cout + co], acc[co]);
        }
    }
    for (int co = 0; co < cout; ++co) Y[i * cout + co] = fmaxf(acc[co], 0.f);
}

// dX[n,q,ci] = sum_k sum_co g[n, q - k + 1, co] W[k,ci,co],  g = dY * (Y > 0)
__global__ void conv1d_bwd_data_kernel(const float* __restrict__ W, const float* __restrict__ Y, const float* __restrict__ dY,
                                       float* __restrict__ dX, int64_t rows, int L, int cin, int cout) {
    __shared__ float w[3 * 8 * 8];
    for (int i = threadIdx.x; i < 3 * cin * cout; i += blockDim.x) w[i] = W[i];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows) return;
    const int q = (int)(i % L);
    float acc[8] = {};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int p = q - k + 1;
        if (p < 0 || p >= L) continue;
        const int64_t o = (i - k + 1) * cout;
        for (int co = 0; co < cout; ++co) {
            const float g = Y[o + co] > 0.f ? dY[o + co] : 0.f;
#pragma unroll
            for (int ci = 0; ci < 8; ++ci)
                if (ci < cin) acc[ci] = fmaf(g, w[(k * cin + ci) * cout + co], acc[ci]);
        }
    }
    for (int ci = 0; ci < cin; ++ci) dX[i * cin + ci] = acc[ci];
}

// dW[k,ci,co] = sum_{n,p} X[n, p + k - 1, ci] g[n,p,co]; db[co] = sum g.  One thread per weight (3 cin cout <= 192) + one per
// bias; a CTA walks whole sequences (staged in shared memory) and adds its partial sums (double) to the scratch
__global__ void __launch_bounds__(256)
conv1d_bwd_w_kernel(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ dY, double* __restrict__ acc_out,
                    int n, int L, int cin, int cout) {
    extern __shared__ float sm[];
    float* xs = sm;                          // [(L + 2) * cin], zero halo
    float* gs = sm + (L + 2) * cin;          // [L * cout]
    const int nw = 3 * cin * cout;
    const int tid = threadIdx.x;
    int k = 0, ci = 0, co = 0;
    if (tid < nw) { k = tid / (cin * cout); ci = (tid / cout) % cin; co = tid % cout; }
    else if (tid < nw + cout) co = tid - nw;
    double acc = 0.0;
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
        for (int i = tid; i < (L + 2) * cin; i += blockDim.x) {
            const int p = i / cin - 1;
            xs[i] = (p >= 0 && p < L) ? X[((size_t)s * L + p) * cin + (i % cin)] : 0.f;
        }
        for (int i = tid; i < L * cout; i += blockDim.x) {
            const size_t o = (size_t)s * L * cout + i;
            gs[i] = Y[o] > 0.f ? dY[o] : 0.f;
        }
        __syncthreads();
        if (tid < nw) {
            float a = 0.f;
            for (int p = 0; p < L; ++p) a = fmaf(xs[(p + k) * cin + ci], gs[p * cout + co], a);
            acc += (double)a;
        } else if (tid < nw + cout) {
            float a = 0.f;
            for (int p = 0; p < L; ++p) a += gs[p * cout + co];
            acc += (double)a;
        }
        __syncthreads();
    }
    if (tid < nw + cout) atomicAdd(&acc_out[tid], acc);
}
__global__ void double_to_float_kernel(const double* __restrict__ src, float* __restrict__ a, int na, float* __restrict__ b, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) a[i] = (float)src[i];
    else if (i < na + nb) b[i - na] = (float)src[i];
}

__global__ void add_bcast_kernel(float* __restrict__ Y, const float* __restrict__ X, int64_t rows, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * C) Y[i] += X[i / C];
}

// ---------------------------------------------------------------------------------------------------------------------------
// BatchNormalization, training mode.  C is a power of two <= 256: a thread of a 256-thread CTA keeps its channel while the CTA
// strides over the rows in steps of 256 / C; per-channel partial sums go through shared memory into a double scratch.
// ---------------------------------------------------------------------------------------------------------------------------
// mode 0: sum x ; 1: sum (x - mean)^2 ; 2: {sum dy, sum dy * xhat}
template <int MODE>
__global__ void __launch_bounds__(256)
bn_reduce_kernel(const float* __restrict__ X, const float* __restrict__ dY, const float* __restrict__ mean, const float* __restrict__ var,
                 float eps, double* __restrict__ out, int64_t rows, int C) {
    __shared__ double s0[256], s1[256];
    const int tid = threadIdx.x, c = tid & (C - 1);
    const int64_t total = rows * C;
    const float mu = MODE >= 1 ? mean[c] : 0.f;
    const float inv = MODE == 2 ? rsqrtf(var[c] + eps) : 0.f;
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + tid; i < total; i += (int64_t)gridDim.x * 256) {
        const float x = X[i];
        if (MODE == 0) a0 += (double)x;
        else if (MODE == 1) { const float d = x - mu; a0 += (double)d * (double)d; }
        else { const float g = dY[i]; a0 += (double)g; a1 += (double)g * (double)((x - mu) * inv); }
    }
    s0[tid] = a0; s1[tid] = a1;
    __syncthreads();
    if (tid < C) {
        double t0 = 0.0, t1 = 0.0;
        for (int j = tid; j < 256; j += C) { t0 += s0[j]; t1 += s1[j]; }
        atomicAdd(&out[tid], t0);
        if (MODE == 2) atomicAdd(&out[C + tid], t1);
    }
}
__global__ void bn_finish_stat_kernel(const double* __restrict__ acc, float* __restrict__ dst, int C, double inv_rows) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) dst[c] = (float)(acc[c] * inv_rows);
}
__global__ void bn_apply_kernel(const float* __restrict__ X, const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ mean, const float* __restrict__ var, float eps, float* __restrict__ Y,
                                int64_t total, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i & (C - 1));
    Y[i] = fmaf(gamma[c] * rsqrtf(var[c] + eps), X[i] - mean[c], beta[c]);
}
// dx = gamma inv / N * (N dy - sum(dy) - xhat sum(dy xhat))
__global__ void bn_bwd_apply_kernel(const float* __restrict__ X, const float* __restrict__ dY, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                    const double* __restrict__ sums, float* __restrict__ dX, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int64_t rows, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) { dbeta[i] = (float)sums[i]; dgamma[i] = (float)sums[C + i]; }
    if (i >= rows * C) return;
    const int c = (int)(i & (C - 1));
    const float inv = rsqrtf(var[c] + eps);
    const float xhat = (X[i] - mean[c]) * inv;
    const float sd = (float)sums[c], sdx = (float)sums[C + c];
    const float n = (float)rows;
    dX[i] = gamma[c] * inv / n * (n * dY[i] - sd - xhat * sdx);
}

// ---------------------------------------------------------------------------------------------------------------------------
// LSTM cell
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float hard_sigmoid(float x) { return fminf(fmaxf(fmaf(0.2f, x, 0.5f), 0.f), 1.f); }

__global__ void lstm_cell_fwd_kernel(float* __restrict__ z, const float* __restrict__ c_prev, float* __restrict__ c,
                                     float* __restrict__ h, int ldh, int B, int u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * u) return;
    const int b = i / u, j = i - b * u;
    float* zr = z + (size_t)b * 4 * u;
    const float ig = hard_sigmoid(zr[j]), fg = hard_sigmoid(zr[u + j]), gg = tanhf(zr[2 * u + j]), og = hard_sigmoid(zr[3 * u + j]);
    const float cp = c_prev ? c_prev[i] : 0.f;
    const float cn = fmaf(fg, cp, ig * gg);
    zr[j] = ig; zr[u + j] = fg; zr[2 * u + j] = gg; zr[3 * u + j] = og;
    c[i] = cn;
    h[(size_t)b * ldh + j] = og * tanhf(cn);
}

__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c,
                                     const float* __restrict__ dh_out, int ldh, const float* __restrict__ dh_rec,
                                     float* __restrict__ dc, float* __restrict__ dz, int B, int u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * u) return;
    const int b = i / u, j = i - b * u;
    const float* g = gates + (size_t)b * 4 * u;
    const float ig = g[j], fg = g[u + j], gg = g[2 * u + j], og = g[3 * u + j];
    const float dh = dh_out[(size_t)b * ldh + j] + (dh_rec ? dh_rec[i] : 0.f);
    const float tc = tanhf(c[i]);
    const float dcn = dc[i] + dh * og * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[i] : 0.f;
    // hard_sigmoid' = 0.2 strictly inside the linear range (the activated value is strictly between 0 and 1), else 0
    auto hs = [](float a) { return (a > 0.f && a < 1.f) ? 0.2f : 0.f; };
    float* d = dz + (size_t)b * 4 * u;
    d[j] = dcn * gg * hs(ig);
    d[u + j] = dcn * cp * hs(fg);
    d[2 * u + j] = dcn * ig * (1.f - gg * gg);
    d[3 * u + j] = dh * tc * hs(og);
    dc[i] = dcn * fg;
}

// ---------------------------------------------------------------------------------------------------------------------------
// losses, dropout, optimiser
// ---------------------------------------------------------------------------------------------------------------------------
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, const float* __restrict__ cw,
                                  float* __restrict__ probs, float* __restrict__ dlogits, float* __restrict__ stats, int B, int nc,
                                  float scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float loss = 0.f, hit = 0.f;
    if (i < B) {
        const float* l = logits + (size_t)i * nc;
        float mx = l[0];
        int am = 0;
        for (int k = 1; k < nc; ++k) if (l[k] > mx) { mx = l[k]; am = k; }
        float e[8], s = 0.f;
        for (int k = 0; k < nc; ++k) { e[k] = expf(l[k] - mx); s += e[k]; }
        const int y = labels[i];
        const float w = cw ? cw[y] : 1.f;
        for (int k = 0; k < nc; ++k) {
            const float p = e[k] / s;
            probs[(size_t)i * nc + k] = p;
            dlogits[(size_t)i * nc + k] = scale * w * (p - (k == y ? 1.f : 0.f));
        }
        // Keras clips the probability to [1e-7, 1 - 1e-7] before the log (backend.sparse_categorical_crossentropy)
        loss = -w * logf(fminf(fmaxf(e[y] / s, 1e-7f), 1.f - 1e-7f));
        hit = am == y ? 1.f : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { loss += __shfl_xor_sync(0xffffffffu, loss, o); hit += __shfl_xor_sync(0xffffffffu, hit, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[0], loss); atomicAdd(&stats[1], hit); }
}

__global__ void center_loss_kernel(const float* __restrict__ feat, const int32_t* __restrict__ labels, const float* __restrict__ centers,
                                   float* __restrict__ dfeat, float* __restrict__ dcenters, float* __restrict__ stats, int B, int dim,
                                   float scale, int centers_elsewhere) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l2 = 0.f;
    if (i < B * dim) {
        const int b = i / dim, k = i - b * dim;
        const int y = labels[b];
        const float d = feat[i] - centers[y * dim + k];
        l2 = d * d;
        const float g = scale * 2.f * d;
        dfeat[i] += g;
        if (!centers_elsewhere) atomicAdd(&dcenters[y * dim + k], -g);
    }
    // per-sample sum over the dim (= 16: half a warp) lanes of a sample, then Keras' "accuracy" of this output against its all-zero
    // target: binary_accuracy = mean(round(l2_i) == 0), round half to even
    float per = l2;
    if (dim == 16) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) per += __shfl_xor_sync(0xffffffffu, per, o);
    }
    float hit = (dim == 16 && (threadIdx.x & 15) == 0 && i < B * dim && per <= 0.5f) ? 1.f : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { l2 += __shfl_xor_sync(0xffffffffu, l2, o); hit += __shfl_xor_sync(0xffffffffu, hit, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&stats[2], l2); atomicAdd(&stats[3], hit); }
}

// deterministic form of the centre gradient: block c, thread k sums the samples of class c in batch order (one writer per entry)
__global__ void center_grad_det_kernel(const float* __restrict__ feat, const int32_t* __restrict__ labels, const float* __restrict__ centers,
                                       float* __restrict__ dcenters, int B, int dim, float scale) {
    const int c = blockIdx.x, k = threadIdx.x;
    if (k >= dim) return;
    float acc = 0.f;
    bool any = false;
    for (int b = 0; b < B; ++b)
        if (labels[b] == c) { any = true; acc -= scale * 2.f * (feat[b * dim + k] - centers[c * dim + k]); }
    if (any) dcenters[c * dim + k] += acc;
}

// keep-mask of Dropout(rate): a counter-based generator (the splitmix64 finaliser over (seed, step, element)), so that a step's
// mask depends on nothing but those three numbers
__global__ void dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t n, uint64_t seed, const int64_t* __restrict__ step_p, float rate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t step = (uint64_t)*step_p;
    uint64_t x = seed ^ (step * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)i * 0xD1B54A32D192ED03ull);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    const float r = (float)(x >> 40) * (1.0f / 16777216.0f);          // 24 bits -> [0, 1)
    mask[i] = r >= rate ? 1 : 0;
}

__global__ void dropout_kernel(float* __restrict__ X, const uint8_t* __restrict__ mask, int64_t n, float scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) X[i] = mask[i] ? X[i] * scale : 0.f;
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            const float* __restrict__ lr_t_p, float b1, float b2, float eps) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lr_t = *lr_t_p;
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void ema_kernel(float* __restrict__ moving, const float* __restrict__ batch, int64_t n, float momentum, float batch_scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) moving[i] = momentum * moving[i] + (1.f - momentum) * batch[i] * batch_scale;
}

bool pow2_le_256(int C) { return C >= 1 && C <= 256 && (C & (C - 1)) == 0; }

// NRV_TRAIN_DETERMINISTIC=1: no fp32 atomics (split-K products, row-chunked column sums), so that a step does the same float
// additions in the same order on every run; ~2x slower.  (The double-precision atomics of the batch-norm / conv-weight reductions
// stay: their order changes bits far below the float the sums are rounded to.)
bool deterministic() {
    static const bool d = getenv("NRV_TRAIN_DETERMINISTIC") && atoi(getenv("NRV_TRAIN_DETERMINISTIC")) != 0;
    return d;
}

}  // namespace

extern "C" {

const char* nrvt_last_error(void) { return g_err.c_str(); }

int nrvt_gemm(void* stream, int ta, int tb, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
              float beta, float* C, int ldc) {
    if (M <= 0 || N <= 0) return 0;
    if (K < 0 || !A || !B || !C) return bad("nrvt_gemm: bad arguments");
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    // few output tiles for the K at hand (the weight gradients: K = all rows of the batch; the per-timestep recurrent products:
    // 16 tiles): split K over the grid's z dimension so that the launch fills the SMs
    const int tiles = (int)(grid.x * grid.y);
    const int splits = std::min(K / 64, 592 / std::max(tiles, 1));
    const bool splitk = tiles < 74 && splits >= 2 && !deterministic();
    int kchunk = 0;
    if (splitk) {
        kchunk = (((K + splits - 1) / splits) + 15) / 16 * 16;
        grid.z = (K + kchunk - 1) / kchunk;
        scale2d_kernel<<<blocks_for((int64_t)M * N, 256), 256, 0, S(stream)>>>(C, M, N, ldc, beta);
    }
#define NRVT_GEMM(TA_, TB_)                                                                                                         \
    do {                                                                                                                            \
        if (splitk) gemm_kernel<TA_, TB_, true><<<grid, 256, 0, S(stream)>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, kchunk); \
        else gemm_kernel<TA_, TB_, false><<<grid, 256, 0, S(stream)>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, 0);            \
    } while (0)
    if (!ta && !tb) NRVT_GEMM(false, false);
    else if (!ta && tb) NRVT_GEMM(false, true);
    else if (ta && !tb) NRVT_GEMM(true, false);
    else NRVT_GEMM(true, true);
#undef NRVT_GEMM
    LAUNCH_CHECK("nrvt_gemm");
    return 0;
}

int nrvt_bias_act(void* stream, float* Y, int M, int N, int ldy, const float* bias, int relu) {
    if ((int64_t)M * N <= 0) return 0;
    bias_act_kernel<<<blocks_for((int64_t)M * N, 256), 256, 0, S(stream)>>>(Y, M, N, ldy, bias, relu);
    LAUNCH_CHECK("nrvt_bias_act");
    return 0;
}

int nrvt_relu_bwd(void* stream, float* dY, const float* Y, int64_t n) {
    if (n <= 0) return 0;
    relu_bwd_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(dY, Y, n);
    LAUNCH_CHECK("nrvt_relu_bwd");
    return 0;
}

int nrvt_colsum(void* stream, const float* X, int M, int N, int ldx, float* out, float beta) {
    if (N <= 0) return 0;
    const int rchunk = deterministic() ? std::max(M, 1) : 512;
    scale2d_kernel<<<blocks_for(N, 256), 256, 0, S(stream)>>>(out, 1, N, N, beta);
    colsum_kernel<<<dim3((N + 31) / 32, (std::max(M, 1) + rchunk - 1) / rchunk), 256, 0, S(stream)>>>(X, M, N, ldx, out, rchunk);
    LAUNCH_CHECK("nrvt_colsum");
    return 0;
}

int nrvt_copy2d(void* stream, float* dst, int ldd, const float* src, int lds, int rows, int cols, int accumulate) {
    if ((int64_t)rows * cols <= 0) return 0;
    copy2d_kernel<<<blocks_for((int64_t)rows * cols, 256), 256, 0, S(stream)>>>(dst, ldd, src, lds, rows, cols, accumulate);
    LAUNCH_CHECK("nrvt_copy2d");
    return 0;
}

int nrvt_conv1d_fwd(void* stream, const float* X, const float* W, const float* b, float* Y, int n, int L, int cin, int cout) {
    if (cin < 1 || cin > 8 || cout < 1 || cout > 8 || L < 1) return bad("nrvt_conv1d_fwd: cin, cout must be 1..8");
    const int64_t rows = (int64_t)n * L;
    if (rows <= 0) return 0;
    conv1d_fwd_kernel<<<blocks_for(rows, 256), 256, 0, S(stream)>>>(X, W, b, Y, rows, L, cin, cout);
    LAUNCH_CHECK("nrvt_conv1d_fwd");
    return 0;
}

int nrvt_conv1d_bwd(void* stream, const float* X, const float* W, const float* Y, const float* dY, float* dX, float* dW, float* db,
                    double* work, int n, int L, int cin, int cout) {
    if (cin < 1 || cin > 8 || cout < 1 || cout > 8 || L < 1 || L > 1024) return bad("nrvt_conv1d_bwd: cin, cout must be 1..8, L <= 1024");
    const int64_t rows = (int64_t)n * L;
    if (rows <= 0) return 0;
    if (dX) {
        conv1d_bwd_data_kernel<<<blocks_for(rows, 256), 256, 0, S(stream)>>>(W, Y, dY, dX, rows, L, cin, cout);
        LAUNCH_CHECK("nrvt_conv1d_bwd (data)");
    }
    // weight / bias gradients accumulate in the caller's double scratch
    if (!work) return bad("nrvt_conv1d_bwd: work is NULL");
    const int nacc = 3 * cin * cout + cout;
    cudaMemsetAsync(work, 0, nacc * sizeof(double), S(stream));
    const size_t smem = ((size_t)(L + 2) * cin + (size_t)L * cout) * sizeof(float);
    conv1d_bwd_w_kernel<<<(unsigned)std::min<int64_t>(n, 1184), 256, smem, S(stream)>>>(X, Y, dY, work, n, L, cin, cout);
    double_to_float_kernel<<<1, 256, 0, S(stream)>>>(work, dW, 3 * cin * cout, db, cout);
    LAUNCH_CHECK("nrvt_conv1d_bwd (weights)");
    return 0;
}

int nrvt_add_bcast(void* stream, float* Y, const float* X, int64_t rows, int C) {
    if (rows * C <= 0) return 0;
    add_bcast_kernel<<<blocks_for(rows * C, 256), 256, 0, S(stream)>>>(Y, X, rows, C);
    LAUNCH_CHECK("nrvt_add_bcast");
    return 0;
}

int nrvt_bn_fwd(void* stream, const float* X, const float* gamma, const float* beta, float eps, float* Y, float* mean, float* var,
                double* work, int64_t rows, int C) {
    if (!pow2_le_256(C)) return bad("nrvt_bn_fwd: C must be a power of two <= 256");
    if (rows <= 0) return 0;
    const int64_t total = rows * C;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 1184);
    cudaMemsetAsync(work, 0, 2 * C * sizeof(double), S(stream));
    bn_reduce_kernel<0><<<grid, 256, 0, S(stream)>>>(X, nullptr, nullptr, nullptr, eps, work, rows, C);
    bn_finish_stat_kernel<<<1, 256, 0, S(stream)>>>(work, mean, C, 1.0 / (double)rows);
    bn_reduce_kernel<1><<<grid, 256, 0, S(stream)>>>(X, nullptr, mean, nullptr, eps, work + C, rows, C);
    bn_finish_stat_kernel<<<1, 256, 0, S(stream)>>>(work + C, var, C, 1.0 / (double)rows);
    bn_apply_kernel<<<blocks_for(total, 256), 256, 0, S(stream)>>>(X, gamma, beta, mean, var, eps, Y, total, C);
    LAUNCH_CHECK("nrvt_bn_fwd");
    return 0;
}

int nrvt_bn_apply(void* stream, const float* X, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                  float* Y, int64_t rows, int C) {
    if (!pow2_le_256(C)) return bad("nrvt_bn_apply: C must be a power of two <= 256");
    if (rows <= 0) return 0;
    bn_apply_kernel<<<blocks_for(rows * C, 256), 256, 0, S(stream)>>>(X, gamma, beta, mean, var, eps, Y, rows * C, C);
    LAUNCH_CHECK("nrvt_bn_apply");
    return 0;
}

int nrvt_bn_bwd(void* stream, const float* X, const float* dY, const float* gamma, const float* mean, const float* var, float eps,
                float* dX, float* dgamma, float* dbeta, double* work, int64_t rows, int C) {
    if (!pow2_le_256(C)) return bad("nrvt_bn_bwd: C must be a power of two <= 256");
    if (rows <= 0) return 0;
    const int64_t total = rows * C;
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 1184);
    cudaMemsetAsync(work, 0, 2 * C * sizeof(double), S(stream));
    bn_reduce_kernel<2><<<grid, 256, 0, S(stream)>>>(X, dY, mean, var, eps, work, rows, C);
    bn_bwd_apply_kernel<<<blocks_for(std::max<int64_t>(total, C), 256), 256, 0, S(stream)>>>(X, dY, gamma, mean, var, eps, work, dX, dgamma,
                                                                                           dbeta, rows, C);
    LAUNCH_CHECK("nrvt_bn_bwd");
    return 0;
}

int nrvt_lstm_cell_fwd(void* stream, float* z, const float* c_prev, float* c, float* h, int ldh, int B, int u) {
    if (B * u <= 0) return 0;
    lstm_cell_fwd_kernel<<<blocks_for((int64_t)B * u, 256), 256, 0, S(stream)>>>(z, c_prev, c, h, ldh, B, u);
    LAUNCH_CHECK("nrvt_lstm_cell_fwd");
    return 0;
}

int nrvt_lstm_cell_bwd(void* stream, const float* gates, const float* c_prev, const float* c, const float* dh_out, int ldh,
                       const float* dh_rec, float* dc, float* dz, int B, int u) {
    if (B * u <= 0) return 0;
    lstm_cell_bwd_kernel<<<blocks_for((int64_t)B * u, 256), 256, 0, S(stream)>>>(gates, c_prev, c, dh_out, ldh, dh_rec, dc, dz, B, u);
    LAUNCH_CHECK("nrvt_lstm_cell_bwd");
    return 0;
}

int nrvt_softmax_ce(void* stream, const float* logits, const int32_t* labels, const float* cw, float* probs, float* dlogits,
                    float* stats, int B, int nc, float scale) {
    if (nc < 1 || nc > 8) return bad("nrvt_softmax_ce: 1 <= classes <= 8");
    if (B <= 0) return 0;
    softmax_ce_kernel<<<blocks_for(B, 128), 128, 0, S(stream)>>>(logits, labels, cw, probs, dlogits, stats, B, nc, scale);
    LAUNCH_CHECK("nrvt_softmax_ce");
    return 0;
}

int nrvt_center_loss(void* stream, const float* feat, const int32_t* labels, const float* centers, float* dfeat, float* dcenters,
                     float* stats, int B, int dim, float scale) {
    if (B * dim <= 0) return 0;
    const int det = deterministic() ? 1 : 0;
    center_loss_kernel<<<blocks_for((int64_t)B * dim, 128), 128, 0, S(stream)>>>(feat, labels, centers, dfeat, dcenters, stats, B, dim, scale, det);
    if (det && dim <= 1024) center_grad_det_kernel<<<8, dim, 0, S(stream)>>>(feat, labels, centers, dcenters, B, dim, scale);   // classes < 8 (nrvt_softmax_ce)
    LAUNCH_CHECK("nrvt_center_loss");
    return 0;
}

int nrvt_dropout_mask(void* stream, uint8_t* mask, int64_t n, uint64_t seed, const int64_t* step, float rate) {
    if (n <= 0) return 0;
    dropout_mask_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(mask, n, seed, step, rate);
    LAUNCH_CHECK("nrvt_dropout_mask");
    return 0;
}

int nrvt_dropout(void* stream, float* X, const uint8_t* mask, int64_t n, float scale) {
    if (n <= 0) return 0;
    dropout_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(X, mask, n, scale);
    LAUNCH_CHECK("nrvt_dropout");
    return 0;
}

int nrvt_adam(void* stream, float* p, const float* g, float* m, float* v, int64_t n, const float* lr_t, float b1, float b2, float eps) {
    if (n <= 0) return 0;
    adam_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(p, g, m, v, n, lr_t, b1, b2, eps);
    LAUNCH_CHECK("nrvt_adam");
    return 0;
}

int nrvt_ema(void* stream, float* moving, const float* batch, int64_t n, float momentum, float batch_scale) {
    if (n <= 0) return 0;
    ema_kernel<<<blocks_for(n, 256), 256, 0, S(stream)>>>(moving, batch, n, momentum, batch_scale);
    LAUNCH_CHECK("nrvt_ema");
    return 0;
}

}  // extern "C"
