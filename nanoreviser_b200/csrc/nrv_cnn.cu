// K2: the per-base signal branch (nanorevcnn.py:17-38, lstmmodel.py:35-41):
//   window gather (preprocessing.py:111-131, fused: windows are never materialised in HBM)
//   -> Conv1D(1->8,k3,same,relu) -> BN -> Conv1D(8->8,k3,same,relu) -> BN -> Add(input broadcast)
//   -> Flatten (index = pos*8 + ch) -> Dense(400->64, linear).
// The reference runs this inside TimeDistributed for each of the W timesteps of every window,
// i.e. 11x redundantly; here it is computed once per base and the LSTM kernels index it by
// (first base of window + t).
#include <cuda_fp16.h>

#include "nrv_common.cuh"

namespace nrv {

constexpr int CNN_TB = 32;          // bases per CTA
constexpr int CNN_THREADS = 256;
constexpr int WIN_LD = 52;          // 50 + zero halo on both sides ('same' padding)
constexpr int FLAT_LDH = 408;       // halves per row of the flattened tile: 816 B rows -> conflict-free ldmatrix

struct CnnSmem {
    float win[CNN_TB][WIN_LD];
    float c1[CNN_TB][WIN_LD][NRV_CNN_CH];
    __half fh[CNN_TB][FLAT_LDH];    // flatten(conv stack) as an fp16 (hi, lo) pair: A operand of the tensor-core dense
    __half fl[CNN_TB][FLAT_LDH];
    float w[264];
    // per base of the tile (stage A): first sample of the window in the batch signal, samples available, left padding, shift, scale
    long long g_lo[CNN_TB];
    int g_len[CNN_TB], g_left[CNN_TB];
    float g_shift[CNN_TB], g_scale[CNN_TB];
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t saddr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// D(16x8, fp32) += A(16x16, fp16, row) . B(16x8, fp16, col)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], const uint2& b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

__device__ __forceinline__ float clamp_f16(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }

// One CTA = 32 bases, BOTH models: the window gather + normalisation (stage A) is shared, the conv stack and the dense run once per
// model on the same windows (the reference evaluates the two networks on identical inputs).
__global__ void __launch_bounds__(CNN_THREADS, 2)
cnn_kernel(CnnDev W0, CnnDev W1, int n_models, const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
           const int32_t* __restrict__ starts, const int32_t* __restrict__ base_read,
           const double* __restrict__ shift, const double* __restrict__ scale,
           const float* __restrict__ explicit_win, int64_t n_bases, float* __restrict__ sig_feat0, float* __restrict__ sig_feat1,
           __half* __restrict__ sf_hi0, __half* __restrict__ sf_lo0, __half* __restrict__ sf_hi1, __half* __restrict__ sf_lo1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CnnSmem& s = *reinterpret_cast<CnnSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * CNN_TB;

    // zero halos of win and c1
    for (int i = tid; i < CNN_TB * 2; i += CNN_THREADS) {
        const int b = i >> 1, e = (i & 1) ? WIN_LD - 1 : 0;
        s.win[b][e] = 0.f;
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) s.c1[b][e][c] = 0.f;
    }
    // ---- stage A: gather + normalise + symmetric zero pad ----------------------------------
    // (x - shift) / scale: the reference divides in fp64 and the network casts to fp32.  x is an int16, shift a multiple of 0.5 and
    // scale (the MAD, no 1.4826 factor) a multiple of 0.25, so numerator and denominator are exact in fp32 and ONE correctly rounded
    // fp32 division gives the same bits: the quotient of two such numbers cannot come within 2^-43 (relative) of an fp32 rounding
    // boundary without hitting it, so rounding the fp64 quotient (2^-53) to fp32 is the same as rounding the exact quotient.
    if (!explicit_win && tid < CNN_TB) {
        const int64_t j = j0 + tid;
        long long lo = 0; int len = 0, left = 0; float sh = 0.f, sc = 1.f;
        if (j < n_bases) {
            const int r = base_read[j];
            const long long o0 = sig_off[r], S = sig_off[r + 1] - o0;
            const long long st = starts[j];
            lo = (st - 25 <= 0) ? 0 : st - 25;
            long long hi = (st + 25 >= S) ? S : st + 25;
            if (hi < lo) hi = lo;
            len = (int)(hi - lo);
            left = (NRV_SIG - len + 1) / 2;
            lo += o0;
            sh = (float)shift[r]; sc = (float)scale[r];
        }
        s.g_lo[tid] = lo; s.g_len[tid] = len; s.g_left[tid] = left; s.g_shift[tid] = sh; s.g_scale[tid] = sc;
    }
    __syncthreads();
    for (int i = tid; i < CNN_TB * NRV_SIG; i += CNN_THREADS) {
        const int b = i / NRV_SIG, p = i - b * NRV_SIG;
        float v = 0.f;
        if (explicit_win) {
            if (j0 + b < n_bases) v = explicit_win[(j0 + b) * NRV_SIG + p];
        } else {
            const int q = p - s.g_left[b];
            if (q >= 0 && q < s.g_len[b]) v = __fdiv_rn((float)signal[s.g_lo[b] + q] - s.g_shift[b], s.g_scale[b]);
        }
        s.win[b][p + 1] = v;
    }
    const float* w1 = s.w;            // [3][8]
    const float* b1 = s.w + 24;
    const float* s1 = s.w + 32;
    const float* t1 = s.w + 40;
    const float* w2 = s.w + 48;       // [3][8][8] (k, cin, cout)
    const float* b2 = s.w + 240;
    const float* s2 = s.w + 248;
    const float* t2 = s.w + 256;
    for (int mi = 0; mi < n_models; ++mi) {
    const CnnDev& W = mi ? W1 : W0;
    float* const sig_feat = mi ? sig_feat1 : sig_feat0;
    __half* const sf_hi = mi ? sf_hi1 : sf_hi0;
    __half* const sf_lo = mi ? sf_lo1 : sf_lo0;
    __syncthreads();                  // windows written (mi = 0) / the previous model's dense has read its flatten tile
    for (int i = tid; i < 264; i += CNN_THREADS) s.w[i] = W.blob[i];
    __syncthreads();
    // ---- stage B: conv1 + relu + BN1 ---------------------------------------------------------
    for (int i = tid; i < CNN_TB * NRV_SIG; i += CNN_THREADS) {
        const int b = i / NRV_SIG, p = i - b * NRV_SIG;
        const float x0 = s.win[b][p], x1 = s.win[b][p + 1], x2 = s.win[b][p + 2];
        float o[NRV_CNN_CH];
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) {
            float a = b1[c];
            a = fmaf(x0, w1[c], a);
            a = fmaf(x1, w1[8 + c], a);
            a = fmaf(x2, w1[16 + c], a);
            a = fmaxf(a, 0.f);
            o[c] = fmaf(a, s1[c], t1[c]);
        }
        *reinterpret_cast<float4*>(&s.c1[b][p + 1][0]) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(&s.c1[b][p + 1][4]) = make_float4(o[4], o[5], o[6], o[7]);
    }
    __syncthreads();
    // ---- stage C: conv2 + relu + BN2 + Add(input) -> flat[b][pos*8 + ch] as fp16 (hi, lo) ------
    // Two positions per thread: every weight broadcast (LDS.128) feeds 8 FMAs and every input row is shared by the two
    // outputs it touches (one position per thread was shared-memory-bound: 1 LDS per FMA).
    for (int i = tid; i < CNN_TB * (NRV_SIG / 2); i += CNN_THREADS) {
        const int b = i / (NRV_SIG / 2), p0 = (i - b * (NRV_SIG / 2)) * 2;
        float in[4][NRV_CNN_CH];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float4 lo4 = *reinterpret_cast<const float4*>(&s.c1[b][p0 + r][0]);
            const float4 hi4 = *reinterpret_cast<const float4*>(&s.c1[b][p0 + r][4]);
            in[r][0] = lo4.x; in[r][1] = lo4.y; in[r][2] = lo4.z; in[r][3] = lo4.w;
            in[r][4] = hi4.x; in[r][5] = hi4.y; in[r][6] = hi4.z; in[r][7] = hi4.w;
        }
        float acc[2][NRV_CNN_CH];
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) acc[0][c] = acc[1][c] = b2[c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int ci = 0; ci < NRV_CNN_CH; ++ci) {
                const float4 wa = *reinterpret_cast<const float4*>(w2 + (k * 8 + ci) * 8);
                const float4 wb = *reinterpret_cast<const float4*>(w2 + (k * 8 + ci) * 8 + 4);
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int co = 0; co < NRV_CNN_CH; ++co) {
                    acc[0][co] = fmaf(in[k][ci], wv[co], acc[0][co]);
                    acc[1][co] = fmaf(in[k + 1][ci], wv[co], acc[1][co]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float xin = s.win[b][p0 + j + 1];
            uint32_t ph[4], pl[4];
#pragma unroll
            for (int c = 0; c < NRV_CNN_CH; c += 2) {
                // clamped to the fp16 range: an outlier spike on a quiet read (tiny MAD) must saturate, not become inf - inf = NaN
                const float y0 = clamp_f16(fmaf(fmaxf(acc[j][c], 0.f), s2[c], t2[c]) + xin);
                const float y1 = clamp_f16(fmaf(fmaxf(acc[j][c + 1], 0.f), s2[c + 1], t2[c + 1]) + xin);
                const __half2 h = __floats2half2_rn(y0, y1);
                const float2 f = __half22float2(h);
                const __half2 l = __floats2half2_rn(y0 - f.x, y1 - f.y);
                ph[c >> 1] = *reinterpret_cast<const uint32_t*>(&h); pl[c >> 1] = *reinterpret_cast<const uint32_t*>(&l);
            }
            *reinterpret_cast<uint4*>(&s.fh[b][(p0 + j) * 8]) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
            *reinterpret_cast<uint4*>(&s.fl[b][(p0 + j) * 8]) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        }
    }
    __syncthreads();
    // ---- stage D: Dense(400 -> 64) on the tensor cores (mma.sync m16n8k16, 3 split-fp16 passes, fp32 accumulate) ----
    // warp w owns output columns [8w, 8w + 8) for both 16-row halves of the tile; B fragments come pre-packed from L2.
    {
        const int warp = tid >> 5, lane = tid & 31;
        float acc[2][4];
        const float bias0 = __ldg(W.dense_b + warp * 8 + (lane & 3) * 2), bias1 = __ldg(W.dense_b + warp * 8 + (lane & 3) * 2 + 1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) { acc[mt][0] = acc[mt][2] = bias0; acc[mt][1] = acc[mt][3] = bias1; }
        const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;
        const uint32_t ah0 = (uint32_t)__cvta_generic_to_shared(&s.fh[lrow][lcol]);
        const uint32_t al0 = (uint32_t)__cvta_generic_to_shared(&s.fl[lrow][lcol]);
        const uint2* bh = W.dfrag_hi + warp * 32 + lane;
        const uint2* bl = W.dfrag_lo + warp * 32 + lane;
#pragma unroll 5
        for (int kt = 0; kt < 25; ++kt) {
            const uint2 b_hi = __ldg(bh + kt * 256), b_lo = __ldg(bl + kt * 256);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                uint32_t a_hi[4], a_lo[4];
                ldmatrix_x4(ah0 + (uint32_t)(mt * 16 * FLAT_LDH + kt * 16) * 2, a_hi);
                ldmatrix_x4(al0 + (uint32_t)(mt * 16 * FLAT_LDH + kt * 16) * 2, a_lo);
                mma_16816(acc[mt], a_lo, b_hi);
                mma_16816(acc[mt], a_hi, b_lo);
                mma_16816(acc[mt], a_hi, b_hi);
            }
        }
        const int col = warp * 8 + (lane & 3) * 2;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int64_t j = j0 + mt * 16 + hh * 8 + (lane >> 2);
                if (j < n_bases) {
                    const float y0 = acc[mt][hh * 2], y1 = acc[mt][hh * 2 + 1];
                    if (sig_feat) *reinterpret_cast<float2*>(sig_feat + j * NRV_SIGFEAT + col) = make_float2(y0, y1);
                    if (sf_hi) {   // fp16 (hi, lo) copy: A operand of the tensor-core projection of total_rnn1's per-base part
                        const float c0 = clamp_f16(y0), c1 = clamp_f16(y1);
                        const __half2 h = __floats2half2_rn(c0, c1);
                        const float2 f = __half22float2(h);
                        *reinterpret_cast<__half2*>(sf_hi + j * NRV_SIGFEAT + col) = h;
                        *reinterpret_cast<__half2*>(sf_lo + j * NRV_SIGFEAT + col) = __floats2half2_rn(c0 - f.x, c1 - f.y);
                    }
                }
            }
    }
    }       // models
}

int launch_cnn(const ModelDev* m1, const ModelDev* m2, const int16_t* signal, const int64_t* sig_off,
               const int32_t* starts, const int64_t* /*base_off*/, const int32_t* base_read,
               const double* shift, const double* scale, const float* explicit_win, int64_t n_bases,
               float* sig_feat1, float* sig_feat2, __half* const sf_hi[2], __half* const sf_lo[2], cudaStream_t st) {
    if (n_bases <= 0 || !m1) return 0;
    static PerDevice attr_set;
    if (attr_set.first()) cudaFuncSetAttribute(cnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CnnSmem));
    const unsigned grid = (unsigned)((n_bases + CNN_TB - 1) / CNN_TB);
    cnn_kernel<<<grid, CNN_THREADS, sizeof(CnnSmem), st>>>(m1->cnn, m2 ? m2->cnn : m1->cnn, m2 ? 2 : 1, signal, sig_off, starts, base_read, shift, scale,
                                                            explicit_win, n_bases, sig_feat1, sig_feat2, sf_hi ? sf_hi[0] : nullptr,
                                                            sf_lo ? sf_lo[0] : nullptr, sf_hi ? sf_hi[1] : nullptr, sf_lo ? sf_lo[1] : nullptr);
    return 1;
}

}  // namespace nrv
