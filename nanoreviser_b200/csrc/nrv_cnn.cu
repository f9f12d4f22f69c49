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

constexpr int C1_ROWS = CNN_TB * WIN_LD;      // conv1 output rows (base, padded position), 8 channels each
struct CnnSmem {
    float win[CNN_TB][WIN_LD];
    // conv1 output as an fp16 (hi, lo) pair, row R = b * 52 + padded position stored at [R + 1] (16 bytes per row; one zero row in
    // front and behind, zero halo rows inside).  The im2col row of output (b, p) for the k = 3 convolution -- taps (p-1, p, p+1) x 8
    // channels -- is then the 48 CONTIGUOUS bytes starting at stored row R: conv2 is a [1664 x 24] x [24 x 8] GEMM whose A operand
    // ldmatrix reads straight from this array with overlapping rows (row stride 16 B).
    __half c1h[C1_ROWS + 2][NRV_CNN_CH];
    __half c1l[C1_ROWS + 2][NRV_CNN_CH];
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

__device__ __forceinline__ void ldmatrix_x2(uint32_t saddr, uint32_t (&r)[2]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(saddr));
}
// D(16x8, fp32) += A(16x8, fp16, row) . B(8x8, fp16, col)
__device__ __forceinline__ void mma_1688(float (&d)[4], const uint32_t (&a)[2], uint32_t b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b));
}

__device__ __forceinline__ float clamp_f16(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}

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

    // zero halos of win; zero rows of the conv1 array (pad rows, halo positions 0 and 51 of every base): never written afterwards
    for (int i = tid; i < CNN_TB * 2; i += CNN_THREADS) {
        const int b = i >> 1, e = (i & 1) ? WIN_LD - 1 : 0;
        s.win[b][e] = 0.f;
        *reinterpret_cast<uint4*>(&s.c1h[1 + b * WIN_LD + e][0]) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(&s.c1l[1 + b * WIN_LD + e][0]) = make_uint4(0, 0, 0, 0);
    }
    if (tid < 2) {
        *reinterpret_cast<uint4*>(&s.c1h[tid ? C1_ROWS + 1 : 0][0]) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(&s.c1l[tid ? C1_ROWS + 1 : 0][0]) = make_uint4(0, 0, 0, 0);
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
        uint32_t ph[4], pl[4];
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; c += 2) split2(clamp_f16(o[c]), clamp_f16(o[c + 1]), ph[c >> 1], pl[c >> 1]);
        *reinterpret_cast<uint4*>(&s.c1h[1 + b * WIN_LD + p + 1][0]) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4*>(&s.c1l[1 + b * WIN_LD + p + 1][0]) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
    }
    __syncthreads();
    // ---- stage C: conv2 + relu + BN2 + Add(input) -> flat[b][pos*8 + ch] as fp16 (hi, lo) ------
    // im2col GEMM on the tensor cores (mma.sync m16n8k16 for taps 0-1, m16n8k8 for tap 2; split-fp16 x 3, fp32 accumulate): 104 tiles
    // of 16 rows over the 1664 (base, padded position) rows, 13 per warp; halo rows are computed and dropped.
    {
        const int warp = tid >> 5, lane = tid & 31;
        // B fragments of this lane: B[k][n] = w2[k * 8 + n], k = tap * 8 + cin, n = cout = lane / 4
        uint32_t bh[3], bl[3];
        {
            const int n = lane >> 2, k0 = (lane & 3) * 2;
#pragma unroll
            for (int j = 0; j < 3; ++j) split2(w2[(j * 8 + k0) * 8 + n], w2[(j * 8 + k0 + 1) * 8 + n], bh[j], bl[j]);
        }
        const int co = (lane & 3) * 2;
        const float bias0 = b2[co], bias1 = b2[co + 1], sc0 = s2[co], sc1 = s2[co + 1], sh0 = t2[co], sh1 = t2[co + 1];
        const uint32_t c1h0 = (uint32_t)__cvta_generic_to_shared(&s.c1h[0][0]), c1l0 = (uint32_t)__cvta_generic_to_shared(&s.c1l[0][0]);
        const uint32_t arow = (uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8) * 16;
        const float* winflat = &s.win[0][0];
        for (int tile = warp; tile < C1_ROWS / 16; tile += CNN_THREADS / 32) {
            const int R0 = tile * 16;
            const uint32_t a16 = (uint32_t)R0 * 16 + arow + (uint32_t)(lane >> 4) * 16, a8 = (uint32_t)R0 * 16 + arow + 32;
            uint32_t ah[4], al[4], th[2], tl[2];
            ldmatrix_x4(c1h0 + a16, ah);
            ldmatrix_x4(c1l0 + a16, al);
            ldmatrix_x2(c1h0 + a8, th);
            ldmatrix_x2(c1l0 + a8, tl);
            float acc[4] = {bias0, bias1, bias0, bias1};
            mma_16816(acc, al, make_uint2(bh[0], bh[1]));
            mma_16816(acc, ah, make_uint2(bl[0], bl[1]));
            mma_16816(acc, ah, make_uint2(bh[0], bh[1]));
            mma_1688(acc, tl, bh[2]);
            mma_1688(acc, th, bl[2]);
            mma_1688(acc, th, bh[2]);
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int R = R0 + (lane >> 2) + hh * 8;
                const int bb = R / WIN_LD, pp = R - bb * WIN_LD;          // padded position 1..50 = position 0..49
                if (pp >= 1 && pp <= NRV_SIG) {
                    const float xin = winflat[R];
                    // clamped to the fp16 range: an outlier spike on a quiet read (tiny MAD) must saturate, not become inf - inf = NaN
                    const float y0 = clamp_f16(fmaf(fmaxf(acc[hh * 2], 0.f), sc0, sh0) + xin);
                    const float y1 = clamp_f16(fmaf(fmaxf(acc[hh * 2 + 1], 0.f), sc1, sh1) + xin);
                    uint32_t h2, l2;
                    split2(y0, y1, h2, l2);
                    *reinterpret_cast<uint32_t*>(&s.fh[bb][(pp - 1) * 8 + co]) = h2;
                    *reinterpret_cast<uint32_t*>(&s.fl[bb][(pp - 1) * 8 + co]) = l2;
                }
            }
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
