// K2: the per-base signal branch (nanorevcnn.py:17-38, lstmmodel.py:35-41):
//   window gather (preprocessing.py:111-131, fused: windows are never materialised in HBM)
//   -> Conv1D(1->8,k3,same,relu) -> BN -> Conv1D(8->8,k3,same,relu) -> BN -> Add(input broadcast)
//   -> Flatten (index = pos*8 + ch) -> Dense(400->64, linear).
// The reference runs this inside TimeDistributed for each of the W timesteps of every window,
// i.e. 11x redundantly; here it is computed once per base and the LSTM kernels index it by
// (first base of window + t).
#include "nrv_common.cuh"

namespace nrv {

constexpr int CNN_TB = 32;          // bases per CTA
constexpr int CNN_THREADS = 256;
constexpr int WIN_LD = 52;          // 50 + zero halo on both sides ('same' padding)
constexpr int FLAT_LD = 404;        // 400 + 4: rows 2 apart land in different banks

struct CnnSmem {
    float win[CNN_TB][WIN_LD];
    float c1[CNN_TB][WIN_LD][NRV_CNN_CH];
    float flat[CNN_TB][FLAT_LD];
    float w[264];
};

__global__ void __launch_bounds__(CNN_THREADS, 2)
cnn_kernel(CnnDev W, const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
           const int32_t* __restrict__ starts, const int32_t* __restrict__ base_read,
           const double* __restrict__ shift, const double* __restrict__ scale,
           const float* __restrict__ explicit_win, int64_t n_bases, float* __restrict__ sig_feat,
           __half* __restrict__ sf_hi, __half* __restrict__ sf_lo) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CnnSmem& s = *reinterpret_cast<CnnSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * CNN_TB;

    for (int i = tid; i < 264; i += CNN_THREADS) s.w[i] = W.blob[i];
    // zero halos of win and c1
    for (int i = tid; i < CNN_TB * 2; i += CNN_THREADS) {
        const int b = i >> 1, e = (i & 1) ? WIN_LD - 1 : 0;
        s.win[b][e] = 0.f;
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) s.c1[b][e][c] = 0.f;
    }
    // ---- stage A: gather + normalise + symmetric zero pad ----------------------------------
    for (int i = tid; i < CNN_TB * NRV_SIG; i += CNN_THREADS) {
        const int b = i / NRV_SIG, p = i - b * NRV_SIG;
        const int64_t j = j0 + b;
        float v = 0.f;
        if (j < n_bases) {
            if (explicit_win) {
                v = explicit_win[j * NRV_SIG + p];
            } else {
                const int r = base_read[j];
                const int16_t* sig = signal + sig_off[r];
                const long long S = sig_off[r + 1] - sig_off[r];
                const long long st = starts[j];
                const long long lo = (st - 25 <= 0) ? 0 : st - 25;
                long long hi = (st + 25 >= S) ? S : st + 25;
                if (hi < lo) hi = lo;
                const int len = (int)(hi - lo);
                const int left = (NRV_SIG - len + 1) / 2;
                if (p >= left && p < left + len)
                    v = (float)(((double)sig[lo + (p - left)] - shift[r]) / scale[r]);
            }
        }
        s.win[b][p + 1] = v;
    }
    __syncthreads();
    const float* w1 = s.w;            // [3][8]
    const float* b1 = s.w + 24;
    const float* s1 = s.w + 32;
    const float* t1 = s.w + 40;
    const float* w2 = s.w + 48;       // [3][8][8] (k, cin, cout)
    const float* b2 = s.w + 240;
    const float* s2 = s.w + 248;
    const float* t2 = s.w + 256;
    // ---- stage B: conv1 + relu + BN1 ---------------------------------------------------------
    for (int i = tid; i < CNN_TB * NRV_SIG; i += CNN_THREADS) {
        const int b = i / NRV_SIG, p = i - b * NRV_SIG;
        const float x0 = s.win[b][p], x1 = s.win[b][p + 1], x2 = s.win[b][p + 2];
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) {
            float a = b1[c];
            a = fmaf(x0, w1[c], a);
            a = fmaf(x1, w1[8 + c], a);
            a = fmaf(x2, w1[16 + c], a);
            a = fmaxf(a, 0.f);
            s.c1[b][p + 1][c] = fmaf(a, s1[c], t1[c]);
        }
    }
    __syncthreads();
    // ---- stage C: conv2 + relu + BN2 + Add(input) -> flat[b][pos*8 + ch] -----------------------
    for (int i = tid; i < CNN_TB * NRV_SIG; i += CNN_THREADS) {
        const int b = i / NRV_SIG, p = i - b * NRV_SIG;
        float acc[NRV_CNN_CH];
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) acc[c] = b2[c];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int ci = 0; ci < NRV_CNN_CH; ++ci) {
                const float xv = s.c1[b][p + k][ci];
#pragma unroll
                for (int co = 0; co < NRV_CNN_CH; ++co) acc[co] = fmaf(xv, w2[(k * 8 + ci) * 8 + co], acc[co]);
            }
        }
        const float xin = s.win[b][p + 1];
#pragma unroll
        for (int c = 0; c < NRV_CNN_CH; ++c) {
            float a = fmaxf(acc[c], 0.f);
            s.flat[b][p * 8 + c] = fmaf(a, s2[c], t2[c]) + xin;
        }
    }
    __syncthreads();
    // ---- stage D: Dense(400 -> 64) -----------------------------------------------------------
    {
        const int tx = tid & 15, ty = tid >> 4;          // 16 column groups x 16 row pairs
        const int r0 = ty * 2;
        float acc0[4], acc1[4];
        const float4 bb = __ldg(reinterpret_cast<const float4*>(W.dense_b) + tx);
        acc0[0] = acc1[0] = bb.x; acc0[1] = acc1[1] = bb.y; acc0[2] = acc1[2] = bb.z; acc0[3] = acc1[3] = bb.w;
        const float4* Wd = reinterpret_cast<const float4*>(W.dense_k) + tx;
#pragma unroll 8
        for (int k = 0; k < 400; ++k) {
            const float4 w = __ldg(Wd + k * 16);
            const float a0 = s.flat[r0][k], a1 = s.flat[r0 + 1][k];
            acc0[0] = fmaf(a0, w.x, acc0[0]); acc0[1] = fmaf(a0, w.y, acc0[1]);
            acc0[2] = fmaf(a0, w.z, acc0[2]); acc0[3] = fmaf(a0, w.w, acc0[3]);
            acc1[0] = fmaf(a1, w.x, acc1[0]); acc1[1] = fmaf(a1, w.y, acc1[1]);
            acc1[2] = fmaf(a1, w.z, acc1[2]); acc1[3] = fmaf(a1, w.w, acc1[3]);
        }
        const int64_t ja = j0 + r0, jb = ja + 1;
        if (ja < n_bases)
            reinterpret_cast<float4*>(sig_feat + ja * NRV_SIGFEAT)[tx] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
        if (jb < n_bases)
            reinterpret_cast<float4*>(sig_feat + jb * NRV_SIGFEAT)[tx] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
        if (sf_hi) {   // fp16 (hi, lo) copy: A operand of the tensor-core projection of total_rnn1's per-base part
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (ja < n_bases) {
                    const __half h = __float2half_rn(acc0[c]);
                    sf_hi[ja * NRV_SIGFEAT + tx * 4 + c] = h;
                    sf_lo[ja * NRV_SIGFEAT + tx * 4 + c] = __float2half_rn(acc0[c] - __half2float(h));
                }
                if (jb < n_bases) {
                    const __half h = __float2half_rn(acc1[c]);
                    sf_hi[jb * NRV_SIGFEAT + tx * 4 + c] = h;
                    sf_lo[jb * NRV_SIGFEAT + tx * 4 + c] = __float2half_rn(acc1[c] - __half2float(h));
                }
            }
        }
    }
}

int launch_cnn(const ModelDev* m1, const ModelDev* m2, const int16_t* signal, const int64_t* sig_off,
               const int32_t* starts, const int64_t* /*base_off*/, const int32_t* base_read,
               const double* shift, const double* scale, const float* explicit_win, int64_t n_bases,
               float* sig_feat1, float* sig_feat2, __half* const sf_hi[2], __half* const sf_lo[2], cudaStream_t st) {
    if (n_bases <= 0) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(cnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CnnSmem));
        attr_set = true;
    }
    const unsigned grid = (unsigned)((n_bases + CNN_TB - 1) / CNN_TB);
    int n = 0;
    if (m1 && sig_feat1) {
        cnn_kernel<<<grid, CNN_THREADS, sizeof(CnnSmem), st>>>(m1->cnn, signal, sig_off, starts, base_read, shift,
                                                                scale, explicit_win, n_bases, sig_feat1, sf_hi ? sf_hi[0] : nullptr,
                                                                sf_lo ? sf_lo[0] : nullptr);
        ++n;
    }
    if (m2 && sig_feat2) {
        cnn_kernel<<<grid, CNN_THREADS, sizeof(CnnSmem), st>>>(m2->cnn, signal, sig_off, starts, base_read, shift,
                                                                scale, explicit_win, n_bases, sig_feat2, sf_hi ? sf_hi[1] : nullptr,
                                                                sf_lo ? sf_lo[1] : nullptr);
        ++n;
    }
    return n;
}

}  // namespace nrv
