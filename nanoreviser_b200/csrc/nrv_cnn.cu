// K2: the per-base signal branch (nanorevcnn.py:17-38, lstmmodel.py:35-41):
//   window gather (preprocessing.py:111-131, fused: windows are never materialised in HBM)
//   -> Conv1D(1->8,k3,same,relu) -> BN -> Conv1D(8->8,k3,same,relu) -> BN -> Add(input broadcast)
//   -> Flatten (index = pos*8 + ch) -> Dense(400->64, linear).
// conv2 is an im2col GEMM on mma.sync (N = 8 output channels fit the m16n8 shapes); the Dense is a tcgen05 GEMM with operands
// swapped (M = 64 output features, N = 32 bases of the tile, K = 400) so that it fits a 32-base tile: the flatten tile is written
// by conv2's epilogue directly in the swizzled K-major operand layout, the weights are streamed by TMA through a ring that reuses
// the shared memory of the conv1 output, the accumulator lives in 32 columns of tensor memory.
// The reference runs this inside TimeDistributed for each of the W timesteps of every window,
// i.e. 11x redundantly; here it is computed once per base and the LSTM kernels index it by
// (first base of window + t).
#include <cuda_fp16.h>

#include "nrv_common.cuh"
#include "nrv_tc.cuh"

namespace nrv {

using namespace tc;

bool make_tmap_f16_k64(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);        // nrv_gemm.cu
bool make_tmap_f16_k16_sw32(CUtensorMap* tm, const void* base, int64_t rows, int K, int box_rows);   // nrv_gemm.cu

constexpr int CNN_TB = 32;          // bases per CTA
constexpr int CNN_THREADS = 256;
constexpr int WIN_LD = 52;          // 50 + zero halo on both sides ('same' padding)

constexpr int C1_ROWS = CNN_TB * WIN_LD;      // conv1 output rows (base, padded position), 8 channels each
// flatten tile = B operand of the dense (N = 32 bases x K = 400, K-major): per part (hi / lo) six chunks of 64 columns
// ([32 rows][128 B], 128-byte swizzle) and one of 16 columns ([32 rows][32 B], 32-byte swizzle)
constexpr int FLAT_CHUNK = CNN_TB * 128;                  // 4 KB
constexpr int FLAT_TAIL_OFF = 6 * FLAT_CHUNK;             // the 16-column chunk
constexpr int FLAT_PART = FLAT_TAIL_OFF + CNN_TB * 32;    // 25,600 B
constexpr int DW_STAGES = 3;                              // ring of dense-weight chunks: [64 rows][128 B] hi at +0, lo at +8 KB
constexpr int DW_STAGE = 16384;
struct CnnSmem {
    alignas(1024) uint8_t flat[2 * FLAT_PART];
    // conv1 output as an fp16 (hi, lo) pair, row R = b * 52 + padded position stored at [R + 1] (16 bytes per row; one zero row in
    // front and behind, zero halo rows inside).  The im2col row of output (b, p) for the k = 3 convolution -- taps (p-1, p, p+1) x 8
    // channels -- is then the 48 CONTIGUOUS bytes starting at stored row R: conv2 is a [1664 x 24] x [24 x 8] GEMM whose A operand
    // ldmatrix reads straight from this array with overlapping rows (row stride 16 B).  Dead once conv2 is done: the dense-weight
    // ring of the tcgen05 stage lives in the same bytes.
    union {
        struct {
            __half c1h[C1_ROWS + 2][NRV_CNN_CH];
            __half c1l[C1_ROWS + 2][NRV_CNN_CH];
        };
        uint8_t wring[DW_STAGES * DW_STAGE];
    };
    float win[CNN_TB][WIN_LD];
    float w[264];
    // per base of the tile (stage A): first sample of the window in the batch signal, samples available, left padding, shift, scale
    long long g_lo[CNN_TB];
    int g_len[CNN_TB], g_left[CNN_TB];
    float g_shift[CNN_TB], g_scale[CNN_TB];
    uint64_t full[DW_STAGES], empty[DW_STAGES], acc_done;
    uint32_t tmem_slot;
};
static_assert(sizeof(((CnnSmem*)0)->wring) <= 2 * (C1_ROWS + 2) * NRV_CNN_CH * 2, "weight ring must fit the conv1 array");
static_assert(sizeof(CnnSmem) + 1024 <= 115200, "two CTAs per SM");

// byte offset, inside one part of the flatten tile, of channel pair (co, co + 1) of position p of base b
__device__ __forceinline__ uint32_t flat_offset(int b, int p, int co) {
    if (p < 48) return (uint32_t)(p >> 3) * FLAT_CHUNK + sw128_offset(b, p & 7) + (uint32_t)co * 2;
    return (uint32_t)FLAT_TAIL_OFF + (uint32_t)(b * 32 + (((p - 48) ^ ((b >> 2) & 1)) << 4)) + (uint32_t)co * 2;
}
// shared-memory descriptor of a K-major operand with rows of 32 bytes (K = 16 fp16) and the 32-byte swizzle (m64_probe.cu)
__device__ __forceinline__ uint64_t umma_desc_k_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;                     // stride between 8-row groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;                              // SWIZZLE_32B
    return d;
}
struct CnnTmaps { CUtensorMap w_hi, w_lo, t_hi, t_lo; };   // dense weights W^T [64][400]: 64-column boxes (hi, lo) and the 16-column tail

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
cnn_kernel(CnnDev W0, CnnDev W1, const __grid_constant__ CnnTmaps T0, const __grid_constant__ CnnTmaps T1, int n_models, const int16_t* __restrict__ signal, const int64_t* __restrict__ sig_off,
           const int32_t* __restrict__ starts, const int32_t* __restrict__ base_read,
           const double* __restrict__ shift, const double* __restrict__ scale,
           const float* __restrict__ explicit_win, int64_t n_bases, float* __restrict__ sig_feat0, float* __restrict__ sig_feat1,
           __half* __restrict__ sf_hi0, __half* __restrict__ sf_lo0, __half* __restrict__ sf_hi1, __half* __restrict__ sf_lo1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CnnSmem& s = *reinterpret_cast<CnnSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x;
    const int64_t j0 = (int64_t)blockIdx.x * CNN_TB;
    if (tid == 0) {
        for (int i = 0; i < DW_STAGES; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
        mbar_init(&s.acc_done, 1);
        fence_mbar_init();
        tma_prefetch_desc(&T0.w_hi); tma_prefetch_desc(&T0.w_lo); tma_prefetch_desc(&T0.t_hi); tma_prefetch_desc(&T0.t_lo);
    }
    if (tid < 32) tmem_alloc(&s.tmem_slot, 32);          // dense accumulator: 64 features (16 lanes per warp quarter) x 32 bases
    uint32_t ring_n = 0;                                  // dense-weight chunks streamed so far (both models): stage = n % 3, phase = n / 3

    // zero halos of win (the zero rows of the conv1 array are rewritten per model: the dense-weight ring overwrites them)
    for (int i = tid; i < CNN_TB * 2; i += CNN_THREADS) {
        const int b = i >> 1, e = (i & 1) ? WIN_LD - 1 : 0;
        s.win[b][e] = 0.f;
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
    // zero rows of the conv1 array: one in front, one behind, halo positions 0 and 51 of every base
    for (int i = tid; i < CNN_TB * 2 + 2; i += CNN_THREADS) {
        const int row = i < CNN_TB * 2 ? 1 + (i >> 1) * WIN_LD + ((i & 1) ? WIN_LD - 1 : 0) : (i == CNN_TB * 2 ? 0 : C1_ROWS + 1);
        *reinterpret_cast<uint4*>(&s.c1h[row][0]) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(&s.c1l[row][0]) = make_uint4(0, 0, 0, 0);
    }
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
                    const uint32_t fo = flat_offset(bb, pp - 1, co);
                    *reinterpret_cast<uint32_t*>(s.flat + fo) = h2;
                    *reinterpret_cast<uint32_t*>(s.flat + FLAT_PART + fo) = l2;
                }
            }
        }
    }
    // the flatten tile was written through the generic proxy and is read by the tensor core (async proxy); the conv1 array is about
    // to be overwritten by TMA
    fence_proxy_async_smem();
    __syncthreads();
    // ---- stage D: Dense(400 -> 64) on tcgen05: D^T[64 features][32 bases] = W^T . flat^T, split fp16 x 3, fp32 accumulate in TMEM ----
    {
        const int warp = tid >> 5;
        const CnnTmaps& TM = mi ? T1 : T0;
        const uint32_t tmem_acc = s.tmem_slot;
        if (warp == 0) {
            // TMA producer: 7 chunks of W^T (hi, lo) through the 3-stage ring
            if (elect_one()) {
                for (int c = 0; c < 7; ++c) {
                    const uint32_t n = ring_n + c, st = n % DW_STAGES, ph = (n / DW_STAGES) & 1;
                    mbar_wait(&s.empty[st], ph ^ 1);
                    uint8_t* dst = s.wring + st * DW_STAGE;
                    if (c < 6) {
                        mbar_arrive_expect_tx(&s.full[st], 2 * 64 * 128);
                        tma_load_2d(dst, &TM.w_hi, &s.full[st], c * 64, 0);
                        tma_load_2d(dst + 8192, &TM.w_lo, &s.full[st], c * 64, 0);
                    } else {
                        mbar_arrive_expect_tx(&s.full[st], 2 * 64 * 32);
                        tma_load_2d(dst, &TM.t_hi, &s.full[st], 384, 0);
                        tma_load_2d(dst + 8192, &TM.t_lo, &s.full[st], 384, 0);
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // MMA issue: M = 64 (features), N = 32 (bases); per K-step of 16: W_hi.f_lo + W_lo.f_hi + W_hi.f_hi
            constexpr uint32_t idesc = umma_idesc_f16_f32(64, CNN_TB);
            const uint32_t fbase = smem_u32(s.flat);
            for (int c = 0; c < 7; ++c) {
                const uint32_t n = ring_n + c, st = n % DW_STAGES, ph = (n / DW_STAGES) & 1;
                mbar_wait(&s.full[st], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t wa = smem_u32(s.wring + st * DW_STAGE);
                    if (c < 6) {
                        const uint32_t fb = fbase + c * FLAT_CHUNK;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t a_hi = umma_desc_k_sw128(wa + k * 32), a_lo = umma_desc_k_sw128(wa + 8192 + k * 32);
                            const uint64_t b_hi = umma_desc_k_sw128(fb + k * 32), b_lo = umma_desc_k_sw128(fb + FLAT_PART + k * 32);
                            umma_f16_ss(tmem_acc, a_hi, b_lo, idesc, (c | k) != 0);
                            umma_f16_ss(tmem_acc, a_lo, b_hi, idesc, 1);
                            umma_f16_ss(tmem_acc, a_hi, b_hi, idesc, 1);
                        }
                    } else {
                        const uint64_t a_hi = umma_desc_k_sw32(wa), a_lo = umma_desc_k_sw32(wa + 8192);
                        const uint64_t b_hi = umma_desc_k_sw32(fbase + FLAT_TAIL_OFF), b_lo = umma_desc_k_sw32(fbase + FLAT_PART + FLAT_TAIL_OFF);
                        umma_f16_ss(tmem_acc, a_hi, b_lo, idesc, 1);
                        umma_f16_ss(tmem_acc, a_lo, b_hi, idesc, 1);
                        umma_f16_ss(tmem_acc, a_hi, b_hi, idesc, 1);
                    }
                    umma_commit(&s.empty[st]);                 // the stage is free when these MMAs have read it
                    if (c == 6) umma_commit(&s.acc_done);
                }
                __syncwarp();
            }
        }
        ring_n += 7;
        mbar_wait(&s.acc_done, mi & 1);
        tc_fence_after();
        // epilogue: TMEM lane (warp % 4) * 32 + l holds feature (warp % 4) * 16 + l for l < 16 (m64_probe.cu); warps 0-3 take bases
        // 0-15, warps 4-7 bases 16-31.  Transposed through shared memory (the flatten tile is dead) for coalesced row stores.
        float* outT = reinterpret_cast<float*>(s.flat);          // [32 bases][64 features]
        {
            const int lane = tid & 31, q = warp & 3, half = warp >> 2;
            uint32_t v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 16))
                : "memory");
            tmem_ld_wait();
            if (lane < 16) {
                const int f = q * 16 + lane;
                const float bias = __ldg(W.dense_b + f);
#pragma unroll
                for (int n = 0; n < 16; ++n) outT[(half * 16 + n) * 64 + f] = __uint_as_float(v[n]) + bias;
            }
        }
        tc_fence_before();
        __syncthreads();
        for (int i = tid; i < CNN_TB * 16; i += CNN_THREADS) {          // 4 features per thread: 16-byte fp32 / 8-byte fp16 stores
            const int b = i >> 4, c4 = (i & 15) * 4;
            const int64_t j = j0 + b;
            if (j >= n_bases) continue;
            const float4 y = *reinterpret_cast<const float4*>(outT + b * 64 + c4);
            if (sig_feat) *reinterpret_cast<float4*>(sig_feat + j * NRV_SIGFEAT + c4) = y;
            if (sf_hi) {   // fp16 (hi, lo) copy: A operand of the tensor-core projection of total_rnn1's per-base part
                uint32_t h01, l01, h23, l23;
                split2(clamp_f16(y.x), clamp_f16(y.y), h01, l01);
                split2(clamp_f16(y.z), clamp_f16(y.w), h23, l23);
                *reinterpret_cast<uint2*>(sf_hi + j * NRV_SIGFEAT + c4) = make_uint2(h01, h23);
                *reinterpret_cast<uint2*>(sf_lo + j * NRV_SIGFEAT + c4) = make_uint2(l01, l23);
            }
        }
    }
    }       // models
    tc_fence_before();
    __syncthreads();
    if (tid < 32) { tc_fence_after(); tmem_dealloc(s.tmem_slot, 32); }
}

int launch_cnn(const ModelDev* m1, const ModelDev* m2, const int16_t* signal, const int64_t* sig_off,
               const int32_t* starts, const int64_t* /*base_off*/, const int32_t* base_read,
               const double* shift, const double* scale, const float* explicit_win, int64_t n_bases,
               float* sig_feat1, float* sig_feat2, __half* const sf_hi[2], __half* const sf_lo[2], cudaStream_t st) {
    if (n_bases <= 0 || !m1) return 0;
    static PerDevice attr_set;
    if (attr_set.first()) cudaFuncSetAttribute(cnn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CnnSmem) + 1024);
    CnnTmaps tm[2];
    for (int k = 0; k < 2; ++k) {
        const CnnDev& c = (k && m2) ? m2->cnn : m1->cnn;
        if (!c.dt_hi || !c.dt_lo) return -1;
        if (!make_tmap_f16_k64(&tm[k].w_hi, c.dt_hi, 64, 400, 64) || !make_tmap_f16_k64(&tm[k].w_lo, c.dt_lo, 64, 400, 64) ||
            !make_tmap_f16_k16_sw32(&tm[k].t_hi, c.dt_hi, 64, 400, 64) || !make_tmap_f16_k16_sw32(&tm[k].t_lo, c.dt_lo, 64, 400, 64))
            return -2;
    }
    const unsigned grid = (unsigned)((n_bases + CNN_TB - 1) / CNN_TB);
    cnn_kernel<<<grid, CNN_THREADS, sizeof(CnnSmem) + 1024, st>>>(m1->cnn, m2 ? m2->cnn : m1->cnn, tm[0], tm[1], m2 ? 2 : 1, signal, sig_off, starts,
                                                                   base_read, shift, scale, explicit_win, n_bases, sig_feat1, sig_feat2,
                                                                   sf_hi ? sf_hi[0] : nullptr, sf_lo ? sf_lo[0] : nullptr,
                                                                   sf_hi ? sf_hi[1] : nullptr, sf_lo ? sf_lo[1] : nullptr);
    return 1;
}

}  // namespace nrv
