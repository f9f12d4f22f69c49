// Probe: tcgen05.mma with the A operand in TMEM (".ts" form) -- layout of a 16-bit A tile written with tcgen05.st.
// D[128][64] = A[128][64] . B[64][64]^T, A from TMEM (lane = row, 32-bit column c = fp16 pair (k = 2c, 2c+1)), B K-major SW128 in smem.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include "../../nanoreviser_b200/csrc/nrv_tc.cuh"
using namespace nrv::tc;

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D, int pair_dummy) {
    __shared__ __align__(1024) uint8_t s_b[64 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 128);
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {          // B rows of 64 halves = 8 x 16 B chunks
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(s_b + sw128_offset(row, c)) = reinterpret_cast<const uint4*>(B + row * 64)[c];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    // A -> TMEM columns [64, 96): this thread's row = threadIdx.x
    const int row = threadIdx.x;
    uint32_t a[32];
    for (int c = 0; c < 32; ++c) {
        const __half lo = A[row * 64 + 2 * c], hi = A[row * 64 + 2 * c + 1];
        a[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 64;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(ta),
        "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]),
        "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(a[16]), "r"(a[17]), "r"(a[18]), "r"(a[19]), "r"(a[20]),
        "r"(a[21]), "r"(a[22]), "r"(a[23]), "r"(a[24]), "r"(a[25]), "r"(a[26]), "r"(a[27]), "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(128, 64);
            const uint64_t bd = umma_desc_k_sw128(smem_u32(s_b));
            for (int k = 0; k < 4; ++k) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb),
                    "r"(tb + 64 + k * 8), "l"(bd + (uint64_t)(k * 2)), "r"(idesc), "r"((uint32_t)(k != 0))
                    : "memory");
            }
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int cb = 0; cb < 2; ++cb) {
        tmem_ld_32x32(tb + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[row * 64 + cb * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 128); }
}


// cta_group::2: M = 256 (128 rows per CTA, A in each CTA's own TMEM), N = 64 (32 B rows in each CTA's smem)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair(const __half* A, const __half* B, float* D) {
    __shared__ __align__(1024) uint8_t s_b[32 * 128];
    __shared__ uint64_t bar, a_ready;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&a_ready, 2); fence_mbar_init(); }
    if (warp == 0) tmem_alloc_pair(&slot, 128);
    for (int i = threadIdx.x; i < 32 * 8; i += 128) {
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(s_b + sw128_offset(row, c)) = reinterpret_cast<const uint4*>(B + (rank * 32 + row) * 64)[c];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tb = slot;
    const int row = rank * 128 + threadIdx.x;
    uint32_t a[32];
    for (int c = 0; c < 32; ++c) {
        const __half lo = A[row * 64 + 2 * c], hi = A[row * 64 + 2 * c + 1];
        a[c] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
    }
    const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 64;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(ta),
        "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]),
        "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]), "r"(a[16]), "r"(a[17]), "r"(a[18]), "r"(a[19]), "r"(a[20]),
        "r"(a[21]), "r"(a[22]), "r"(a[23]), "r"(a[24]), "r"(a[25]), "r"(a[26]), "r"(a[27]), "r"(a[28]), "r"(a[29]), "r"(a[30]), "r"(a[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive_remote(&a_ready, 0);       // both CTAs' A tiles are in TMEM
    if (warp == 0 && rank == 0) {
        mbar_wait(&a_ready, 0);
        tc_fence_after();
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(256, 64);
            const uint64_t bd = umma_desc_k_sw128(smem_u32(s_b));
            for (int k = 0; k < 4; ++k) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb),
                    "r"(tb + 64 + k * 8), "l"(bd + (uint64_t)(k * 2)), "r"(idesc), "r"((uint32_t)(k != 0))
                    : "memory");
            }
            umma_commit_pair(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int cb = 0; cb < 2; ++cb) {
        tmem_ld_32x32(tb + ((uint32_t)(warp * 32) << 16) + cb * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[row * 64 + cb * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tb, 128); }
}

int main() {
    std::vector<__half> A(128 * 64), B(64 * 64);
    std::vector<float> Af(128 * 64), Bf(64 * 64), D(128 * 64), R(128 * 64, 0.f);
    srand(1);
    for (size_t i = 0; i < A.size(); ++i) { Af[i] = (rand() % 2001 - 1000) / 1000.f; A[i] = __float2half(Af[i]); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (rand() % 2001 - 1000) / 1000.f; B[i] = __float2half(Bf[i]); Bf[i] = __half2float(B[i]); }
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)Af[r * 64 + k] * Bf[n * 64 + k]; R[r * 64 + n] = (float)s; }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dD, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0; for (size_t i = 0; i < D.size(); ++i) mx = fmax(mx, fabs(D[i] - R[i]));
    printf("TS-mode cta_group::1 max abs err = %g  (D[0]=%g ref %g, D[5*64+7]=%g ref %g)\n", mx, D[0], R[0], D[5 * 64 + 7], R[5 * 64 + 7]);
    {   // pair: A [256][64]
        std::vector<__half> A2(256 * 64); std::vector<float> A2f(256 * 64), D2(256 * 64), R2(256 * 64);
        for (size_t i = 0; i < A2.size(); ++i) { A2[i] = __float2half((rand() % 2001 - 1000) / 1000.f); A2f[i] = __half2float(A2[i]); }
        for (int r = 0; r < 256; ++r) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)A2f[r * 64 + k] * Bf[n * 64 + k]; R2[r * 64 + n] = (float)s; }
        __half* dA2; float* dD2;
        cudaMalloc(&dA2, A2.size() * 2); cudaMalloc(&dD2, D2.size() * 4);
        cudaMemcpy(dA2, A2.data(), A2.size() * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD2, 0, D2.size() * 4);
        probe_pair<<<2, 128>>>(dA2, dB, dD2);
        e = cudaDeviceSynchronize();
        printf("pair kernel: %s\n", cudaGetErrorString(e));
        cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost);
        double m0 = 0, m1 = 0;
        for (int i = 0; i < 128 * 64; ++i) { m0 = fmax(m0, fabs(D2[i] - R2[i])); m1 = fmax(m1, fabs(D2[128 * 64 + i] - R2[128 * 64 + i])); }
        printf("TS-mode cta_group::2 max abs err: CTA0 rows %g, CTA1 rows %g\n", m0, m1);
    }
    return 0;
}
