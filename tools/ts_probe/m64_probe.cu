// Probe for the tcgen05 dense of the CNN kernel: cta_group::1, M = 64 (A rows), N = 32, both operands K-major in shared memory,
// K = 64 (one SWIZZLE_128B chunk) + 16 (one SWIZZLE_32B chunk).  Pins (a) which TMEM lanes hold the 64 accumulator rows
// (hypothesis: row r -> lane (r / 16) * 32 + r % 16, "16 data-path lanes per warp") and (b) the SWIZZLE_32B operand layout /
// descriptor (rows of 32 B, 16-byte chunk index XOR bit 2 of the row, SBO = 256 B, layout type 6).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../../nanoreviser_b200/csrc/nrv_tc.cuh"
using namespace nrv::tc;

__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;      // SBO: 8 rows x 32 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;               // SWIZZLE_32B
    return d;
}
__device__ __forceinline__ uint32_t sw32_offset(int row, int chunk) { return (uint32_t)(row * 32 + ((chunk ^ ((row >> 2) & 1)) << 4)); }

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D) {
    __shared__ __align__(1024) uint8_t sA0[64 * 128];
    __shared__ __align__(1024) uint8_t sB0[32 * 128];
    __shared__ __align__(1024) uint8_t sA1[64 * 32];
    __shared__ __align__(1024) uint8_t sB1[32 * 32];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 32);
    for (int i = threadIdx.x; i < 64 * 10; i += 128) {         // A rows of 80 halves = 10 x 16 B chunks
        const int row = i / 10, c = i % 10;
        const uint4 v = reinterpret_cast<const uint4*>(A + row * 80)[c];
        if (c < 8) *reinterpret_cast<uint4*>(sA0 + sw128_offset(row, c)) = v;
        else *reinterpret_cast<uint4*>(sA1 + sw32_offset(row, c - 8)) = v;
    }
    for (int i = threadIdx.x; i < 32 * 10; i += 128) {
        const int row = i / 10, c = i % 10;
        const uint4 v = reinterpret_cast<const uint4*>(B + row * 80)[c];
        if (c < 8) *reinterpret_cast<uint4*>(sB0 + sw128_offset(row, c)) = v;
        else *reinterpret_cast<uint4*>(sB1 + sw32_offset(row, c - 8)) = v;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    {   // poison the accumulator columns so that untouched lanes are recognisable
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(-12345.f);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(tb + ((uint32_t)(warp * 32) << 16)),
            "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
            "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
            "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
            : "memory");
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(64, 32);
            for (int k = 0; k < 4; ++k)
                umma_f16_ss(tb, umma_desc_k_sw128(smem_u32(sA0) + k * 32), umma_desc_k_sw128(smem_u32(sB0) + k * 32), idesc, k != 0);
            umma_f16_ss(tb, desc_sw32(smem_u32(sA1)), desc_sw32(smem_u32(sB1)), idesc, 1);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld_32x32(tb + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[threadIdx.x * 32 + j] = __uint_as_float(v[j]);       // D[lane][column]
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 32); }
}

int main() {
    std::vector<__half> A(64 * 80), B(32 * 80);
    std::vector<float> Af(A.size()), Bf(B.size()), D(128 * 32), R(64 * 32);
    srand(5);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Bf[i] = __half2float(B[i]); }
    for (int r = 0; r < 64; ++r) for (int n = 0; n < 32; ++n) { double s = 0; for (int k = 0; k < 80; ++k) s += (double)Af[r * 80 + k] * Bf[n * 80 + k]; R[r * 32 + n] = (float)s; }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    // which lane holds row r?  match every lane against every reference row
    int n_match = 0;
    double worst = 0;
    for (int lane = 0; lane < 128; ++lane) {
        int best = -1; double be = 1e30;
        for (int r = 0; r < 64; ++r) {
            double err = 0;
            for (int n = 0; n < 32; ++n) err = fmax(err, fabs(D[lane * 32 + n] - R[r * 32 + n]));
            if (err < be) { be = err; best = r; }
        }
        const bool poison = D[lane * 32] == -12345.f;
        if (!poison && be < 1e-2) { printf("lane %3d = row %2d (err %.2e)%s\n", lane, best, be, best == (lane / 32) * 16 + lane % 32 && lane % 32 < 16 ? "" : "   <-- not the hypothesis"); ++n_match; worst = fmax(worst, be); }
        else if (!poison) printf("lane %3d: written, matches no row (closest %d, err %.3g)\n", lane, best, be);
    }
    printf("%d lanes hold a row, worst err %.3g\n", n_match, worst);
    return n_match == 64 ? 0 : 1;
}
