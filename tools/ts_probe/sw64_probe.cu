// Probe for the ping-pong read_rnn11 kernel: A operand with rows of 64 bytes (K = 32 fp16) and the 64-byte swizzle against a B operand
// with the 128-byte swizzle (K = 64 tile of which only the first 32 columns are used).  M = 128, N = 64, cta_group::1.
// SWIZZLE_64B: rows of 64 B, 16-byte chunk index XOR bits [1,3) of the row, 8-row groups 512 B apart, layout type 4.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include "../../nanoreviser_b200/csrc/nrv_tc.cuh"
using namespace nrv::tc;

__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;      // SBO: 8 rows x 64 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;               // SWIZZLE_64B
    return d;
}
__device__ __forceinline__ uint32_t sw64_offset(int row, int chunk) { return (uint32_t)(row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)); }

__global__ void __launch_bounds__(128, 1) probe(const __half* A, const __half* B, float* D) {
    __shared__ __align__(1024) uint8_t sA[128 * 64];
    __shared__ __align__(1024) uint8_t sB[64 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 64);
    for (int i = threadIdx.x; i < 128 * 4; i += 128) {          // A rows of 32 halves = 4 x 16 B chunks
        const int row = i >> 2, c = i & 3;
        *reinterpret_cast<uint4*>(sA + sw64_offset(row, c)) = reinterpret_cast<const uint4*>(A + row * 32)[c];
    }
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {           // B rows of 64 halves (only K < 32 is multiplied)
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + sw128_offset(row, c)) = reinterpret_cast<const uint4*>(B + row * 64)[c];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(128, 64);
            for (int k = 0; k < 2; ++k)
                umma_f16_ss(tb, desc_sw64(smem_u32(sA) + k * 32), umma_desc_k_sw128(smem_u32(sB) + k * 32), idesc, k != 0);
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
        for (int j = 0; j < 32; ++j) D[threadIdx.x * 64 + cb * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

int main() {
    std::vector<__half> A(128 * 32), B(64 * 64);
    std::vector<float> Af(A.size()), Bf(B.size()), D(128 * 64), R(128 * 64);
    srand(9);
    for (size_t i = 0; i < A.size(); ++i) { A[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Af[i] = __half2float(A[i]); }
    for (size_t i = 0; i < B.size(); ++i) { B[i] = __float2half((rand() % 2001 - 1000) / 1000.f); Bf[i] = __half2float(B[i]); }
    for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) { double s = 0; for (int k = 0; k < 32; ++k) s += (double)Af[r * 32 + k] * Bf[n * 64 + k]; R[r * 64 + n] = (float)s; }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (size_t i = 0; i < D.size(); ++i) mx = fmax(mx, fabs(D[i] - R[i]));
    printf("SW64 A x SW128 B (K = 32): max abs err = %g (D[0] = %g ref %g, D[77*64+5] = %g ref %g)\n", mx, D[0], R[0], D[77 * 64 + 5], R[77 * 64 + 5]);
    return mx < 1e-3 ? 0 : 1;
}
