// Probe: tcgen05.mma kind::f8f6f4 with the A operand in TMEM (".ts" form) -- layout of an 8-bit (e4m3) A tile written with
// tcgen05.st.  D[128][64] = A[128][128] . B[64][128]^T.  Hypothesis under test: lane = row, 32-bit column c = the four
// K-consecutive bytes k = 4c .. 4c + 3 (little endian); one K = 32 instruction = 8 columns.  B is K-major SW128 in shared
// memory (128 e4m3 per 128-byte row).  Prints the max abs error against a host reference (products of e4m3 values are exact in fp32).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp8.h>
#include "../../nanoreviser_b200/csrc/nrv_tc.cuh"
using namespace nrv::tc;

__global__ void __launch_bounds__(128, 1) probe8(const uint8_t* A, const uint8_t* B, float* D) {
    __shared__ __align__(1024) uint8_t s_b[64 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&slot, 128);
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {          // B rows of 128 bytes = 8 x 16 B chunks
        const int row = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(s_b + sw128_offset(row, c)) = reinterpret_cast<const uint4*>(B + row * 128)[c];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    const int row = threadIdx.x;
    const uint32_t ta = tb + ((uint32_t)(warp * 32) << 16) + 64;      // A -> TMEM columns [64, 96)
    for (int c4 = 0; c4 < 8; ++c4) {
        const uint4 v = reinterpret_cast<const uint4*>(A + row * 128)[c4];    // 16 consecutive K bytes -> 4 columns
        tmem_st_32x4(ta + c4 * 4, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(128, 64);    // format codes 0 = E4M3 for kind::f8f6f4
            const uint64_t bd = umma_desc_k_sw128(smem_u32(s_b));
            for (int k = 0; k < 4; ++k) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb),
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

static float e4m3_to_float(uint8_t b) {
    const int s = b >> 7, e = (b >> 3) & 15, m = b & 7;
    float v = e == 0 ? ldexpf((float)m, -9) : ldexpf(1.f + m / 8.f, e - 7);
    return s ? -v : v;
}

int main() {
    std::vector<uint8_t> A(128 * 128), B(64 * 128);
    std::vector<float> D(128 * 64), R(128 * 64);
    srand(7);
    auto rnd8 = []() { uint8_t b; do { b = (uint8_t)(rand() & 0xff); } while ((b & 0x7f) == 0x7f || ((b >> 3) & 15) > 9); return b; };   // no NaN, |v| < 8
    for (auto& a : A) a = rnd8();
    for (auto& b : B) b = rnd8();
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 128; ++k) s += (double)e4m3_to_float(A[r * 128 + k]) * e4m3_to_float(B[n * 128 + k]);
            R[r * 64 + n] = (float)s;
        }
    uint8_t *dA, *dB; float* dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    probe8<<<1, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0, mr = 0;
    for (size_t i = 0; i < D.size(); ++i) { mx = fmax(mx, fabs(D[i] - R[i])); mr = fmax(mr, fabs(R[i])); }
    printf("TS-mode e4m3 cta_group::1 max abs err = %g (max |ref| %g; D[0]=%g ref %g, D[5*64+7]=%g ref %g)\n", mx, mr, D[0], R[0], D[5 * 64 + 7], R[5 * 64 + 7]);
    return mx < 1e-3 * mr ? 0 : 1;
}
