// LSTM cell arithmetic shared by the tensor-core recurrence kernels (Keras-2.2.4 LSTM: gates i,f,c,o;
// recurrent_activation = hard_sigmoid; lstmmodel.py:44-51 of the reference).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace nrv {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tanh(x) = 2 / (1 + exp(-2x)) - 1: MUFU.EX2 + MUFU.RCP and three FMA-pipe instructions.  The reciprocal is in (0, 1], so the
// absolute error stays ~3e-7 everywhere (same cancellation near 0 as 1 - 2/(exp(2|x|)+1), one instruction and the copysign
// less); exp -> +inf gives rcp = 0 -> -1, exp -> 0 gives +1: saturates cleanly.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = ex2_approx(x * -2.8853900817779268f);
    return fmaf(2.f, rcp_approx(e + 1.f), -1.f);
}
// Same function with the reciprocal on the FMA pipe (bit-trick seed + 3 Newton steps, ~1e-7 relative) instead of MUFU.RCP: halves
// the XU work of a tanh at the price of 7 more issue slots.  Experiment switch NRV_TANH_NR (1: gate tanh, 2: both tanhs of a unit).
__device__ __forceinline__ float tanh_fast_nr(float x) {
    const float e = fminf(ex2_approx(x * -2.8853900817779268f), 1e30f);
    const float d = e + 1.f;
    float y = __int_as_float(0x7EF311C7 - __float_as_int(d));
    y = y * fmaf(-d, y, 2.f);
    y = y * fmaf(-d, y, 2.f);
    y = y * fmaf(-d, y, 2.f);
    return fmaf(2.f, y, -1.f);
}
__device__ __forceinline__ float hsig(float x) { return __saturatef(fmaf(0.2f, x, 0.5f)); }

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// One 32-column block (8 units x gates i,f,c,o) of the Keras-2.2.4 LSTM cell for this thread's row:
// z = acc (if any) + zin ; c, h update ; h -> fp16 (hi, lo) packed as 2 x 16 bytes.
__device__ __forceinline__ uint32_t half2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ void lstm_cell_block(const uint32_t (&v)[32], bool have_acc, const float4 (&zq)[8], float* c8,
                                                uint4& phi, uint4& plo) {
    float hv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float zi = zq[j].x, zf = zq[j].y, zc = zq[j].z, zo = zq[j].w;
        if (have_acc) {
            zi += __uint_as_float(v[4 * j + 0]); zf += __uint_as_float(v[4 * j + 1]);
            zc += __uint_as_float(v[4 * j + 2]); zo += __uint_as_float(v[4 * j + 3]);
        }
        const float ig = hsig(zi), fg = hsig(zf), gg = tanh_fast(zc), og = hsig(zo);
        const float cn = fmaf(fg, c8[j], ig * gg);
        c8[j] = cn;
        hv[j] = og * tanh_fast(cn);
    }
    // h = hi + lo as fp16 pairs.  Packed conversions (cvt.rn.f16x2.f32 -> F2FP, ALU pipe): the scalar F2F.F16.F32 runs on the
    // quarter-rate XU pipe next to the four MUFU ops of every unit, which is what bounds this epilogue (ncu: XU 43 %).
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const __half2 hi = __floats2half2_rn(hv[2 * p], hv[2 * p + 1]);
        const float2 hf = __half22float2(hi);
        const __half2 lo = __floats2half2_rn(hv[2 * p] - hf.x, hv[2 * p + 1] - hf.y);
        ph[p] = half2_bits(hi); pl[p] = half2_bits(lo);
    }
    phi = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    plo = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

}  // namespace nrv
