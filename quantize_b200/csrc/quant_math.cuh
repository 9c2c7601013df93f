// Activation-quantizer arithmetic shared by the standalone quantizer kernels (actquant.cu) and the fused
// quantizing producer of the 1x1 tensor-core kernel (conv_umma.cu).
#pragma once
#include "common.cuh"

namespace qb200 {

struct QuantParams {
    float s, z, lo, hi;  // the reference's parameters
    float r;             // RN(1/s)
    float xlo, xhi;      // inputs outside [xlo, xhi] quantize to qmin / qmax whatever their value (2 steps of margin)
    uint32_t lo4, hi4;   // qmin / qmax replicated into 4 bytes (valid when byte_clamp)
    int byte_clamp;      // 0 <= qmin <= qmax <= 255, both integral, scale normal: the branch-free path applies
};

__device__ __forceinline__ QuantParams load_params(const float* p_scale, const float* p_zero, const float* p_qmin,
                                                   const float* p_qmax) {
    QuantParams p;
    p.s = __ldg(p_scale);
    p.z = __ldg(p_zero);
    p.lo = __ldg(p_qmin);
    p.hi = __ldg(p_qmax);
    p.r = __frcp_rn(p.s);
    p.xlo = __fmul_rn(__fadd_rn(__fadd_rn(p.lo, p.z), -2.f), p.s);
    p.xhi = __fmul_rn(__fadd_rn(__fadd_rn(p.hi, p.z), 2.f), p.s);
    const bool range_ok = p.lo >= 0.f && p.hi <= 255.f && p.lo <= p.hi && p.lo == rintf(p.lo) && p.hi == rintf(p.hi);
    // the division refinement below needs a normal positive scale with headroom and a bounded zero point
    const bool scale_ok = p.s > 1e-30f && p.s < 1e30f && fabsf(p.z) < 1048576.f;
    p.byte_clamp = (range_ok && scale_ok) ? 1 : 0;
    const uint32_t l = (uint32_t)(int)fminf(fmaxf(p.lo, 0.f), 255.f), h = (uint32_t)(int)fminf(fmaxf(p.hi, 0.f), 255.f);
    p.lo4 = l * 0x01010101u;
    p.hi4 = h * 0x01010101u;
    return p;
}

// the reference's arithmetic, verbatim, for the four channels of one output word (results clamped)
static __device__ __noinline__ uint32_t quant_word_exact(float x0, float x1, float x2, float x3, float s, float z, float lo, float hi) {
    const float xs[4] = {x0, x1, x2, x3};
    uint32_t w = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t = __fsub_rn(__fdiv_rn(xs[i], s), z);  // x / scale - zero, no contraction
        t = rintf(t);                                 // torch.round: half to even
        t = fminf(fmaxf(t, lo), hi);                  // clamp(qmin, qmax); a NaN activation maps to qmin
        w |= ((uint32_t)(int)t & 0xFFu) << (8 * i);
    }
    return w;
}

// rint(x / s - z), branch-free and bit-identical to the reference's IEEE arithmetic (before the final clamp):
//   * x is first clamped to [xlo, xhi] (values beyond quantize to qmin / qmax anyway; a NaN becomes xlo -> qmin), so
//     every intermediate below is finite and |t| < 2^21;
//   * q = RN(x / s) without a divide: q0 = RN(x * r) with r = RN(1/s) is within 1.5 ulp; one residual correction
//     q1 = RN(q0 + RN(x - q0*s) * r) (the residual is exact in an FMA) is faithful, and by Markstein's theorem a
//     second one is the correctly rounded quotient;
//   * t = RN(q - z) as the reference; adding 1.5 * 2^23 rounds t to the nearest-even integer in the low mantissa bits.
__device__ __forceinline__ int quant_int(float x, const QuantParams& p) {
    x = fminf(fmaxf(x, p.xlo), p.xhi);
    float q = __fmul_rn(x, p.r);
    float e = __fmaf_rn(-q, p.s, x);
    q = __fmaf_rn(e, p.r, q);
    e = __fmaf_rn(-q, p.s, x);
    q = __fmaf_rn(e, p.r, q);
    const float t = __fsub_rn(q, p.z);
    const float u = __fadd_rn(t, 12582912.f);
    return __float_as_int(u) - 0x4B400000;
}

// One output word = four channels of one pixel: saturating pack to u8, then byte-wise clamp to [qmin, qmax].
// Quantizers whose range is not inside [0, 255] (or with a degenerate scale) take the reference's exact arithmetic;
// the choice is uniform over the kernel.
__device__ __forceinline__ uint32_t quant_word(float x0, float x1, float x2, float x3, const QuantParams& p) {
    if (!p.byte_clamp) return quant_word_exact(x0, x1, x2, x3, p.s, p.z, p.lo, p.hi);
    const int q0 = quant_int(x0, p), q1 = quant_int(x1, p), q2 = quant_int(x2, p), q3 = quant_int(x3, p);
    uint32_t hi16, w;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi16) : "r"(q3), "r"(q2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(q1), "r"(q0), "r"(hi16));
    return __vminu4(__vmaxu4(w, p.lo4), p.hi4);
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}


}  // namespace qb200
