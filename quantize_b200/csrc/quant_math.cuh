// Activation-quantizer arithmetic shared by the standalone quantizer kernels (actquant.cu) and the fused
// quantizing producer of the 1x1 tensor-core kernel (conv_umma.cu).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace qb200 {

struct QuantParams {
    float s, z, lo, hi;  // the reference's parameters
    float r;             // RN(1/s)
    float xlo, xhi;      // inputs outside [xlo, xhi] quantize to qmin / qmax; xlo / xhi themselves quantize to exactly qmin / qmax
    int ilo, ihi;        // qmin / qmax as integers (valid when byte_clamp)
    int full_range;      // [qmin, qmax] == [0, 255]: the saturating pack is the whole clamp
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
    // Tight input clamp: xlo = RN((qmin + z) * s) quantizes to exactly qmin (its t = RN(RN(xlo / s) - z) is within an ulp
    // of qmin, far from the .5 boundary) and the computed quantizer is monotone in x (a chain of monotone roundings), so
    // every clamped input lands in [qmin, qmax] and every input below / above the clamp would have been clamped to
    // qmin / qmax by the reference anyway.  No saturation or integer clamp is needed after the rounding.
    p.xlo = __fmul_rn(__fadd_rn(p.lo, p.z), p.s);
    p.xhi = __fmul_rn(__fadd_rn(p.hi, p.z), p.s);
    const bool range_ok = p.lo >= 0.f && p.hi <= 255.f && p.lo <= p.hi && p.lo == rintf(p.lo) && p.hi == rintf(p.hi);
    // the division refinement below needs a normal positive scale with headroom and a bounded zero point
    const bool scale_ok = p.s > 1e-30f && p.s < 1e30f && fabsf(p.z) < 1048576.f;
    p.byte_clamp = (range_ok && scale_ok) ? 1 : 0;
    p.ilo = (int)fminf(fmaxf(p.lo, 0.f), 255.f);
    p.ihi = (int)fminf(fmaxf(p.hi, 0.f), 255.f);
    p.full_range = (p.byte_clamp && p.ilo == 0 && p.ihi == 255) ? 1 : 0;
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

// Two values at once with Blackwell's packed fp32 pipe (FMUL2 / FFMA2 / FADD2: two independent round-to-nearest
// operations per issue slot, each rounded exactly like its scalar form): 3.5 instead of 7 issue slots per value for the
// dependent chain.  -q*s + x is computed as q*(-s) + x and q - z as q + (-z), which are the same real numbers.
__device__ __forceinline__ uint64_t f32x2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void quant_int2(float x0, float x1, const QuantParams& p, int& i0, int& i1) {
    x0 = fminf(fmaxf(x0, p.xlo), p.xhi);
    x1 = fminf(fmaxf(x1, p.xlo), p.xhi);
    const uint64_t x = f32x2_pack(x0, x1), r = f32x2_pack(p.r, p.r), ns = f32x2_pack(-p.s, -p.s),
                   nz = f32x2_pack(-p.z, -p.z), magic = f32x2_pack(12582912.f, 12582912.f);
    uint64_t q, e;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(x), "l"(r));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(e) : "l"(q), "l"(ns), "l"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(e), "l"(r), "l"(q));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(e) : "l"(q), "l"(ns), "l"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(e), "l"(r), "l"(q));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(q), "l"(nz));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(q), "l"(magic));
    uint32_t u0, u1;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(u0), "=r"(u1) : "l"(q));
    i0 = (int)u0;   // 0x4B400000 + q with q in [qmin, qmax] within [0, 255]: the quantized value is the LOW BYTE
    i1 = (int)u1;
}

// four quant_int2 results (value in the low byte of each) -> one word, i0 in the low byte: three byte permutes.  (The
// byte-wise SIMD min / max intrinsics are emulated on sm_100 — 14 instructions per word — and with the tight input clamp
// neither they nor a saturating pack are needed.)
template <bool kFull>
__device__ __forceinline__ uint32_t pack_clamp4(int i0, int i1, int i2, int i3, const QuantParams&) {
    const uint32_t lo = __byte_perm((uint32_t)i0, (uint32_t)i1, 0x0040), hi = __byte_perm((uint32_t)i2, (uint32_t)i3, 0x0040);
    return __byte_perm(lo, hi, 0x5410);
}

// One output word = four channels of one pixel.
// Quantizers whose range is not inside [0, 255] (or with a degenerate scale) take the reference's exact arithmetic;
// the choice is uniform over the kernel.
__device__ __forceinline__ uint32_t quant_word(float x0, float x1, float x2, float x3, const QuantParams& p) {
    if (!p.byte_clamp) return quant_word_exact(x0, x1, x2, x3, p.s, p.z, p.lo, p.hi);
    int q0, q1, q2, q3;
    quant_int2(x0, x1, p, q0, q1);
    quant_int2(x2, x3, p, q2, q3);
    return p.full_range ? pack_clamp4<true>(q0, q1, q2, q3, p) : pack_clamp4<false>(q0, q1, q2, q3, p);
}

// NW words from 4*NW consecutive channels of one pixel, the range / exactness switches hoisted out of the unrolled loop
template <int NW>
__device__ __forceinline__ void quant_row(const float (&r)[4 * NW], uint32_t (&w)[NW], const QuantParams& p) {
    if (!p.byte_clamp) {
#pragma unroll   // (a rolled loop would index r[] dynamically and push the caller's arrays into local memory)
        for (int k = 0; k < NW; ++k) w[k] = quant_word_exact(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3], p.s, p.z, p.lo, p.hi);
        return;
    }
    auto body = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            int q0, q1, q2, q3;
            quant_int2(r[4 * k], r[4 * k + 1], p, q0, q1);
            quant_int2(r[4 * k + 2], r[4 * k + 3], p, q2, q3);
            w[k] = pack_clamp4<kFull>(q0, q1, q2, q3, p);
        }
    };
    if (p.full_range) body(std::true_type{});
    else body(std::false_type{});
}

// 4*NQ channels x four consecutive pixels (v[c] = the 4 pixels of channel c, as loaded) -> w[pixel][word]: NQ words
// of four channels per pixel.  Pixel pairs of one channel sit in adjacent registers, so they feed the packed pipe
// without moves.
template <int NQ>
__device__ __forceinline__ void quant_tile(const float4 (&v)[4 * NQ], uint32_t (&w)[4][NQ], const QuantParams& p) {
    if (!p.byte_clamp) {
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            w[0][k] = quant_word_exact(v[4 * k].x, v[4 * k + 1].x, v[4 * k + 2].x, v[4 * k + 3].x, p.s, p.z, p.lo, p.hi);
            w[1][k] = quant_word_exact(v[4 * k].y, v[4 * k + 1].y, v[4 * k + 2].y, v[4 * k + 3].y, p.s, p.z, p.lo, p.hi);
            w[2][k] = quant_word_exact(v[4 * k].z, v[4 * k + 1].z, v[4 * k + 2].z, v[4 * k + 3].z, p.s, p.z, p.lo, p.hi);
            w[3][k] = quant_word_exact(v[4 * k].w, v[4 * k + 1].w, v[4 * k + 2].w, v[4 * k + 3].w, p.s, p.z, p.lo, p.hi);
        }
        return;
    }
    auto body = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
            int i[4][4];  // [channel][pixel]
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                quant_int2(v[4 * k + c].x, v[4 * k + c].y, p, i[c][0], i[c][1]);
                quant_int2(v[4 * k + c].z, v[4 * k + c].w, p, i[c][2], i[c][3]);
            }
#pragma unroll
            for (int px = 0; px < 4; ++px) w[px][k] = pack_clamp4<kFull>(i[0][px], i[1][px], i[2][px], i[3][px], p);
        }
    };
    if (p.full_range) body(std::true_type{});
    else body(std::false_type{});
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}


}  // namespace qb200
