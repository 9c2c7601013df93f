// CUDA-core convolution kernels:
//   * conv_q8_direct_kernel     int8 direct conv on the quantized NHWC activations, any groups (depthwise layers,
//                               and the cross-check for the tcgen05 path) — same epilogue as the tensor-core kernel;
//   * weightonly_kernel         the reference op's weight-only semantic in fp32, in the reference's own accumulation
//                               order (quantconv2d_float_input.cu:83-119), so results are bit-identical to it.
#include "common.cuh"
#include "conv_common.cuh"

namespace qb200 {
namespace {

// ------------------------------------------------------------------------------------------------
// int8 direct conv.  One thread per output element, q fastest (coalesced NCHW stores).
// ------------------------------------------------------------------------------------------------
template <bool kSignedW>
__global__ void __launch_bounds__(256)
conv_q8_direct_kernel(const uint8_t* __restrict__ qa, const uint8_t* __restrict__ wq, ConvGeom g, EpilogueParams ep,
                      void* __restrict__ out) {
    const int64_t total = (int64_t)g.N * g.K * g.P * g.Q;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int q = (int)(idx % g.Q);
    int64_t t = idx / g.Q;
    const int p = (int)(t % g.P);
    t /= g.P;
    const int k = (int)(t % g.K);
    const int n = (int)(t / g.K);
    const int Kg = g.K / g.groups;
    const int grp = k / Kg;

    int32_t acc = 0;
    for (int r = 0; r < g.R; ++r) {
        const int ih = p * g.stride - g.pad + r;
        if (ih < 0 || ih >= g.H) continue;
        for (int s = 0; s < g.S; ++s) {
            const int iw = q * g.stride - g.pad + s;
            if (iw < 0 || iw >= g.W) continue;
            const uint8_t* ap = qa + (((int64_t)n * g.H + ih) * g.W + iw) * g.Cp + grp * g.Cg;
            const uint8_t* wp = wq + (((int64_t)k * g.R + r) * g.S + s) * g.Cgp;
            int c = 0;
            if (((reinterpret_cast<uintptr_t>(ap) | reinterpret_cast<uintptr_t>(wp)) & 3) == 0) {
                for (; c + 4 <= g.Cg; c += 4) {
                    const uint32_t av = *reinterpret_cast<const uint32_t*>(ap + c);
                    const uint32_t wv = *reinterpret_cast<const uint32_t*>(wp + c);
                    if (kSignedW) asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(av), "r"(wv));  // u8 x s8
                    else asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(av), "r"(wv));         // u8 x u8
                }
            }
            for (; c < g.Cg; ++c) {
                const int32_t wv = kSignedW ? (int32_t)(int8_t)wp[c] : (int32_t)wp[c];
                acc += (int32_t)ap[c] * wv;
            }
        }
    }
    if (ep.out_kind == QB200_OUT_ACC) {
        static_cast<int32_t*>(out)[idx] = acc;
    } else {
        const EpilogueScalars es = load_epilogue_scalars(ep);
        const PixelWindow pw = pixel_window(g, p, q);
        static_cast<float*>(out)[idx] = epilogue_tail(dequant_one(acc, k, g, ep, es, pw), idx, ep);
    }
}

// ------------------------------------------------------------------------------------------------
// weight-only fp32 conv (reference order).  Block = one (n, k) and 256 consecutive output pixels;
// the k-th filter is dequantized once into shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
weightonly_kernel(const float* __restrict__ x, const uint8_t* __restrict__ w_packed, const float* __restrict__ w_scale,
                  const float* __restrict__ w_zero, int per_tensor, int nb, int sign, const float* __restrict__ bias,
                  float* __restrict__ out, ConvGeom g) {
    extern __shared__ float wf[];  // [C*R*S]
    const int k = blockIdx.y;
    const int n = blockIdx.z;
    const int crs = g.C * g.R * g.S;
    const uint32_t offset = sign ? (1u << (nb - 1)) : 0u;
    const uint32_t mask = (1u << nb) - 1u;
    const float zero = per_tensor ? w_zero[0] : w_zero[k];
    const float scale = per_tensor ? w_scale[0] : w_scale[k];
    for (int i = threadIdx.x; i < crs; i += blockDim.x) {
        const int64_t e = (int64_t)k * crs + i;                         // :94
        const int64_t bit = e * nb;
        const int64_t byte = bit >> 3;                                  // :95
        const int off = (int)(bit & 7);                                 // :96
        uint32_t v = (w_packed[byte] >> off) & mask;                    // :97
        if (off + nb > 8) v |= ((uint32_t)w_packed[byte + 1] << (8 - off)) & mask;  // :98-99
        const uint8_t u = (uint8_t)(v - offset);                        // :102
        const float wv = sign ? (float)(int8_t)u : (float)u;            // :103
        wf[i] = __fmul_rn(__fsub_rn(wv, zero), scale);                  // :104-106
    }
    __syncthreads();
    const int pq = blockIdx.x * blockDim.x + threadIdx.x;
    if (pq >= g.P * g.Q) return;
    const int p = pq / g.Q, q = pq % g.Q;
    float o = bias ? bias[k] : 0.f;                                     // :83
    const float* xn = x + (int64_t)n * g.C * g.H * g.W;
    for (int c = 0; c < g.C; ++c)                                       // :86
        for (int r = 0; r < g.R; ++r) {                                 // :87
            const int ih = p * g.stride - g.pad + r;                    // :89
            if (ih < 0 || ih >= g.H) continue;
            for (int s = 0; s < g.S; ++s) {                             // :88
                const int iw = q * g.stride - g.pad + s;                // :90
                if (iw < 0 || iw >= g.W) continue;                      // :92
                o = __fmaf_rn(__ldg(xn + ((int64_t)c * g.H + ih) * g.W + iw), wf[(c * g.R + r) * g.S + s], o);  // :112
            }
        }
    out[(((int64_t)n * g.K + k) * g.P + p) * g.Q + q] = o;              // :119
}

// ------------------------------------------------------------------------------------------------
// weight-only fp32 linear (reference order: quantlinear_float_input.cu:60-105): acc = 0; acc = fma(x[k], wf[k], acc)
// for k ascending (nvcc contracts the reference's `+= a * b`); out = acc + bias.  Block = 32 output features x 8 rows;
// a 32 x 32 tile of dequantized weights is staged in shared memory per pass (each weight is unpacked once per 8 rows).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linear_weightonly_kernel(const float* __restrict__ x, const uint8_t* __restrict__ w_packed, const float* __restrict__ w_scale,
                         const float* __restrict__ w_zero, int per_tensor, int nb, int sign, const float* __restrict__ bias,
                         float* __restrict__ out, int batch, int in_f, int out_f) {
    __shared__ float wt[32][33];   // [k][col]
    __shared__ float xt[8][32];    // [row][k]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + tx, row = blockIdx.y * 8 + ty;
    const uint32_t offset = sign ? (1u << (nb - 1)) : 0u;
    const uint32_t mask = (1u << nb) - 1u;
    float acc = 0.f;
    for (int k0 = 0; k0 < in_f; k0 += 32) {
        // weights: thread (ty, tx) dequantizes rows ty, ty+8, ty+16, ty+24 of the k-tile for column blockIdx.x*32 + tx
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = ty + 8 * j;
            float wv = 0.f;
            if (k0 + kk < in_f && col < out_f) {
                const int64_t e = (int64_t)col * in_f + k0 + kk;            // :73
                const int64_t bit = e * nb;
                const int64_t byte = bit >> 3;                               // :74
                const int off = (int)(bit & 7);                              // :75
                uint32_t v = (w_packed[byte] >> off) & mask;                 // :76
                if (off + nb > 8) v |= ((uint32_t)w_packed[byte + 1] << (8 - off)) & mask;   // :77-78
                const uint8_t u = (uint8_t)(v - offset);                     // :81
                const float f = sign ? (float)(int8_t)u : (float)u;          // :82
                const float zero = per_tensor ? w_zero[0] : w_zero[col], scale = per_tensor ? w_scale[0] : w_scale[col];
                wv = __fmul_rn(__fsub_rn(f, zero), scale);                   // :83-85
            }
            wt[kk][tx] = wv;
        }
        xt[ty][tx] = (row < batch && k0 + tx < in_f) ? __ldg(x + (int64_t)row * in_f + k0 + tx) : 0.f;   // :66-68
        __syncthreads();
        const int kn = min(32, in_f - k0);
        for (int j = 0; j < kn; ++j) acc = __fmaf_rn(xt[ty][j], wt[j][tx], acc);   // :94-96
        __syncthreads();
    }
    if (row < batch && col < out_f) out[(int64_t)row * out_f + col] = __fadd_rn(acc, bias ? bias[col] : 0.f);   // :103-104
}

}  // namespace

int launch_conv_direct(const ConvGeom& g, const uint8_t* qa, const uint8_t* wq, const EpilogueParams& ep, void* out,
                       cudaStream_t st) {
    const int64_t total = (int64_t)g.N * g.K * g.P * g.Q;
    if (total == 0) return 0;
    const unsigned blocks = (unsigned)ceil_div64(total, 256);
    if (g.w_sign) conv_q8_direct_kernel<true><<<blocks, 256, 0, st>>>(qa, wq, g, ep, out);
    else conv_q8_direct_kernel<false><<<blocks, 256, 0, st>>>(qa, wq, g, ep, out);
    QB_LAUNCH_CHECK();
    return 0;
}

}  // namespace qb200

extern "C" {

int qb200_quantconv2d_weightonly(const qb200_conv_shape* s, const float* x, const uint8_t* w_packed,
                                 const float* w_scale, const float* w_zero, int32_t n_w_scale, const float* bias,
                                 float* out, void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    QB_REQUIRE(s->C == s->Cg, QB200_EUNSUPPORTED,
               "quantconv2d_float_input without activation quantizer supports groups == 1 only (as the reference)");
    QB_REQUIRE(n_w_scale == 1 || n_w_scale == s->K, QB200_EINVAL, "weight_scale must have 1 or K elements");
    if (s->N == 0) return 0;
    QB_REQUIRE(x && w_packed && w_scale && w_zero && out, QB200_EINVAL, "weightonly: null pointer");
    const ConvGeom g = make_geom(*s);
    const size_t smem = (size_t)g.C * g.R * g.S * sizeof(float);
    QB_REQUIRE(smem <= 200 * 1024, QB200_EUNSUPPORTED, "weightonly: filter of %zu bytes exceeds shared memory", smem);
    if (smem > 48 * 1024) QB_CUDA(cudaFuncSetAttribute(weightonly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((g.P * g.Q + 255) / 256, g.K, g.N);
    QB_REQUIRE(g.K <= 65535 && g.N <= 65535, QB200_EUNSUPPORTED, "weightonly: K and N must be <= 65535");
    weightonly_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        x, w_packed, w_scale, w_zero, n_w_scale == 1, s->w_bits, s->w_sign, bias, out, g);
    QB_LAUNCH_CHECK();
    return 0;
}

int qb200_quantlinear_weightonly(const float* x, int64_t batch, int32_t in_features, int32_t out_features,
                                 const uint8_t* w_packed, int32_t w_bits, int32_t w_sign, const float* w_scale,
                                 const float* w_zero, int32_t n_w_scale, const float* bias, float* out, void* stream) {
    using namespace qb200;
    QB_REQUIRE(batch >= 0 && in_features > 0 && out_features > 0, QB200_EINVAL, "quantlinear: bad sizes");
    QB_REQUIRE(w_bits > 0 && w_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    QB_REQUIRE(n_w_scale == 1 || n_w_scale == out_features, QB200_EINVAL, "weight_scale must have 1 or out_features elements");
    if (batch == 0) return 0;
    QB_REQUIRE(x && w_packed && w_scale && w_zero && out, QB200_EINVAL, "quantlinear: null pointer");
    QB_REQUIRE((batch + 7) / 8 <= 65535, QB200_EUNSUPPORTED, "quantlinear: batch too large");
    const dim3 grid((unsigned)((out_features + 31) / 32), (unsigned)((batch + 7) / 8));
    linear_weightonly_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w_packed, w_scale, w_zero, n_w_scale == 1, w_bits, w_sign, bias, out, (int)batch, in_features, out_features);
    QB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
