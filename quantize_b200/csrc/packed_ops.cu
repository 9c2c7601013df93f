// Ops whose ACTIVATIONS arrive as a tpack'ed stream:
//   quantconv2d   reference engine/kernels/functions/quantconv2d.cu:49-264   (packed u8 NCHW activations x packed weights)
//   quantlinear   reference engine/kernels/functions/quantlinear.cu:39-133, :231-297
//
// quantconv2d.  The reference dequantizes both operands per MAC in fp32, (q - zero) * scale (quantconv2d.cu:113-115,
// :127-129).  With a per-tensor input quantizer and symmetric weights (weight_zero == 0) that factors into the integer
// GEMM this engine already runs:
//     sum_{in-bounds taps} (u - off - z_i) * qw  =  sum u*qw  -  (off + z_i) * sum_{in-bounds taps} qw
// where u is the STORED (offset-binary) value of the stream, an unsigned byte — exactly the A operand of the u8 x s8
// tensor-core kernel — so the op is: unpack_act_nhwc (stream -> NHWC(Cp) bytes) + the tcgen05 conv with activation
// "zero point" -(off + z_i) (its epilogue computes s_a * s_w * (acc + z_a * wsum) + bias).  Per-input-channel input
// scales or asymmetric weights do not factor; they take dequant_packed (stream -> fp32 NCHW, the reference's (q - z) * s)
// followed by the fp32 weight-only kernel, which accumulates in the reference's order (bit-identical to its kernel).
//
// quantlinear.  (q + zero) convention, per-row input scale, k ascending:  tmp += (qi + zi[row]) * (qw + zw[col]) * s[row][col]
// (quantlinear.cu:110-120) — restated with the same roundings (a*b rounded, then one FMA with the scale product), the
// dequantized operands staged as fp32 tiles in shared memory so that every element is unpacked once per tile.
#include <algorithm>
#include "common.cuh"
#include "conv_common.cuh"

namespace qb200 {
namespace {

// element i of an n-bit LSB-first stream (tpack.cu:286-312 / quantconv2d.cu:105-110); `bytes` guards the straddle read
__device__ __forceinline__ uint32_t stream_get(const uint8_t* __restrict__ s, int64_t i, int nb, uint32_t mask, int64_t bytes) {
    const int64_t bit = i * nb;
    const int64_t byte = bit >> 3;
    const int off = (int)(bit & 7);
    uint32_t v = (uint32_t)__ldg(s + byte) >> off;
    if (off + nb > 8 && byte + 1 < bytes) v |= (uint32_t)__ldg(s + byte + 1) << (8 - off);
    return v & mask;
}

constexpr int kUpPix = 128;   // pixels per block (thread = pixel: lanes read adjacent bit fields of one channel plane)
constexpr int kUpCw = 128;    // channel bytes per block
constexpr int kUpRow = 144;   // shared row stride (16-byte multiple)

// packed NCHW stream -> NHWC(Cp) bytes holding the stored values; channels C..Cp-1 are 0.
// block 0 / thread 0 also writes the activation "zero point" of the integer form: zadj = -(offset + z_i).
__global__ void __launch_bounds__(kUpPix)
unpack_act_nhwc_kernel(const uint8_t* __restrict__ packed, int nb, int64_t bytes, int64_t total_pix, int C, int Cp, int HW,
                       uint8_t* __restrict__ q, const float* __restrict__ in_zero, float offset, float* __restrict__ zadj) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(16) uint8_t tile[kUpPix * kUpRow];
    if (zadj != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
        *zadj = -__fadd_rn(offset, __ldg(in_zero));
    const uint32_t mask = (1u << nb) - 1u;
    const int t = threadIdx.x;
    const int64_t g0 = (int64_t)blockIdx.x * kUpPix;
    const int64_t g = g0 + t;
    const int c_base = blockIdx.y * kUpCw;
    const int cw = min(kUpCw, Cp - c_base);   // multiple of 32
    if (g < total_pix) {
        const int64_t n = g / HW;
        const int pix = (int)(g - n * HW);
        for (int c = 0; c < cw; c += 4) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int ch = c_base + c + b;
                if (ch < C) word |= stream_get(packed, (n * C + ch) * (int64_t)HW + pix, nb, mask, bytes) << (8 * b);
            }
            *reinterpret_cast<uint32_t*>(tile + t * kUpRow + c) = word;
        }
    }
    __syncthreads();
    const int n_rows = (int)min((int64_t)kUpPix, total_pix - g0);
    const int chunks = cw >> 4;
    for (int i = t; i < n_rows * chunks; i += kUpPix) {
        const int row = i / chunks, col = i - row * chunks;
        *reinterpret_cast<uint4*>(q + (g0 + row) * (int64_t)Cp + c_base + col * 16) =
            *reinterpret_cast<const uint4*>(tile + row * kUpRow + col * 16);
    }
}

// stream -> fp32, the reference's dequantization:  plus_zero ? (q + z) * s : (q - z) * s,  z / s per tensor or indexed by
// channel = (i / inner) % C.   q = (input_t)(u - offset)  (quantconv2d.cu:111-112)
__global__ void __launch_bounds__(256)
dequant_packed_kernel(const uint8_t* __restrict__ packed, int nb, int sign, int64_t bytes, int64_t n, int64_t inner, int C,
                      const float* __restrict__ scale, const float* __restrict__ zero, int per_tensor, int plus_zero,
                      float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t mask = (1u << nb) - 1u, offset = sign ? (1u << (nb - 1)) : 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t u = (uint8_t)(stream_get(packed, i, nb, mask, bytes) - offset);
        const float f = sign ? (float)(int8_t)u : (float)u;
        const int c = per_tensor ? 0 : (int)((i / inner) % C);
        const float z = __ldg(zero + c), s = __ldg(scale + c);
        out[i] = __fmul_rn(plus_zero ? __fadd_rn(f, z) : __fsub_rn(f, z), s);
    }
}

// quantlinear: 32 x 32 output tile per block, 256 threads, thread (ty, tx) owns rows ty, ty+8, ty+16, ty+24 of column tx.
__global__ void __launch_bounds__(256)
quantlinear_kernel(const uint8_t* __restrict__ in_packed, int in_bits, int in_sign, int64_t in_bytes,
                   const float* __restrict__ in_scale, const float* __restrict__ in_zero,
                   const uint8_t* __restrict__ w_packed, int w_bits, int w_sign, int64_t w_bytes,
                   const float* __restrict__ w_scale, const float* __restrict__ w_zero, const float* __restrict__ bias,
                   float* __restrict__ out, int batch, int in_f, int out_f) {
    __shared__ float at[32][33];   // [row][k]   (qi + zi[row])            quantlinear.cu:110-112
    __shared__ float bt[32][33];   // [k][col]   (qw + zw[col])            :115-117
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int row0 = blockIdx.y * 32, col = blockIdx.x * 32 + tx;
    const uint32_t imask = (1u << in_bits) - 1u, ioff = in_sign ? (1u << (in_bits - 1)) : 0u;
    const uint32_t wmask = (1u << w_bits) - 1u, woff = w_sign ? (1u << (w_bits - 1)) : 0u;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, sc[4];
    const float ws = col < out_f ? __ldg(w_scale + col) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = row0 + ty + 8 * j;
        sc[j] = r < batch ? __fmul_rn(__ldg(in_scale + r), ws) : 0.f;       // :96
    }
    for (int k0 = 0; k0 < in_f; k0 += 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rr = ty + 8 * j;
            // activations: row row0+rr, element k0+tx
            float av = 0.f;
            if (row0 + rr < batch && k0 + tx < in_f) {
                const uint8_t u = (uint8_t)(stream_get(in_packed, (int64_t)(row0 + rr) * in_f + k0 + tx, in_bits, imask, in_bytes) - ioff);
                av = __fadd_rn(in_sign ? (float)(int8_t)u : (float)u, __ldg(in_zero + row0 + rr));
            }
            at[rr][tx] = av;
            // weights: output feature blockIdx.x*32 + rr, element k0+tx, stored transposed
            float wv = 0.f;
            const int oc = blockIdx.x * 32 + rr;
            if (oc < out_f && k0 + tx < in_f) {
                const uint8_t u = (uint8_t)(stream_get(w_packed, (int64_t)oc * in_f + k0 + tx, w_bits, wmask, w_bytes) - woff);
                wv = __fadd_rn(w_sign ? (float)(int8_t)u : (float)u, __ldg(w_zero + oc));
            }
            bt[tx][rr] = wv;
        }
        __syncthreads();
        const int kn = min(32, in_f - k0);
        for (int k = 0; k < kn; ++k) {
            const float b = bt[k][tx];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = __fmaf_rn(__fmul_rn(at[ty + 8 * j][k], b), sc[j], acc[j]);   // :120
        }
        __syncthreads();
    }
    if (col < out_f) {
        const float b = bias ? __ldg(bias + col) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + ty + 8 * j;
            if (r < batch) out[(int64_t)r * out_f + col] = __fadd_rn(acc[j], b);   // :127
        }
    }
}

}  // namespace
}  // namespace qb200

extern "C" {

int qb200_unpack_act_nhwc(const uint8_t* packed, int32_t n_bits, int32_t sign, int32_t N, int32_t C, int32_t H, int32_t W,
                          uint8_t* q_nhwc, const float* in_zero, float* zero_adj, void* stream) {
    using namespace qb200;
    QB_REQUIRE(n_bits > 0 && n_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    QB_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, QB200_EINVAL, "unpack_act: bad shape");
    if (N == 0) return 0;
    QB_REQUIRE(packed && q_nhwc, QB200_EINVAL, "unpack_act: null pointer");
    QB_REQUIRE((zero_adj == nullptr) == (in_zero == nullptr), QB200_EINVAL, "unpack_act: in_zero and zero_adj go together");
    QB_REQUIRE(reinterpret_cast<uintptr_t>(q_nhwc) % 16 == 0, QB200_EINVAL, "unpack_act: output must be 16-B aligned");
    const int Cp = qb200_padded_channels(C);
    const int HW = H * W;
    const int64_t total = (int64_t)N * HW;
    const int64_t bytes = qb200_packed_bytes(total * C, n_bits);
    dim3 grid((unsigned)ceil_div64(total, kUpPix), (unsigned)((Cp + kUpCw - 1) / kUpCw));
    const float offset = sign ? (float)(1 << (n_bits - 1)) : 0.f;
    QB_CUDA(launch_pdl(unpack_act_nhwc_kernel, grid, dim3(kUpPix), 0, static_cast<cudaStream_t>(stream), packed, (int)n_bits, bytes, total,
                       (int)C, Cp, HW, q_nhwc, in_zero, offset, zero_adj));
    QB_LAUNCH_CHECK();
    return 0;
}

int qb200_dequant_packed_f32(const uint8_t* packed, int32_t n_bits, int32_t sign, int64_t n_elements, int64_t inner, int32_t C,
                             const float* scale, const float* zero, int32_t n_scale, int32_t plus_zero, float* out,
                             void* stream) {
    using namespace qb200;
    QB_REQUIRE(n_bits > 0 && n_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    QB_REQUIRE(n_elements >= 0 && inner > 0 && C > 0 && (n_scale == 1 || n_scale == C), QB200_EINVAL, "dequant_packed: bad sizes");
    if (n_elements == 0) return 0;
    QB_REQUIRE(packed && scale && zero && out, QB200_EINVAL, "dequant_packed: null pointer");
    const int blocks = (int)std::min<int64_t>(ceil_div64(n_elements, 256), (int64_t)kNumSMs * 16);
    QB_CUDA(launch_pdl(dequant_packed_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), packed, (int)n_bits,
                       (int)(sign != 0), qb200_packed_bytes(n_elements, n_bits), n_elements, inner, (int)C, scale, zero,
                       (int)(n_scale == 1), (int)(plus_zero != 0), out));
    QB_LAUNCH_CHECK();
    return 0;
}

int qb200_quantconv2d_packed(const qb200_conv_shape* s, const uint8_t* in_packed, int32_t in_bits, int32_t in_sign,
                             const float* in_scale, const float* in_zero, const void* prepared, const float* w_scale,
                             int32_t n_w_scale, const float* bias, void* workspace, void* out, int32_t out_kind, void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (s->N == 0) return 0;
    QB_REQUIRE(in_packed && in_scale && in_zero && workspace, QB200_EINVAL, "quantconv2d: null pointer");
    // workspace: NHWC(Cp) bytes, then (256-byte aligned) one float for the integer form's zero point
    const size_t q_bytes = align_up_sz((size_t)s->N * s->H * s->W * qb200_padded_channels(s->C), 256);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* zadj = reinterpret_cast<float*>(ws + q_bytes);
    if (int rc = qb200_unpack_act_nhwc(in_packed, in_bits, in_sign, s->N, s->C, s->H, s->W, ws, in_zero, zadj, stream)) return rc;
    qb200_act_quant aq = {in_scale, zadj, nullptr, nullptr};
    return qb200_conv2d_q8_nhwc(s, ws, prepared, w_scale, n_w_scale, bias, &aq, out, out_kind, stream);
}

size_t qb200_quantconv2d_packed_workspace_bytes(const qb200_conv_shape* s) {
    if (!s) return 0;
    return qb200::align_up_sz((size_t)s->N * s->H * s->W * qb200_padded_channels(s->C), 256) + 256;
}

int qb200_quantlinear_packed(const uint8_t* in_packed, int32_t in_bits, int32_t in_sign, const float* in_scale,
                             const float* in_zero, int64_t batch, int32_t in_features, int32_t out_features,
                             const uint8_t* w_packed, int32_t w_bits, int32_t w_sign, const float* w_scale, const float* w_zero,
                             const float* bias, float* out, void* stream) {
    using namespace qb200;
    QB_REQUIRE(batch >= 0 && in_features > 0 && out_features > 0, QB200_EINVAL, "quantlinear: bad sizes");
    QB_REQUIRE(in_bits > 0 && in_bits <= 8 && w_bits > 0 && w_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    if (batch == 0) return 0;
    QB_REQUIRE(in_packed && in_scale && in_zero && w_packed && w_scale && w_zero && out, QB200_EINVAL, "quantlinear: null pointer");
    QB_REQUIRE((batch + 31) / 32 <= 65535, QB200_EUNSUPPORTED, "quantlinear: batch too large");
    const dim3 grid((unsigned)((out_features + 31) / 32), (unsigned)((batch + 31) / 32));
    quantlinear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in_packed, in_bits, in_sign != 0, qb200_packed_bytes(batch * in_features, in_bits), in_scale, in_zero, w_packed, w_bits,
        w_sign != 0, qb200_packed_bytes((int64_t)out_features * in_features, w_bits), w_scale, w_zero, bias, out, (int)batch,
        in_features, out_features);
    QB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
