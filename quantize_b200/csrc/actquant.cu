// Activation quantizer: fp32 NCHW -> u8 NHWC(Cp), the first half of the fused hot path.
//
// Restates Quantizer.simulate's quantize step (reference modelzoo/modules/quantizer.py:31 Round.forward,
// :215 `self.round(x, scale, zero).clamp(self.qmin, self.qmax)`) with per-tensor scale/zero:
//     q = clamp(rint(x / scale - zero), qmin, qmax)          all in fp32, IEEE divide, round-half-even
// and changes the layout to channel-last so that the implicit-GEMM conv can fetch K-contiguous im2col
// rows with TMA.  Channels C..Cp-1 are written as 0 (they meet zero weights in the MMA).
//
// Machine mapping: one thread owns one pixel (n,h,w flattened), so every fp32 load is coalesced along
// the pixel dimension (the contiguous one in NCHW); 16 channels are quantized into one 16-byte register
// vector, staged in an XOR-swizzled shared tile of 128 pixels x 128 channel-bytes, and the tile is
// written out as whole contiguous NHWC rows (128-byte lines when Cp >= 128).
// HBM-bound: algorithmic bytes per element = 4 (read) + Cp/C (write).
#include "common.cuh"

namespace qb200 {
namespace {

constexpr int kPix = 128;  // pixels per block == threads per block
constexpr int kCw = 128;   // channel bytes per pass

__device__ __forceinline__ uint32_t quant1(float x, float s, float z, float lo, float hi) {
    float t = __fsub_rn(__fdiv_rn(x, s), z);  // x / scale - zero, no contraction
    t = rintf(t);                             // torch.round: half to even
    t = fminf(fmaxf(t, lo), hi);              // clamp(qmin, qmax); NaN -> lo like torch.clamp? (see note)
    return (uint32_t)(int)t & 0xFFu;
}
// note: torch.clamp propagates NaN; a NaN activation has no uint8 image, the reference would produce a NaN
// output.  The kernel maps NaN to qmin; NaN inputs are outside the contract of a calibrated quantizer.

__global__ void __launch_bounds__(kPix)
act_quantize_nhwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ q, int64_t total_pix, int C,
                         int Cp, int HW, const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                         const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    __shared__ uint4 tile[kPix * (kCw / 16)];
    const int t = threadIdx.x;
    const int64_t g0 = (int64_t)blockIdx.x * kPix;
    const int64_t g = g0 + t;
    const int c_base = blockIdx.y * kCw;
    const int cw = min(kCw, Cp - c_base);  // multiple of 32
    const int chunks = cw >> 4;

    const float s = __ldg(p_scale), z = __ldg(p_zero), lo = __ldg(p_qmin), hi = __ldg(p_qmax);

    if (g < total_pix) {
        const int64_t n = g / HW;
        const int pix = (int)(g - n * HW);
        const float* xp = x + (n * C + c_base) * (int64_t)HW + pix;
        for (int j = 0; j < chunks; ++j) {
            float v[16];
            const int c0 = c_base + j * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (c0 + i < C) ? __ldg(xp + (int64_t)(j * 16 + i) * HW) : 0.f;
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t acc = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = k * 4 + b;
                    const uint32_t qv = (c0 + i < C) ? quant1(v[i], s, z, lo, hi) : 0u;
                    acc |= qv << (8 * b);
                }
                w[k] = acc;
            }
            tile[t * (kCw / 16) + (j ^ (t & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncthreads();
    // copy-out: rows of cw bytes at stride Cp; chunk index i -> (row, col)
    const int n_rows = (int)min((int64_t)kPix, total_pix - g0);
    const int total_chunks = n_rows * chunks;
    for (int i = t; i < total_chunks; i += kPix) {
        const int row = i / chunks, col = i - row * chunks;
        const uint4 v = tile[row * (kCw / 16) + (col ^ (row & 7))];
        *reinterpret_cast<uint4*>(q + (g0 + row) * (int64_t)Cp + c_base + col * 16) = v;
    }
}

}  // namespace
}  // namespace qb200

extern "C" {

int32_t qb200_padded_channels(int32_t C) { return (C + 31) / 32 * 32; }

int qb200_act_quantize_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W,
                            const qb200_act_quant* aq, uint8_t* q_nhwc, void* stream) {
    using namespace qb200;
    QB_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0, QB200_EINVAL, "act_quantize: bad shape");
    QB_REQUIRE(aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL,
               "act_quantize: activation quantizer parameters missing");
    if (N == 0) return 0;
    QB_REQUIRE(x && q_nhwc, QB200_EINVAL, "act_quantize: null pointer");
    QB_REQUIRE(reinterpret_cast<uintptr_t>(q_nhwc) % 16 == 0, QB200_EINVAL, "act_quantize: output must be 16-B aligned");
    const int Cp = qb200_padded_channels(C);
    const int64_t total = (int64_t)N * H * W;
    dim3 grid((unsigned)ceil_div64(total, kPix), (unsigned)((Cp + kCw - 1) / kCw));
    act_quantize_nhwc_kernel<<<grid, kPix, 0, static_cast<cudaStream_t>(stream)>>>(
        x, q_nhwc, total, C, Cp, H * W, aq->scale, aq->zero, aq->qmin, aq->qmax);
    QB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
