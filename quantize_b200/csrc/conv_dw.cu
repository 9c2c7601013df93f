// Depthwise layers (groups == C, one input channel per filter: MobileNetV2's 3x3 convs) as ONE fused kernel:
// fp32 NCHW in -> quantize -> integer 2-D stencil -> dequant (+ bias, optional tail) -> fp32 NCHW out.
//
// A depthwise layer has no reduction over channels, so there is nothing for the tensor pipe to do and the layer is
// purely HBM-bound (4 bytes in + 4 bytes out per element).  One block owns a band of output rows of several channels of
// one image: the input patch it needs is read once with coalesced loads, quantized once into shared memory with
// the engine's exact quantizer (quant_math.cuh), and every output is R*S integer multiply-adds from shared memory.
// Out-of-image taps contribute nothing to sum(qa*qw) and are excluded from the zero-point term, exactly like the
// reference module (zero padding of the DEQUANTIZED activation, quantconv2d.py:207-210); the dequant uses the same two
// fused multiply-adds as every other kernel of the engine, so results are bit-identical to the CUDA-core kernel.
#include <algorithm>
#include "common.cuh"
#include "conv_common.cuh"
#include "quant_math.cuh"

namespace qb200 {
namespace {

constexpr int kDwThreads = 256, kDwWarps = kDwThreads / 32;

// One block = CH consecutive channels x a band of TH output rows (full width) of one image.  Enough bytes per block
// (tens of KB of fp32 loads issued before the first use) to cover HBM latency: the first version, one 32 x 16 tile of
// one channel per block, had 2.4 KB in flight per block and ran at 0.5 TB/s.
// KS: compile-time kernel size (3) or 0 = run-time R x S loops (the generic loops cost ~170 instructions per output)
template <bool kSignedW, int KS>
__global__ void __launch_bounds__(kDwThreads)
conv_dw_fused_kernel(const float* __restrict__ x, const uint8_t* __restrict__ wq, ConvGeom g, EpilogueParams ep,
                     void* __restrict__ out, int CH, int TH, int bands, int in_w_alloc,
                     const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                     const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ uint8_t patch[];                    // [CH][in_h][in_w_alloc] quantized inputs (0 outside the image)
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int n = blockIdx.y;
    const int kg = blockIdx.x / bands, band = blockIdx.x - kg * bands;
    const int k0 = kg * CH, nch = min(CH, g.K - k0);
    const int oh0 = band * TH, n_rows = min(TH, g.P - oh0);
    const int in_h = (TH - 1) * g.stride + g.R, in_w = (g.Q - 1) * g.stride + g.S;
    const int ih0 = oh0 * g.stride - g.pad, iw0 = -g.pad;
    const int taps = g.R * g.S;
    const int mult = g.K / g.groups;                      // output channels per input channel (Cg == 1)
    int* wsm = reinterpret_cast<int*>(patch + (((size_t)CH * in_h * in_w_alloc + 15) & ~(size_t)15));   // [CH][taps]
    for (int i = threadIdx.x; i < nch * taps; i += kDwThreads) {
        const uint8_t b = wq[((int64_t)k0 * taps + i) * g.Cgp];
        wsm[i] = kSignedW ? (int)(int8_t)b : (int)b;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- load + quantize: a warp takes (channel, input row) pairs, lanes run along the row ----
    for (int pr = warp; pr < nch * in_h; pr += kDwWarps) {
        const int ch = pr / in_h, r = pr - ch * in_h;
        const int ih = ih0 + r;
        const int c = (k0 + ch) / mult;
        uint8_t* prow = patch + ((size_t)ch * in_h + r) * in_w_alloc;
        const bool rok = ih >= 0 && ih < g.H;
        const float* xr = x + (((int64_t)n * g.C + c) * g.H + (rok ? ih : 0)) * g.W;
        for (int col = lane; col < in_w; col += 32) {
            const int iw = iw0 + col;
            uint8_t q = 0;
            if (rok && iw >= 0 && iw < g.W) {
                const float v = __ldg(xr + iw);
                if (p.byte_clamp) q = (uint8_t)min(max(quant_int(v, p), p.ilo), p.ihi);
                else q = (uint8_t)(quant_word_exact(v, 0.f, 0.f, 0.f, p.s, p.z, p.lo, p.hi) & 0xFFu);
            }
            prow[col] = q;
        }
    }
    __syncthreads();

    // ---- stencil + dequant: a warp takes (channel, output row) pairs, lanes run along the row ----
    const EpilogueScalars es = load_epilogue_scalars(ep);
    for (int pr = warp; pr < nch * n_rows; pr += kDwWarps) {
        const int ch = pr / n_rows, orow = pr - ch * n_rows;
        const int k = k0 + ch, oh = oh0 + orow;
        const float scale = __fmul_rn(es.s_a, __ldg(ep.w_scale + (ep.per_tensor_w ? 0 : k)));
        const float bias = ep.bias ? __ldg(ep.bias + k) : 0.f;
        const int* wk = wsm + ch * taps;
        const uint8_t* pbase = patch + ((size_t)ch * in_h + orow * g.stride) * in_w_alloc;
        const int h_in = oh * g.stride - g.pad;
        int w9[KS ? KS * KS : 1], wall = 0;
        if (KS) {
#pragma unroll
            for (int i = 0; i < KS * KS; ++i) { w9[i] = wk[i]; wall += w9[i]; }
        }
        const bool rows_in = KS && h_in >= 0 && h_in + KS <= g.H;
        for (int ow = lane; ow < g.Q; ow += 32) {
            const int w_in = ow * g.stride - g.pad;
            const uint8_t* pp = pbase + ow * g.stride;
            int acc = 0, ws = 0;
            if (KS) {
#pragma unroll
                for (int r = 0; r < KS; ++r)
#pragma unroll
                    for (int s2 = 0; s2 < KS; ++s2) acc += (int)pp[r * in_w_alloc + s2] * w9[r * KS + s2];
                if (rows_in && w_in >= 0 && w_in + KS <= g.W) {
                    ws = wall;
                } else {
#pragma unroll
                    for (int r = 0; r < KS; ++r)
#pragma unroll
                        for (int s2 = 0; s2 < KS; ++s2)
                            ws += (h_in + r >= 0 && h_in + r < g.H && w_in + s2 >= 0 && w_in + s2 < g.W) ? w9[r * KS + s2] : 0;
                }
            } else {
                for (int r = 0; r < g.R; ++r) {
                    const bool rok = h_in + r >= 0 && h_in + r < g.H;
                    for (int s2 = 0; s2 < g.S; ++s2) {
                        const int w = wk[r * g.S + s2];
                        acc += (int)pp[r * in_w_alloc + s2] * w;           // out-of-image patch entries are 0
                        ws += (rok && w_in + s2 >= 0 && w_in + s2 < g.W) ? w : 0;
                    }
                }
            }
            const int64_t idx = (((int64_t)n * g.K + k) * g.P + oh) * g.Q + ow;
            if (ep.out_kind == QB200_OUT_ACC) {
                static_cast<int32_t*>(out)[idx] = acc;
            } else {
                float t = (float)acc;
                if (es.z_a != 0.f) t = __fmaf_rn(es.z_a, (float)ws, t);
                static_cast<float*>(out)[idx] = epilogue_tail(__fmaf_rn(scale, t, bias), idx, ep);
            }
        }
    }
}

}  // namespace

bool dw_fused_supported(const ConvGeom& g) { return g.groups == g.C && g.Cg == 1 && g.groups > 1 && g.K <= 65535 && g.N <= 65535; }

int launch_conv_dw_fused(const ConvGeom& g, const float* x, const uint8_t* wq, const EpilogueParams& ep,
                         const qb200_act_quant* aq, void* out, cudaStream_t st) {
    QB_REQUIRE(dw_fused_supported(g), QB200_EUNSUPPORTED, "conv_dw: not a depthwise layer");
    QB_REQUIRE(aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL, "conv_dw: activation quantizer parameters missing");
    QB_REQUIRE(ep.q8_out == nullptr, QB200_EUNSUPPORTED, "conv_dw: no quantized hand-off from depthwise layers");
    // rows per band: up to 16; channels per block: as many as give ~8K outputs (and fit 40 KB of patch)
    const int TH = std::min(16, g.P);
    const int bands = (g.P + TH - 1) / TH;
    const int in_h = (TH - 1) * g.stride + g.R, in_w = (g.Q - 1) * g.stride + g.S;
    const int in_w_alloc = (in_w + 3) & ~3;
    QB_REQUIRE((size_t)in_h * in_w_alloc <= 40 * 1024, QB200_EUNSUPPORTED, "conv_dw: plane too wide");
    int CH = std::max(1, 8192 / (TH * g.Q));
    CH = std::min(CH, std::min(64, g.K));
    while (CH > 1 && (size_t)CH * in_h * in_w_alloc > 40 * 1024) --CH;
    const size_t smem = (((size_t)CH * in_h * in_w_alloc + 15) & ~(size_t)15) + (size_t)CH * g.R * g.S * sizeof(int);
    const int kgroups = (g.K + CH - 1) / CH;
    const dim3 grid((unsigned)(kgroups * bands), (unsigned)g.N);
    const bool k3 = g.R == 3 && g.S == 3;
#define QB_DW_LAUNCH(SIGNED, KSZ)                                                                                              \
    QB_CUDA(launch_pdl(conv_dw_fused_kernel<SIGNED, KSZ>, grid, dim3(kDwThreads), smem, st, x, wq, g, ep, out, CH, TH, bands, \
                       in_w_alloc, aq->scale, aq->zero, aq->qmin, aq->qmax))
    if (g.w_sign && k3) QB_DW_LAUNCH(true, 3);
    else if (g.w_sign) QB_DW_LAUNCH(true, 0);
    else if (k3) QB_DW_LAUNCH(false, 3);
    else QB_DW_LAUNCH(false, 0);
#undef QB_DW_LAUNCH
    QB_LAUNCH_CHECK();
    return 0;
}

}  // namespace qb200
