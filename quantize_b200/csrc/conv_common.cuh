// Geometry + dequant epilogue shared by the CUDA-core and the tcgen05 convolution kernels.
#pragma once
#include "common.cuh"

namespace qb200 {

struct ConvGeom {
    int N, C, H, W, K, Cg, R, S, stride, pad, P, Q;
    int groups, Cp, Cgp, w_sign;
};

struct PreparedLayout {
    int Cgp;
    size_t wq_bytes;
    size_t wpre_off;
    int Kcol;          // bytes per row of the materialised-im2col weight matrix (0: layer does not use it)
    size_t wcol_off;
    // tap-major copy for the halo variant: [cblocks][R*S][K][KC] (KC = the k-block width implied by Cgp), so that the
    // weight tiles of several taps are ONE 3-D TMA box; present for spatial kernels with groups == 1 and C > 4
    int tapKC;         // 0: absent
    size_t wtap_off;
    size_t total;
};
static inline int kblock_bytes(int Cp) { return (Cp % 128 == 0) ? 128 : (Cp % 64 == 0) ? 64 : 32; }
// Layers with very few input channels (the RGB stem) waste the tensor pipe and the TMA unit when channels are padded
// to 32 per tap; for them the activation pass writes the im2col matrix itself (one 4-byte word per tap = the pixel's
// <=4 channels) and the conv runs as a plain GEMM over Kcol "channels".
static inline bool uses_im2col_rows(const qb200_conv_shape& s) { return s.C == s.Cg && s.C <= 4 && s.R * s.S > 1; }
static inline int im2col_row_bytes(int R, int S) {
    const int kq = R * S * 4;
    return kq > 64 ? (kq + 127) / 128 * 128 : (kq > 32 ? 64 : 32);
}
// Grouped rows for wide filters (the 7x7 stem): 8 bytes per (channel, filter row) = that row's S <= 8 taps of ONE channel
// (bytes S..7 meet zero weights), i.e. a window of 8 consecutive quantized pixels: 3*7*8 = 168 -> 192 bytes per output
// pixel instead of 49 words = 196 -> 256.  Used when it is the smaller of the two.
static inline int im2col8_row_bytes(int C, int R) {
    const int kq = C * R * 8;
    return kq > 64 ? (kq + 63) / 64 * 64 : (kq > 32 ? 64 : 32);
}
static inline bool im2col_grouped(int C, int R, int S) { return S <= 8 && im2col8_row_bytes(C, R) < im2col_row_bytes(R, S); }
static inline int im2col_kcol(int C, int R, int S) { return im2col_grouped(C, R, S) ? im2col8_row_bytes(C, R) : im2col_row_bytes(R, S); }
PreparedLayout prepared_layout(const qb200_conv_shape& s);
int validate_shape(const qb200_conv_shape* s);

static inline ConvGeom make_geom(const qb200_conv_shape& s) {
    ConvGeom g;
    g.N = s.N; g.C = s.C; g.H = s.H; g.W = s.W; g.K = s.K; g.Cg = s.Cg; g.R = s.R; g.S = s.S;
    g.stride = s.stride; g.pad = s.pad;
    g.P = (s.H + 2 * s.pad - s.R) / s.stride + 1;  // quantconv2d_float_input.cu:178
    g.Q = (s.W + 2 * s.pad - s.S) / s.stride + 1;  // :179
    g.groups = s.C / s.Cg;
    g.Cp = qb200_padded_channels(s.C);
    g.Cgp = prepared_layout(s).Cgp;
    g.w_sign = s.w_sign ? 1 : 0;
    return g;
}

// out = s_a * s_w[k] * (acc + z_a * wsum_valid(k, p, q)) + bias[k]
struct EpilogueParams {
    const float* a_scale;   // device, 1 element
    const float* a_zero;    // device, 1 element
    const float* w_scale;   // device, 1 or K elements
    const float* bias;      // device, K elements or null
    const int32_t* wpre;    // device, [K][R+1][S+1]
    int per_tensor_w;
    int out_kind;
    // optional fused tail (extension beyond the reference op): out = relu(out + residual)
    const float* residual;  // device, same shape as out, or null
    int relu;
    // optional quantized hand-off: the result, quantized with the consumer's activation quantizer, written into the
    // consumer's workspace (NHWC or zero-padded NHWC, q8_cp bytes per pixel): byte address of pixel (n, p, q) =
    // q8_out + ((n * q8_img_pixels + p * q8_row_pixels + q + q8_pixel_off) * q8_cp
    uint8_t* q8_out;        // null: none
    const float *q8_scale, *q8_zero, *q8_qmin, *q8_qmax;
    int q8_cp, q8_img_pixels, q8_row_pixels, q8_pixel_off;
    int store_f32;          // 0: skip the fp32 / int32 store (hand-off only)
};

struct EpilogueScalars {
    float s_a, z_a;
};

// in-bounds tap window of an output pixel: taps r in [r0, r1), s in [s0, s1)
struct PixelWindow {
    int r0, r1, s0, s1;
};

__device__ __forceinline__ EpilogueScalars load_epilogue_scalars(const EpilogueParams& ep) {
    EpilogueScalars es;
    es.s_a = __ldg(ep.a_scale);
    es.z_a = __ldg(ep.a_zero);
    return es;
}

__device__ __forceinline__ PixelWindow pixel_window(const ConvGeom& g, int p, int q) {
    PixelWindow w;
    const int h0 = p * g.stride - g.pad, w0 = q * g.stride - g.pad;  // quantconv2d_float_input.cu:89-90
    w.r0 = max(0, -h0);
    w.r1 = min(g.R, g.H - h0);
    w.s0 = max(0, -w0);
    w.s1 = min(g.S, g.W - w0);
    return w;
}

__device__ __forceinline__ int32_t window_wsum(const int32_t* __restrict__ wpre, int k, int R, int S,
                                               const PixelWindow& w) {
    const int32_t* t = wpre + (int64_t)k * (R + 1) * (S + 1);
    const int s1 = S + 1;
    return __ldg(t + w.r1 * s1 + w.s1) - __ldg(t + w.r0 * s1 + w.s1) - __ldg(t + w.r1 * s1 + w.s0) +
           __ldg(t + w.r0 * s1 + w.s0);
}

__device__ __forceinline__ float dequant_one(int32_t acc, int k, const ConvGeom& g, const EpilogueParams& ep,
                                             const EpilogueScalars& es, const PixelWindow& pw) {
    float t = (float)acc;
    if (es.z_a != 0.f) t = __fmaf_rn(es.z_a, (float)window_wsum(ep.wpre, k, g.R, g.S, pw), t);
    const float sw = __ldg(ep.w_scale + (ep.per_tensor_w ? 0 : k));
    const float b = ep.bias ? __ldg(ep.bias + k) : 0.f;
    float r = __fmaf_rn(__fmul_rn(es.s_a, sw), t, b);
    return r;
}

__device__ __forceinline__ float epilogue_tail(float r, int64_t idx, const EpilogueParams& ep) {
    if (ep.residual) r = __fadd_rn(r, __ldg(ep.residual + idx));
    if (ep.relu) r = fmaxf(r, 0.f);
    return r;
}

int launch_conv_direct(const ConvGeom& g, const uint8_t* qa, const uint8_t* wq, const EpilogueParams& ep, void* out,
                       cudaStream_t st);
// gemm_rows > 0: qa is a materialised im2col matrix [N*P*Q][gemm_rows bytes] and wq is [K][gemm_rows]; the main loop
// then runs as a 1x1 convolution over it while the epilogue keeps the real geometry g.
// x_fused != nullptr: 1x1 / stride-1 layer whose A operand is quantized inside the kernel from the fp32 NCHW input
// (see umma_fused_quant_supported); qa is then ignored.
// halo: qa is the zero-padded NHWC buffer written by launch_act_quantize_padded (see umma_halo_supported).
// pair_mode: CTA-pair variant (tcgen05.mma.cta_group::2) — 0 never, 1 where it was measured to win (deep spatial kernels
// with 256-wide channel tiles), 2 wherever it is supported (tests).
// fq_force: the fused-quantize modes that are gated on the problem size (A-stationary channel tiles, flat pixel tiles,
// cp.async input: only when there is at least one pixel tile per SM) wherever they are supported (tests: small shapes).
int launch_conv_umma(const ConvGeom& g, const uint8_t* qa, const uint8_t* wq, const EpilogueParams& ep, void* out,
                     cudaStream_t st, int gemm_rows = 0, const float* x_fused = nullptr,
                     const qb200_act_quant* aq_fused = nullptr, bool halo = false, int pair_mode = 1, bool fq_force = false);
bool umma_halo_supported(const ConvGeom& g);
bool umma_halo_profitable(const ConvGeom& g);
int launch_act_quantize_padded(const float* x, const ConvGeom& g, const qb200_act_quant* aq, uint8_t* q, cudaStream_t st);
// zero the pad pixels of a [N][H + 2*pad][W + 2*pad][Cp] buffer whose interior another kernel fills (quantized hand-off)
int launch_zero_pad_borders(uint8_t* q, int N, int H, int W, int pad, int Cp, cudaStream_t st);
bool umma_fused_quant_supported(const ConvGeom& g, const float* x);
bool umma_fused_quant_profitable(const ConvGeom& g);
// few-channel layer with grouped im2col rows (the 7x7 RGB stem): rows built in shared memory from the fp32 input
// (launch_conv_umma with gemm_rows > 0 AND x_fused)
bool umma_stem_supported(const ConvGeom& g, const float* x);
// 1x1 / stride > 1 / pad 0: quantize only the pixels the conv reads into a compact [N, P, Q, Cp] buffer
int launch_act_quantize_subsampled(const float* x, const ConvGeom& g, const qb200_act_quant* aq, uint8_t* q, cudaStream_t st);
static inline bool uses_subsampled_input(const ConvGeom& g) { return g.R == 1 && g.S == 1 && g.pad == 0 && g.stride > 1; }
// the geometry the conv kernels see for such a layer: a stride-1 1x1 conv over the compact buffer
static inline ConvGeom subsampled_geom(ConvGeom g) {
    g.H = g.P;
    g.W = g.Q;
    g.stride = 1;
    return g;
}
int launch_act_quantize_im2col(const float* x, const ConvGeom& g, int Kcol, const qb200_act_quant* aq, uint8_t* a_col,
                               cudaStream_t st);
bool umma_supported(const ConvGeom& g);
// depthwise layers (groups == C, Cg == 1): quantizer + stencil + dequant in one CUDA-core kernel (conv_dw.cu)
bool dw_fused_supported(const ConvGeom& g);
int launch_conv_dw_fused(const ConvGeom& g, const float* x, const uint8_t* wq, const EpilogueParams& ep,
                         const qb200_act_quant* aq, void* out, cudaStream_t st);
int watchdog_code();

}  // namespace qb200
