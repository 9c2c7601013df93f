// C-ABI entry points of the fused quantized convolution (see include/qb200.h).
#include <stdlib.h>
#include "common.cuh"
#include "conv_common.cuh"

namespace qb200 {
namespace {

int g_conv_algo = QB200_ALGO_AUTO;

// Few-channel layers (the RGB stem) go through materialised im2col rows: 192 B per output pixel, 617 MB for ResNet-50's
// stem at batch 256 — written by the quantizer and read back by the conv, 1.2 GB of DRAM traffic the op contract does not
// contain.  Processing the batch in chunks of a few images through the SAME workspace region keeps the rows in the
// 126 MB L2: chunk c+1's quantizer overwrites what chunk c's conv has just read.  0 = whole batch at once (the default:
// measured on B200 the chunked stem is SLOWER — 16 images: 1070 us, 32: 690 us, whole batch 433 us — because the row
// writer has one block per 8 output rows and a 16-image chunk is 224 blocks on 148 SMs; kept as a switch, off).
int g_rows_chunk = [] {
    const char* e = getenv("QB200_ROWS_CHUNK");
    return e ? atoi(e) : 0;
}();

int resolve_algo(const ConvGeom& g) {
    int algo = g_conv_algo;
    if (algo == QB200_ALGO_AUTO) algo = umma_supported(g) ? QB200_ALGO_UMMA : QB200_ALGO_DIRECT;
    if (algo == QB200_ALGO_UMMA_TWO_KERNELS || algo == QB200_ALGO_UMMA_FUSED_QUANT || algo == QB200_ALGO_UMMA_PAIR) algo = QB200_ALGO_UMMA;
    return algo;
}

// 1x1 / stride-1 layers: the quantizer runs inside the tensor-core kernel's producer warps (one kernel, no workspace)
// depthwise layers: one fused CUDA-core kernel (product default; the explicit algo settings keep the two-kernel paths)
bool dw_single_kernel(const ConvGeom& g) { return g_conv_algo == QB200_ALGO_AUTO && dw_fused_supported(g); }

// few-channel 7x7-type stems: the grouped im2col rows are built in shared memory inside the tensor-core kernel (no
// 617 MB round trip of the rows through HBM).  QB200_STEM_FUSED=0 keeps the two-kernel path (A/B measurements).
bool stem_single_kernel(const ConvGeom& g, const float* x) {
    static const bool on = [] {
        const char* e = getenv("QB200_STEM_FUSED");
        return !(e && e[0] == '0');
    }();
    if (!on || g_conv_algo == QB200_ALGO_UMMA_TWO_KERNELS || g_conv_algo == QB200_ALGO_UMMA_PAIR || resolve_algo(g) != QB200_ALGO_UMMA)
        return false;
    return umma_stem_supported(g, x);
}

bool single_kernel(const ConvGeom& g, const float* x) {
    if (dw_single_kernel(g)) return true;
    if (stem_single_kernel(g, x)) return true;
    if (g_conv_algo == QB200_ALGO_UMMA_TWO_KERNELS || g_conv_algo == QB200_ALGO_UMMA_PAIR || resolve_algo(g) != QB200_ALGO_UMMA) return false;
    if (!umma_fused_quant_supported(g, x)) return false;
    return g_conv_algo == QB200_ALGO_UMMA_FUSED_QUANT || umma_fused_quant_profitable(g);
}

// stride-1 spatial kernels: zero-padded workspace + one halo load per tile instead of one im2col load per tap
bool workspace_is_padded(const ConvGeom& g) {
    if (g_conv_algo == QB200_ALGO_UMMA_TWO_KERNELS || g_conv_algo == QB200_ALGO_UMMA_PAIR || resolve_algo(g) != QB200_ALGO_UMMA ||
        !umma_halo_supported(g))
        return false;
    return g_conv_algo == QB200_ALGO_UMMA_FUSED_QUANT || umma_halo_profitable(g);   // algo 4 forces every variant (tests)
}  // (umma_halo_supported implies the tap-major weight copy exists: spatial kernel, groups == 1, C > 4)

// does the conv kernel chosen for this shape read materialised im2col rows (few-channel layers) or NHWC(Cp) bytes?
bool workspace_is_im2col(const ConvGeom& g, const PreparedLayout& L) { return resolve_algo(g) == QB200_ALGO_UMMA && L.Kcol > 0; }

// Quantized hand-off: can the tensor-core producer `g` write the workspace of consumer `n` directly?
// (consumer layouts handled: NHWC(Cp) and zero-padded NHWC; not im2col rows or the sub-sampled compact buffer)
bool handoff_ok(const qb200_conv_shape& s, const qb200_conv_shape& n) {
    const ConvGeom g = make_geom(s), gn = make_geom(n);
    if (resolve_algo(g) != QB200_ALGO_UMMA) return false;
    if (n.N != s.N || n.C != s.K || n.H != g.P || n.W != g.Q) return false;
    if (workspace_is_im2col(gn, prepared_layout(n)) || uses_subsampled_input(gn)) return false;
    return true;
}

// images per chunk for a layer whose workspace holds im2col rows (0: run the whole batch at once)
int rows_chunk_images(const qb200_conv_shape& s, const qb200_conv_tail* tail) {
    if (g_rows_chunk <= 0 || s.N <= g_rows_chunk || (tail && tail->next_shape)) return 0;
    const ConvGeom g = make_geom(s);
    if (!workspace_is_im2col(g, prepared_layout(s))) return 0;
    // only worth it when the rows of the whole batch do not fit L2 anyway
    const PreparedLayout L = prepared_layout(s);
    if ((int64_t)s.N * g.P * g.Q * L.Kcol < (64ll << 20)) return 0;
    return g_rows_chunk;
}

int quantize_input(const qb200_conv_shape* s, const float* x, const qb200_act_quant* aq, uint8_t* ws, cudaStream_t st) {
    QB_REQUIRE(x && ws, QB200_EINVAL, "conv: null pointer");
    const ConvGeom g = make_geom(*s);
    const PreparedLayout L = prepared_layout(*s);
    if (workspace_is_im2col(g, L)) return launch_act_quantize_im2col(x, g, L.Kcol, aq, ws, st);
    if (uses_subsampled_input(g)) return launch_act_quantize_subsampled(x, g, aq, ws, st);
    if (workspace_is_padded(g)) return launch_act_quantize_padded(x, g, aq, ws, st);
    return qb200_act_quantize_nhwc(x, s->N, s->C, s->H, s->W, aq, ws, st);
}

// from_ws: `q` is the workspace written by quantize_input (layout per workspace_is_im2col); else NHWC(Cp) bytes
int run_conv(const qb200_conv_shape* s, const uint8_t* q, bool from_ws, const void* prepared, const float* w_scale,
             int32_t n_w_scale, const float* bias, const qb200_act_quant* aq, void* out, int32_t out_kind, cudaStream_t st,
             const float* x_fused = nullptr, const qb200_conv_tail* tail = nullptr) {
    QB_REQUIRE(n_w_scale == 1 || n_w_scale == s->K, QB200_EINVAL, "weight_scale must have 1 or K elements");
    QB_REQUIRE(out_kind == QB200_OUT_F32 || out_kind == QB200_OUT_ACC, QB200_EINVAL, "conv: bad out_kind");
    QB_REQUIRE(aq && aq->scale && aq->zero, QB200_EINVAL, "conv: activation quantizer parameters missing");
    QB_REQUIRE((q || x_fused) && prepared && w_scale, QB200_EINVAL, "conv: null pointer");
    const ConvGeom g = make_geom(*s);
    const PreparedLayout L = prepared_layout(*s);
    const uint8_t* wq = static_cast<const uint8_t*>(prepared);
    EpilogueParams ep;
    ep.a_scale = aq->scale;
    ep.a_zero = aq->zero;
    ep.w_scale = w_scale;
    ep.bias = bias;
    ep.wpre = reinterpret_cast<const int32_t*>(wq + L.wpre_off);
    ep.per_tensor_w = n_w_scale == 1;
    ep.out_kind = out_kind;
    ep.residual = tail ? tail->residual : nullptr;
    ep.relu = tail ? (tail->relu != 0) : 0;
    ep.q8_out = nullptr;
    ep.q8_scale = ep.q8_zero = ep.q8_qmin = ep.q8_qmax = nullptr;
    ep.q8_cp = ep.q8_img_pixels = ep.q8_row_pixels = ep.q8_pixel_off = 0;
    ep.store_f32 = 1;
    const bool handoff = tail && tail->next_shape;
    QB_REQUIRE(out_kind == QB200_OUT_F32 || (!ep.residual && !ep.relu && !handoff), QB200_EINVAL,
               "conv: the fused tail applies to the fp32 output only");
    QB_REQUIRE(out || handoff, QB200_EINVAL, "conv: null output");
    if (handoff) {
        const qb200_conv_shape& n = *tail->next_shape;
        if (int rc = validate_shape(&n)) return rc;
        const qb200_act_quant* nq = tail->next_quant;
        QB_REQUIRE(nq && nq->scale && nq->zero && nq->qmin && nq->qmax && tail->next_workspace, QB200_EINVAL,
                   "conv: hand-off needs the consumer's quantizer and workspace");
        QB_REQUIRE(handoff_ok(*s, n) && !x_fused, QB200_EUNSUPPORTED, "conv: this layer pair cannot be chained (qb200_conv_handoff_supported)");
        QB_REQUIRE(reinterpret_cast<uintptr_t>(tail->next_workspace) % 16 == 0, QB200_EINVAL, "conv: hand-off workspace must be 16-byte aligned");
        const ConvGeom gn = make_geom(n);
        ep.q8_out = static_cast<uint8_t*>(tail->next_workspace);
        ep.q8_scale = nq->scale; ep.q8_zero = nq->zero; ep.q8_qmin = nq->qmin; ep.q8_qmax = nq->qmax;
        ep.q8_cp = gn.Cp;
        if (workspace_is_padded(gn)) {
            const int Wp = gn.W + 2 * gn.pad;
            ep.q8_img_pixels = (gn.H + 2 * gn.pad) * Wp;
            ep.q8_row_pixels = Wp;
            ep.q8_pixel_off = gn.pad * Wp + gn.pad;
            if (int rc = launch_zero_pad_borders(ep.q8_out, gn.N, gn.H, gn.W, gn.pad, gn.Cp, st)) return rc;
        } else {
            ep.q8_img_pixels = gn.H * gn.W;
            ep.q8_row_pixels = gn.W;
        }
        ep.store_f32 = out != nullptr;
    }
    if (x_fused && dw_single_kernel(g)) return launch_conv_dw_fused(g, x_fused, wq, ep, aq, out, st);
    if (x_fused && stem_single_kernel(g, x_fused)) return launch_conv_umma(g, nullptr, wq + L.wcol_off, ep, out, st, L.Kcol, x_fused, aq);
    if (x_fused) return launch_conv_umma(g, nullptr, wq, ep, out, st, 0, x_fused, aq, false, 1, g_conv_algo == QB200_ALGO_UMMA_FUSED_QUANT);
    if (from_ws && workspace_is_im2col(g, L)) return launch_conv_umma(g, q, wq + L.wcol_off, ep, out, st, L.Kcol);
    // strided 1x1 layers read a compact buffer holding only the sampled pixels: a stride-1 conv over [N, P, Q, Cp]
    if (from_ws && workspace_is_padded(g) && L.tapKC) return launch_conv_umma(g, q, wq + L.wtap_off, ep, out, st, 0, nullptr, nullptr, true);
    const ConvGeom gk = (from_ws && uses_subsampled_input(g)) ? subsampled_geom(g) : g;
    if (resolve_algo(g) == QB200_ALGO_UMMA)
        return launch_conv_umma(gk, q, wq, ep, out, st, 0, nullptr, nullptr, false,
                                g_conv_algo == QB200_ALGO_UMMA_TWO_KERNELS ? 0 : (g_conv_algo == QB200_ALGO_UMMA_PAIR ? 2 : 1));
    return launch_conv_direct(gk, q, wq, ep, out, st);
}

}  // namespace
}  // namespace qb200

extern "C" {

int qb200_watchdog_code(void) { return qb200::watchdog_code(); }

void qb200_set_rows_chunk(int images) { qb200::g_rows_chunk = images; }
int qb200_conv_rows_chunk(const qb200_conv_shape* s) {
    if (!s || qb200::validate_shape(s)) return 0;
    return qb200::rows_chunk_images(*s, nullptr);
}

void qb200_set_conv_algo(int algo) { qb200::g_conv_algo = algo; }
int qb200_get_conv_algo(void) { return qb200::g_conv_algo; }

int qb200_conv2d_q8_nhwc(const qb200_conv_shape* s, const uint8_t* q_nhwc, const void* prepared, const float* w_scale,
                         int32_t n_w_scale, const float* bias, const qb200_act_quant* aq, void* out, int32_t out_kind,
                         void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (s->N == 0) return 0;
    return run_conv(s, q_nhwc, false, prepared, w_scale, n_w_scale, bias, aq, out, out_kind, static_cast<cudaStream_t>(stream));
}

int qb200_conv_is_single_kernel(const qb200_conv_shape* s, const float* x) {
    using namespace qb200;
    if (validate_shape(s)) return 0;
    return single_kernel(make_geom(*s), x) ? 1 : 0;
}

int qb200_conv_quantize_input(const qb200_conv_shape* s, const float* x, const qb200_act_quant* aq, void* workspace,
                              void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (s->N == 0) return 0;
    return quantize_input(s, x, aq, static_cast<uint8_t*>(workspace), static_cast<cudaStream_t>(stream));
}

int qb200_conv_from_workspace(const qb200_conv_shape* s, const void* workspace, const void* prepared, const float* w_scale,
                              int32_t n_w_scale, const float* bias, const qb200_act_quant* aq, void* out, int32_t out_kind,
                              void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (s->N == 0) return 0;
    return run_conv(s, static_cast<const uint8_t*>(workspace), true, prepared, w_scale, n_w_scale, bias, aq, out, out_kind,
                    static_cast<cudaStream_t>(stream));
}

int qb200_conv_from_workspace_ex(const qb200_conv_shape* s, const void* workspace, const void* prepared, const float* w_scale,
                                 int32_t n_w_scale, const float* bias, const qb200_act_quant* aq, const qb200_conv_tail* tail,
                                 void* out, int32_t out_kind, void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (s->N == 0) return 0;
    return run_conv(s, static_cast<const uint8_t*>(workspace), true, prepared, w_scale, n_w_scale, bias, aq, out, out_kind,
                    static_cast<cudaStream_t>(stream), nullptr, tail);
}

int qb200_conv_handoff_supported(const qb200_conv_shape* s, const qb200_conv_shape* next_shape) {
    using namespace qb200;
    if (!next_shape || validate_shape(s) || validate_shape(next_shape)) return 0;
    return handoff_ok(*s, *next_shape) ? 1 : 0;
}

int qb200_quantconv2d_fused_ex(const qb200_conv_shape* s, const float* x, const void* prepared, const float* w_scale,
                               int32_t n_w_scale, const float* bias, const qb200_act_quant* aq, const qb200_conv_tail* tail,
                               void* workspace, void* out, int32_t out_kind, void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (s->N == 0) return 0;
    QB_REQUIRE(workspace != nullptr, QB200_EINVAL, "conv: workspace missing (qb200_conv_workspace_bytes)");
    QB_REQUIRE(x != nullptr, QB200_EINVAL, "conv: null input");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // (a residual tail needs 32 more live registers in the epilogue than the 608-thread fused-quantize kernel has)
    const qb200::ConvGeom g0 = make_geom(*s);
    if (const int chunk = rows_chunk_images(*s, tail)) {
        // im2col-rows layer in L2-sized chunks of images (see g_rows_chunk): same kernels, same bits, less DRAM traffic
        const int64_t in_img = (int64_t)s->C * s->H * s->W, out_img = (int64_t)s->K * g0.P * g0.Q;
        for (int n0 = 0; n0 < s->N; n0 += chunk) {
            qb200_conv_shape sc = *s;
            sc.N = s->N - n0 < chunk ? s->N - n0 : chunk;
            qb200_conv_tail tc;
            if (tail) {
                tc = *tail;
                if (tc.residual) tc.residual += n0 * out_img;
            }
            if (int rc = quantize_input(&sc, x + n0 * in_img, aq, static_cast<uint8_t*>(workspace), st)) return rc;
            void* out_c = out_kind == QB200_OUT_ACC ? static_cast<void*>(static_cast<int32_t*>(out) + n0 * out_img)
                                                    : static_cast<void*>(static_cast<float*>(out) + n0 * out_img);
            if (int rc = run_conv(&sc, static_cast<const uint8_t*>(workspace), true, prepared, w_scale, n_w_scale, bias, aq, out_c,
                                  out_kind, st, nullptr, tail ? &tc : nullptr))
                return rc;
        }
        return 0;
    }
    if (dw_single_kernel(g0) && !(tail && tail->next_shape))   // (the depthwise kernel has the residual / ReLU tail)
        return run_conv(s, nullptr, false, prepared, w_scale, n_w_scale, bias, aq, out, out_kind, st, x, tail);
    if (single_kernel(g0, x) && !dw_single_kernel(g0) && !(tail && (tail->residual || tail->next_shape)))
        return run_conv(s, nullptr, false, prepared, w_scale, n_w_scale, bias, aq, out, out_kind, st, x, tail);
    if (int rc = quantize_input(s, x, aq, static_cast<uint8_t*>(workspace), st)) return rc;
    return run_conv(s, static_cast<const uint8_t*>(workspace), true, prepared, w_scale, n_w_scale, bias, aq, out, out_kind, st,
                    nullptr, tail);
}

int qb200_quantconv2d_fused(const qb200_conv_shape* s, const float* x, const void* prepared, const float* w_scale,
                            int32_t n_w_scale, const float* bias, const qb200_act_quant* aq, void* workspace, void* out,
                            int32_t out_kind, void* stream) {
    return qb200_quantconv2d_fused_ex(s, x, prepared, w_scale, n_w_scale, bias, aq, nullptr, workspace, out, out_kind, stream);
}

}  // extern "C"
