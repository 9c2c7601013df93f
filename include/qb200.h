/*
 * qb200.h — C-ABI of the B200 (sm_100a) quantized-operator engine.
 *
 * This is the drop-in boundary for the hot path of JingInAI/Quantize's `quant_engine` torch extension:
 *   tpack / tunpack                  reference: engine/kernels/tpack/tpack.h:17-32, tpack.cu:96-128, :327-359
 *   quantconv2d_float_input          reference: engine/kernels/functions/funcs.h:143-151,
 *                                               quantconv2d_float_input.cu:45-121 (kernel), :140-220 (host)
 * plus the activation-quantize formula the fused conv folds in
 *   Quantizer.simulate               reference: modelzoo/modules/quantizer.py:196-226 (q = clamp(round(x/s - z)))
 *
 * Everything here is plain C: device pointers, sizes, a cudaStream_t passed as void*.  No torch types.
 * The Python-facing `quant_engine` module (quantize_b200/csrc/pybind.cpp) is a thin shim over these
 * entry points; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative QB200_E* code on argument errors, or a positive
 *     cudaError_t; qb200_last_error() returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - all launches are asynchronous on `stream` (a cudaStream_t; NULL = legacy default stream).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with a CUDA error.
 */
#ifndef QB200_H_
#define QB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB200_VERSION 200

/* error codes (negative; positive values are cudaError_t) */
#define QB200_OK 0
#define QB200_EINVAL (-1)       /* bad argument */
#define QB200_EUNSUPPORTED (-2) /* valid in the reference but not implemented by this engine */
#define QB200_EDRIVER (-3)      /* CUDA driver entry point (cuTensorMapEncode*) unavailable / failed */

/* element types accepted by qb200_tpack (the reference dispatches AT_DISPATCH_ALL_TYPES_AND(Half), tpack.cu:120) */
typedef enum {
    QB200_U8 = 0,
    QB200_I8 = 1,
    QB200_I16 = 2,
    QB200_I32 = 3,
    QB200_I64 = 4,
    QB200_F16 = 5,
    QB200_F32 = 6,
    QB200_F64 = 7,
    QB200_BF16 = 8
} qb200_dtype;

int qb200_version(void);
const char* qb200_last_error(void);
/* number of this library's kernel launches issued by the calling thread since the last reset (bench.py's gpu_launches) */
uint64_t qb200_launch_count(void);
void qb200_launch_count_reset(void);

/* ------------------------------------------------------------------------------------------------
 * Sub-byte packing  (reference: tpack.cu:30-84 kernel, :96-128 launcher, :203-255 host logic)
 *
 *   stored_i = (uint8)((int8)x_i + (sign ? 2^(n_bits-1) : 0));  stream bits [i*n, (i+1)*n) = stored_i, LSB first.
 *   out must hold qb200_packed_bytes(n_elements, n_bits) bytes; every byte is written (pad bits = 0), so
 *   the caller does not have to zero it (the reference needs torch::zeros, tpack.cu:225).
 *   range_flag (device int32, may be NULL): bit 0 is OR-ed in when any element is outside
 *   [-(2^(n-1)), 2^(n-1)-1] (signed) / [0, 2^n-1] (unsigned) or is NaN — the fused form of the reference's
 *   x.min()/x.max() host check (tpack.cu:211-215).  The caller zeroes it before the call.
 * ---------------------------------------------------------------------------------------------- */
int64_t qb200_packed_bytes(int64_t n_elements, int n_bits);
int qb200_tpack(const void* x, int x_dtype, int64_t n_elements, int n_bits, int sign,
                uint8_t* out, int32_t* range_flag, void* stream);

/* reference: tpack.cu:267-315 kernel, :327-359 launcher.  out is int8 (sign) or uint8, n_elements long. */
int qb200_tunpack(const uint8_t* packed, int64_t n_elements, int n_bits, int sign,
                  void* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Quantized convolution with float input  (reference: quantconv2d_float_input.cu:45-121, :140-220)
 *
 * Shape block shared by every conv entry point.  groups = C / Cg must divide C and K; the reference op
 * only expresses groups == 1 (its signature has no groups argument, funcs.h:143-151) — groups > 1 is the
 * compatible extension used for depthwise layers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t N, C, H, W;     /* input  [N, C, H, W] fp32 NCHW contiguous                         */
    int32_t K, Cg, R, S;    /* weight [K, Cg, R, S] (weight_des[2:6]); groups = C / Cg          */
    int32_t stride, pad;    /* square, symmetric (quantconv2dop.py:82-85)                       */
    int32_t w_bits, w_sign; /* weight_des[0], weight_des[1]                                    */
} qb200_conv_shape;

/* P = (H + 2*pad - R)/stride + 1 (integer division, quantconv2d_float_input.cu:178-179) */
int qb200_conv_out_hw(const qb200_conv_shape* s, int32_t* P, int32_t* Q);

/* Channel padding of the quantized NHWC activation buffer and of the prepared weights:
 * Cp = round_up(C, 32) (TMA needs 16-B pixel strides; the MMA consumes K in 32-byte slices). */
int32_t qb200_padded_channels(int32_t C);

/* Prepared (MMA-ready) weights.  A derived, cacheable view of the reference's packed byte stream — the
 * checkpointed tensor itself is never re-laid-out.  Layout of `prepared` (all sections 256-B aligned):
 *   wq   [K][R][S][Cgp]   one byte per weight, channel-last, zero padded.  int8 when w_sign, else uint8
 *                         (unsigned weights feed the MMA as an unsigned B operand).
 *                         Cgp = qb200_padded_channels(Cg) when groups == 1, else round_up(Cg, 4).
 *   wpre int32 [K][R+1][S+1]  exclusive 2-D prefix sums over (r,s) of sum_c wq — the zero-point term of
 *                         border pixels is a 4-corner lookup (see qb200_quantconv2d_fused).
 */
size_t qb200_conv_prepared_bytes(const qb200_conv_shape* s);
int qb200_conv_prepare_weights(const qb200_conv_shape* s, const uint8_t* w_packed, void* prepared, void* stream);

/* Activation quantization parameters (Quantizer.scale/zero/qmin/qmax, quantizer.py:119-123), per-tensor.
 * DEVICE pointers to one float each, so the hot path never synchronises to read them. */
typedef struct {
    const float* scale;
    const float* zero;
    const float* qmin;
    const float* qmax;
} qb200_act_quant;

/* workspace (bytes) the fused conv needs for the quantized activations: N*H*W*Cp, or N*P*Q*Kcol for few-channel layers */
size_t qb200_conv_workspace_bytes(const qb200_conv_shape* s);

/* q = clamp(rint(x / scale - zero), qmin, qmax) as uint8, NCHW fp32 -> NHWC(Cp) u8, channels C..Cp-1 = 0
 * (quantizer.py:31, :215).  Exposed on its own for tests and for callers that keep activations quantized. */
int qb200_act_quantize_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W,
                            const qb200_act_quant* aq, uint8_t* q_nhwc, void* stream);

/* output selector */
#define QB200_OUT_F32 0 /* dequantized fp32 NCHW  (the op's contract)                               */
#define QB200_OUT_ACC 1 /* raw int32 accumulators sum(qa*qw), NCHW (parity: bit-exact vs the oracle) */

/* conv kernel selector (tests and benchmarks; 0 is what the product uses) */
#define QB200_ALGO_AUTO 0   /* tcgen05 implicit GEMM when groups == 1, CUDA-core kernel otherwise       */
#define QB200_ALGO_DIRECT 1 /* CUDA-core direct conv (any groups)                                      */
#define QB200_ALGO_UMMA 2   /* TMA im2col + tcgen05.mma kind::i8 + TMEM accumulators (groups == 1)      */
#define QB200_ALGO_UMMA_TWO_KERNELS 3 /* as 2, plain variant only: quantizer kernel + im2col-TMA conv kernel, one CTA per tile
                                       * (no fused-quantize / halo / CTA-pair variants; A/B tests) */
#define QB200_ALGO_UMMA_FUSED_QUANT 4 /* as 2, and use the fused-quantize / halo variants wherever supported (tests)    */
#define QB200_ALGO_UMMA_PAIR 5        /* as 3, but the CTA-pair (tcgen05.mma.cta_group::2) variant wherever supported (tests) */
/* Every mbarrier wait of the tensor-core kernel is bounded (~4 s): on expiry the kernel records which wait it was
 * (1 TMA producer, 2 MMA/accumulator, 3 MMA/operands, 4 epilogue, 5-7 fused-quantize producers) and traps instead of
 * hanging the GPU.  0 = never fired.  Readable after the trap (mapped host memory). */
int qb200_watchdog_code(void);
void qb200_set_conv_algo(int algo);
int qb200_get_conv_algo(void);

/* The fused hot path: activation-quantize + int8 implicit-GEMM conv + per-channel dequant + bias.
 *   out[n,k,p,q] = s_a * w_scale[k|0] * ( sum qa*qw + z_a * sum_{in-bounds taps} qw ) + bias[k]
 * which equals the reference module's packed forward (quantconv2d.py:207-210) for symmetric weights
 * (w_zero == 0, the only case the shipped configs produce, range/minmax.py:123-135).
 * w_scale: n_w_scale == 1 (per-tensor) or K (per-channel), as in quantconv2d_float_input.cu:104-106.
 * w_zero must be all-zero for this entry point (checked by the caller; see qb200_quantconv2d_weightonly).
 * bias may be NULL.  out: fp32 or int32 [N,K,P,Q] NCHW. */
int qb200_quantconv2d_fused(const qb200_conv_shape* s, const float* x, const void* prepared,
                            const float* w_scale, int32_t n_w_scale, const float* bias,
                            const qb200_act_quant* aq, void* workspace, void* out, int32_t out_kind,
                            void* stream);

/* Optional fused tail — an extension beyond the reference op for callers that own the surrounding graph (ResNet
 * blocks): out = relu(out + residual), each step rounded as the separate fp32 ops would.  residual: device fp32
 * [N,K,P,Q] or NULL; relu: 0/1.  With tail == NULL the call is exactly qb200_quantconv2d_fused.
 *
 * Quantized hand-off (the int8-out epilogue; reference counterpart: the packed-activation op
 * engine/kernels/functions/quantconv2d.cu:49-264 fed by Quantizer.pack, quantizer.py:228-246): when next_shape is set,
 * the epilogue also quantizes its result with the CONSUMER layer's activation quantizer,
 *     q = clamp(rint(v / next_quant.scale - next_quant.zero), qmin, qmax),   v = the fp32 value after the tail,
 * and writes it straight into the consumer's activation workspace (the bytes qb200_conv_quantize_input would have
 * produced from the fp32 tensor — bit-identical, the same arithmetic on the same fp32 value).  The consumer then runs
 * with qb200_conv_from_workspace(_ex).  next_shape must describe a layer whose input is this layer's output
 * (N, C == K, H == P, W == Q); `out` may be NULL when only the quantized result is wanted.
 * qb200_conv_handoff_supported tells whether this producer/consumer pair can be chained (tensor-core producer,
 * consumer workspace in NHWC or zero-padded NHWC layout); when it returns 0 the caller keeps the fp32 path. */
typedef struct {
    const float* residual;
    int32_t relu;
    const qb200_conv_shape* next_shape; /* NULL: no hand-off */
    const qb200_act_quant* next_quant;
    void* next_workspace;               /* qb200_conv_workspace_bytes(next_shape) bytes */
} qb200_conv_tail;
int qb200_quantconv2d_fused_ex(const qb200_conv_shape* s, const float* x, const void* prepared,
                               const float* w_scale, int32_t n_w_scale, const float* bias,
                               const qb200_act_quant* aq, const qb200_conv_tail* tail, void* workspace, void* out,
                               int32_t out_kind, void* stream);
int qb200_conv_handoff_supported(const qb200_conv_shape* s, const qb200_conv_shape* next_shape);

/* Few-channel layers (the RGB stem) run through materialised im2col rows; when the rows of the whole batch exceed L2,
 * qb200_quantconv2d_fused processes the batch in chunks of `images` images through the same workspace region so the rows
 * stay L2-resident (same kernels, same bits).  qb200_conv_rows_chunk: the chunk this layer will use (0 = whole batch);
 * off by default (measured slower on B200: the chunks are too small to fill the GPU); environment QB200_ROWS_CHUNK. */
void qb200_set_rows_chunk(int images);
int qb200_conv_rows_chunk(const qb200_conv_shape* s);

/* 1 when qb200_quantconv2d_fused runs this layer as ONE kernel (the quantizer runs in the conv kernel's producer warps
 * and no workspace is written), else 0.  Supported for 1x1, stride 1, pad 0, C % 64 == 0, H*W % 4 == 0, 16-byte
 * aligned x; chosen by default where it was measured to win (C == 64, feature map >= 28x28). */
int qb200_conv_is_single_kernel(const qb200_conv_shape* s, const float* x);

/* The two kernels of the fused op as separate calls (same result as qb200_quantconv2d_fused; lets a caller time or
 * overlap them).  The workspace layout is private to the pair: NHWC(Cp) bytes, or — for layers with <= 4 input
 * channels such as the RGB stem — materialised im2col rows. */
int qb200_conv_quantize_input(const qb200_conv_shape* s, const float* x, const qb200_act_quant* aq, void* workspace,
                              void* stream);
int qb200_conv_from_workspace(const qb200_conv_shape* s, const void* workspace, const void* prepared,
                              const float* w_scale, int32_t n_w_scale, const float* bias, const qb200_act_quant* aq,
                              void* out, int32_t out_kind, void* stream);
/* ... with the optional tail (residual / ReLU / quantized hand-off to the next layer) */
int qb200_conv_from_workspace_ex(const qb200_conv_shape* s, const void* workspace, const void* prepared,
                                 const float* w_scale, int32_t n_w_scale, const float* bias, const qb200_act_quant* aq,
                                 const qb200_conv_tail* tail, void* out, int32_t out_kind, void* stream);

/* Same conv on already-quantized NHWC(Cp) activations (callers that keep activations quantized). */
int qb200_conv2d_q8_nhwc(const qb200_conv_shape* s, const uint8_t* q_nhwc, const void* prepared,
                         const float* w_scale, int32_t n_w_scale, const float* bias,
                         const qb200_act_quant* aq, void* out, int32_t out_kind, void* stream);

/* Weight-only semantic of the reference op (8-argument call, no activation quantizer):
 *   out = bias + sum x * ((qw - w_zero[k|0]) * w_scale[k|0])        quantconv2d_float_input.cu:83-119
 * fp32 accumulate.  groups must be 1 (as in the reference). */
int qb200_quantconv2d_weightonly(const qb200_conv_shape* s, const float* x, const uint8_t* w_packed,
                                 const float* w_scale, const float* w_zero, int32_t n_w_scale,
                                 const float* bias, float* out, void* stream);

/* quantlinear_float_input, weight-only semantic of the reference op
 * (engine/kernels/functions/quantlinear_float_input.cu:36-106 kernel, :120-182 host):
 *   out[b, o] = ( sum_k x[b, k] * ((qw[o, k] - w_zero[o|0]) * w_scale[o|0]) ) + bias[o]
 * fp32, k ascending with fused multiply-adds and the bias added last — the reference kernel's order, so results are
 * bit-identical to it whenever in_features is a multiple of 32 (its shared tiles keep stale entries otherwise).
 * w_packed: the tpack stream of the [out_features, in_features] integer weights.
 * The activation-quantized integer path of a linear layer is the 1x1 case of the convolution: call
 * qb200_quantconv2d_fused with N = batch, C = Cg = in_features, H = W = 1, K = out_features, R = S = 1. */
int qb200_quantlinear_weightonly(const float* x, int64_t batch, int32_t in_features, int32_t out_features,
                                 const uint8_t* w_packed, int32_t w_bits, int32_t w_sign, const float* w_scale,
                                 const float* w_zero, int32_t n_w_scale, const float* bias, float* out, void* stream);

/* Quantizer.simulate of a per-tensor quantizer (modelzoo/modules/quantizer.py:194, :215-218) in one
 * pass:  out = (clamp(rint(x / scale - zero), qmin, qmax) + zero) * scale,  each step rounded like the separate fp32
 * torch kernels the reference launches (div, sub, round, clamp, add, mul) — bit-identical to them.  Any range
 * (signed ranges and ranges beyond a byte take an IEEE-division path).  x, out: n fp32 values (may alias). */
int qb200_fake_quantize_f32(const float* x, int64_t n, const qb200_act_quant* aq, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Ops whose activations arrive as a tpack'ed stream
 * ---------------------------------------------------------------------------------------------- */

/* tpack'ed NCHW activation stream (Quantizer.pack, quantizer.py:228-246, element order of tpack.cu:203-255) ->
 * NHWC(Cp) bytes holding the STORED values u = q + (sign ? 2^(n_bits-1) : 0), channels C..Cp-1 = 0: the A operand of the
 * tensor-core conv.  When in_zero / zero_adj are given the kernel also writes *zero_adj = -(offset + *in_zero), the
 * activation zero point of the integer form of quantconv2d (see qb200_quantconv2d_packed). */
int qb200_unpack_act_nhwc(const uint8_t* packed, int32_t n_bits, int32_t sign, int32_t N, int32_t C, int32_t H, int32_t W,
                          uint8_t* q_nhwc, const float* in_zero, float* zero_adj, void* stream);

/* stream -> fp32 with the reference's per-element dequantization (quantconv2d.cu:111-115, quantlinear.cu:110-112):
 *   q = (sign ? int8 : uint8)(u - offset);   out = plus_zero ? (q + zero[c]) * scale[c] : (q - zero[c]) * scale[c]
 * c = 0 when n_scale == 1, else (i / inner) % C  (per input channel of an NCHW tensor: inner = H*W). */
int qb200_dequant_packed_f32(const uint8_t* packed, int32_t n_bits, int32_t sign, int64_t n_elements, int64_t inner, int32_t C,
                             const float* scale, const float* zero, int32_t n_scale, int32_t plus_zero, float* out,
                             void* stream);

/* quantconv2d (reference engine/kernels/functions/quantconv2d.cu:49-147 kernel, :164-264 host; funcs.h:113-124) with a
 * per-tensor input quantizer and symmetric weights, as an integer GEMM on the tensor cores:
 *   out[n,k,p,q] = in_scale * w_scale[k|0] * ( sum u*qw - (offset + in_zero) * sum_{in-bounds taps} qw ) + bias[k]
 * = sum (q_in - in_zero) * in_scale * qw * w_scale + bias, the reference's (q - zero) * scale convention on both operands.
 * s describes the layer (N, C, H, W = the activation tensor the stream encodes); prepared = qb200_conv_prepare_weights.
 * workspace: qb200_quantconv2d_packed_workspace_bytes(s) bytes.  out_kind QB200_OUT_ACC returns sum u*qw (int32).
 * Per-input-channel input scales and asymmetric weights do not factor into an integer GEMM: the caller composes
 * qb200_dequant_packed_f32 + qb200_quantconv2d_weightonly for those (bit-identical to the reference kernel). */
size_t qb200_quantconv2d_packed_workspace_bytes(const qb200_conv_shape* s);
int qb200_quantconv2d_packed(const qb200_conv_shape* s, const uint8_t* in_packed, int32_t in_bits, int32_t in_sign,
                             const float* in_scale, const float* in_zero, const void* prepared, const float* w_scale,
                             int32_t n_w_scale, const float* bias, void* workspace, void* out, int32_t out_kind, void* stream);

/* quantlinear (reference engine/kernels/functions/quantlinear.cu:39-133 kernel, :231-297 host; funcs.h:37-46):
 *   out[b,o] = ( sum_k (qi[b,k] + in_zero[b]) * (qw[o,k] + w_zero[o]) * (in_scale[b] * w_scale[o]) ) + bias[o]
 * fp32, k ascending, each term rounded as the reference kernel rounds it (product, then one FMA with the scale product),
 * bias added last (0 when NULL, as the reference's wrapper substitutes zeros) — bit-identical to the reference kernel
 * whenever in_features % 32 == 0 (its shared tiles keep stale entries otherwise).  in_scale / in_zero: batch elements,
 * w_scale / w_zero: out_features elements (the reference expands 0-d tensors on the host, quantlinear.cu:275-289). */
int qb200_quantlinear_packed(const uint8_t* in_packed, int32_t in_bits, int32_t in_sign, const float* in_scale,
                             const float* in_zero, int64_t batch, int32_t in_features, int32_t out_features,
                             const uint8_t* w_packed, int32_t w_bits, int32_t w_sign, const float* w_scale, const float* w_zero,
                             const float* bias, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Calibration reductions (reference modelzoo/modules/range/minmax.py:62-108, :44-60, :184-203)
 * The tensor is viewed as [A][R][B] and reduced over A and B:  per tensor A=1,R=1,B=numel;  weights per channel
 * A=1,R=K,B=C*R*S;  activations per channel A=N,R=C,B=H*W.
 * ---------------------------------------------------------------------------------------------- */

/* out_min[r], out_max[r] = the estimator's (xmin, xmax) of row r in ONE pass over x:
 *   symmetric == 0:  (min x, max x)          symmetric != 0:  (0, max |x|)        (minmax.py:75-77, :89-91)
 * NaN propagates as in torch.min / torch.max.  update_mode folds the estimator's state update into the same call:
 *   0  none;   1  run = (min(run_min, xmin), max(run_max, xmax))                          MinMax.update   (:44-60)
 *   2  run = momentum * x + one_minus_momentum * run, rounded like the torch expression   MAMinMax.update (:184-203)
 * (modes 1, 2: run_min / run_max hold the previous state, are updated in place, and out_* receive the new state).
 * workspace: qb200_minmax_workspace_bytes(R) bytes. */
size_t qb200_minmax_workspace_bytes(int64_t R);
int qb200_minmax_f32(const float* x, int64_t A, int64_t R, int64_t B, int32_t symmetric, float* out_min, float* out_max,
                     int32_t update_mode, float momentum, float one_minus_momentum, float* run_min, float* run_max,
                     void* workspace, void* stream);

/* out[r] = the k-th smallest (1-based; of |x| when use_abs) element of row r — torch.kthvalue's value (minmax.py:78-84,
 * :92-98), exact (4-pass radix select), NaN sorts last.  1 <= k <= A*B.  workspace: qb200_kthvalue_workspace_bytes(R). */
size_t qb200_kthvalue_workspace_bytes(int64_t R);
int qb200_kthvalue_f32(const float* x, int64_t A, int64_t R, int64_t B, int32_t use_abs, int64_t k, float* out, void* workspace,
                       void* stream);

/* Max pooling over fp32 NCHW planes (planes = N*C), square kernel / stride, -inf padding, floor output size:
 * the op between the stem conv and the first residual stage of the ResNet family (torchvision resnet.py; the reference
 * runs torch.nn.MaxPool2d there).  Bit-identical to torch.nn.functional.max_pool2d; exists because that op, not a
 * conv, was the largest single kernel of the packed ResNet-50 forward. */
int qb200_maxpool2d_f32(const float* x, int64_t planes, int32_t H, int32_t W, int32_t kernel, int32_t stride,
                        int32_t pad, float* out, void* stream);

/* Global average pooling of fp32 planes: out[plane] = sum(x[plane][0..HW)) / HW — the op between the last residual stage
 * and the classifier (torchvision resnet.py: AdaptiveAvgPool2d((1, 1)); the reference runs the torch module there).
 * One warp per plane; sums in a fixed lane-strided + shuffle-tree order (deterministic, within fp32 rounding of torch). */
int qb200_avgpool_global_f32(const float* x, int64_t planes, int32_t HW, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QB200_H_ */
