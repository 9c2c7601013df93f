/*
 * qoracle.c — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference hot path, plain C.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library; it is the checker, never the product path (the product fails loudly without its CUDA library).
 *
 * Parity status: PINNED.  tpack/tunpack are checked against the packing known-answer vectors generated
 * from the reference's own compiled tpack (tests/golden/pack_kat.json) and, in this
 * container, against oracle/_ref (the unmodified reference sources compiled by oracle/build_ref.py);
 * the activation-quantize / conv / dequant chain is checked against fixtures produced by importing the
 * reference's Python modules (tests/golden/make_golden.py -> tests/golden/conv_*.npz).
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference).
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp (oracle/build_oracle.py). -ffp-contract=off
 * matters: the fp32 expressions below must round after every operation exactly as the reference's
 * separate torch ops / non-contracted CUDA statements do.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---- engine/kernels/tpack/tpack.cu:203-255 (host) + :30-84 (kernel): pack ----------------------
 * returns 0, or 1 when a value is outside the representable range (tpack.cu:211-215 raises). */
int qo_tpack_f32(const float* x, int64_t n, int n_bits, int sign, uint8_t* out /* ceil(n*bits/8), any content */)
{
    const int64_t n_out = (n * n_bits + 7) / 8;           /* tpack.cu:224 */
    const float lo = sign ? -(float)(1 << (n_bits - 1)) : 0.0f;
    const float hi = sign ? (float)((1 << (n_bits - 1)) - 1) : (float)((1 << n_bits) - 1);
    const uint8_t offset = sign ? (uint8_t)(1 << (n_bits - 1)) : 0;   /* tpack.cu:106-110 */
    int bad = 0;
    memset(out, 0, (size_t)n_out);                        /* tpack.cu:225 torch::zeros */
    for (int64_t i = 0; i < n; ++i) {
        if (!(x[i] >= lo && x[i] <= hi)) bad = 1;         /* tpack.cu:211-215 via min()/max() */
        uint8_t el = (uint8_t)(int8_t)(int)x[i];          /* tpack.cu:50  (char)x[index]      */
        el = (uint8_t)(el + offset);                      /* tpack.cu:51                      */
        el &= (uint8_t)((1 << n_bits) - 1);               /* in range this is a no-op          */
        const int64_t bit = i * n_bits;                   /* tpack.cu:54                      */
        const int64_t byte = bit / 8;                     /* tpack.cu:57                      */
        const int off = (int)(bit % 8);                   /* tpack.cu:60                      */
        out[byte] |= (uint8_t)(el << off);                /* tpack.cu:63-69                   */
        if (off + n_bits > 8)                             /* tpack.cu:71-81                   */
            out[byte + 1] |= (uint8_t)(el >> (8 - off));
    }
    return bad;
}

/* ---- tpack.cu:429-476 (host) + :267-315 (kernel): unpack -------------------------------------- */
void qo_tunpack(const uint8_t* packed, int64_t n, int n_bits, int sign, uint8_t* out /* int8 bit patterns if sign */)
{
    const uint8_t offset = sign ? (uint8_t)(1 << (n_bits - 1)) : 0;
    const uint8_t mask = (uint8_t)((1 << n_bits) - 1);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t bit = i * n_bits;
        const int64_t byte = bit / 8;
        const int off = (int)(bit % 8);
        uint8_t el = (uint8_t)((packed[byte] >> off) & mask);              /* tpack.cu:296 */
        if (off + n_bits > 8)                                              /* tpack.cu:298-305 */
            el |= (uint8_t)((packed[byte + 1] << (8 - off)) & mask);
        out[i] = (uint8_t)(el - offset);                                   /* tpack.cu:308-309 */
    }
}

/* ---- modelzoo/modules/quantizer.py:31 (Round.forward), :215 (simulate): activation quantize ----
 * q = clamp(round_half_even(x / scale - zero), qmin, qmax), all fp32 like the torch ops. */
void qo_act_quantize(const float* x, int64_t n, float scale, float zero, float qmin, float qmax, float* q)
{
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) {
        float t = x[i] / scale;      /* x/scale            */
        t = t - zero;                /* - zero             */
        t = nearbyintf(t);           /* .round(): half-even */
        t = t < qmin ? qmin : t;     /* .clamp(qmin, qmax) */
        t = t > qmax ? qmax : t;
        q[i] = t;
    }
}

/* ---- exact integer convolution of the quantized operands --------------------------------------
 * acc[n,k,p,q] = sum_{c,r,s in bounds} qa[n, g*Cg + c, ih, iw] * qw[k, c, r, s]
 * Index arithmetic follows quantconv2d_float_input.cu:86-94 (ih = oh*stride - padding + kh, bounds
 * check :92, OIHW weight index :94); groups follow torch conv semantics (quantconv2d.py:166-168 →
 * nn.Conv2d._conv_forward) because the reference op itself cannot express groups.
 * Also returns wsum[k,p,q] = sum over the same in-bounds taps of qw (the zero-point term). */
void qo_conv_acc(const uint8_t* qa, const int8_t* qw, int N, int C, int H, int W, int K, int Cg, int R, int S,
                 int stride, int pad, int32_t* acc, int32_t* wsum /* [K,P,Q] or NULL */)
{
    const int P = (H + 2 * pad - R) / stride + 1;          /* quantconv2d_float_input.cu:178 */
    const int Q = (W + 2 * pad - S) / stride + 1;          /* :179 */
    const int groups = C / Cg;
    const int Kg = K / groups;
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const int g = k / Kg;
            for (int p = 0; p < P; ++p)
                for (int q = 0; q < Q; ++q) {
                    int64_t a = 0;
                    int32_t ws = 0;
                    for (int c = 0; c < Cg; ++c)
                        for (int r = 0; r < R; ++r)
                            for (int s = 0; s < S; ++s) {
                                const int ih = p * stride - pad + r;
                                const int iw = q * stride - pad + s;
                                if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
                                    const int32_t wv = qw[((k * Cg + c) * R + r) * S + s];
                                    a += (int32_t)qa[(((int64_t)n * C + g * Cg + c) * H + ih) * W + iw] * wv;
                                    ws += wv;
                                }
                            }
                    acc[(((int64_t)n * K + k) * P + p) * Q + q] = (int32_t)a;
                    if (wsum && n == 0) wsum[(k * P + p) * Q + q] = ws;
                }
        }
}

/* ---- integer form of the packed forward (quantconv2d.py:207-210 with w_zero == 0) -------------
 * out = s_a*s_w[k] * (acc + z_a * wsum_valid) + bias[k]   (SURVEY Appendix A "INT8 FORM") */
void qo_dequant(const int32_t* acc, const int32_t* wsum, int N, int K, int P, int Q, float s_a, float z_a,
                const float* s_w, int n_sw, const float* bias, float* out)
{
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float sw = s_w[n_sw == 1 ? 0 : k];
            const float b = bias ? bias[k] : 0.0f;
            for (int i = 0; i < P * Q; ++i) {
                const int64_t o = ((int64_t)n * K + k) * P * Q + i;
                /* double here on purpose: this is the mathematically exact value the fp32 kernel is
                 * allowed to deviate from by the 1e-3 relative tolerance of the north star */
                const double t = (double)acc[o] + (double)z_a * (double)wsum[(int64_t)k * P * Q + i];
                out[o] = (float)((double)s_a * (double)sw * t + (double)b);
            }
        }
}

/* ---- engine/kernels/functions/quantconv2d_float_input.cu:45-121: the op's weight-only semantic --
 * fp32 sequential accumulate in the kernel's loop order (ic -> kh -> kw), bias first (:83). */
void qo_quantconv2d_float_input(const float* x, const uint8_t* w_packed, const float* w_scale, const float* w_zero,
                                int per_tensor, int n_bits, int sign, const float* bias, float* out,
                                int N, int C, int H, int W, int K, int R, int S, int stride, int pad)
{
    const int P = (H + 2 * pad - R) / stride + 1;
    const int Q = (W + 2 * pad - S) / stride + 1;
    const uint8_t offset = sign ? (uint8_t)(1 << (n_bits - 1)) : 0;   /* :185 */
    const uint8_t mask = (uint8_t)((1 << n_bits) - 1);
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            for (int p = 0; p < P; ++p)
                for (int q = 0; q < Q; ++q) {
                    float o = bias ? bias[k] : 0.0f;                                   /* :83 */
                    for (int c = 0; c < C; ++c)                                        /* :86 */
                        for (int r = 0; r < R; ++r)                                    /* :87 */
                            for (int s = 0; s < S; ++s) {                              /* :88 */
                                const int ih = p * stride - pad + r;                   /* :89 */
                                const int iw = q * stride - pad + s;                   /* :90 */
                                if (ih >= 0 && ih < H && iw >= 0 && iw < W) {          /* :92 */
                                    const int64_t e = (((int64_t)k * C + c) * R + r) * S + s;  /* :94 */
                                    const int64_t byte = e * n_bits / 8;               /* :95 */
                                    const int bit = (int)(e * n_bits % 8);             /* :96 */
                                    uint8_t v = (uint8_t)((w_packed[byte] >> bit) & mask);      /* :97 */
                                    if (bit + n_bits > 8)                              /* :98-99 */
                                        v |= (uint8_t)((w_packed[byte + 1] << (8 - bit)) & mask);
                                    v = (uint8_t)(v - offset);                         /* :102 */
                                    const float wv = sign ? (float)(int8_t)v : (float)v;       /* :103 */
                                    const float wf = per_tensor ? (wv - w_zero[0]) * w_scale[0]
                                                                : (wv - w_zero[k]) * w_scale[k];   /* :104-106 */
                                    const float xv = x[(((int64_t)n * C + c) * H + ih) * W + iw];  /* :109 */
                                    o = fmaf(xv, wf, o);   /* :112 — nvcc's default -fmad contracts `+= a*b` to FFMA */
                                }
                            }
                    out[(((int64_t)n * K + k) * P + p) * Q + q] = o;                  /* :119 */
                }
}

/* ---- engine/kernels/functions/quantlinear_float_input.cu:36-106: weight-only linear -----------------
 * acc = 0; acc += x[b,k] * wf[o,k] for k ascending (:94-96; nvcc contracts to FFMA); out = acc + bias (:103-104).
 * The reference kernel walks 32-wide shared tiles and, when input_size is not a multiple of 32, multiplies stale tile
 * entries of the last pass (:66-68, :71 guard the loads, :94 does not guard the loop); this restatement sums exactly the
 * input_size terms, so it equals the reference for input_size % 32 == 0. */
void qo_quantlinear_float_input(const float* x, const uint8_t* w_packed, const float* w_scale, const float* w_zero,
                                int per_tensor, int n_bits, int sign, const float* bias, float* out,
                                int batch, int in_f, int out_f)
{
    const uint8_t offset = sign ? (uint8_t)(1 << (n_bits - 1)) : 0;   /* :163 */
    const uint8_t mask = (uint8_t)((1 << n_bits) - 1);
#pragma omp parallel for collapse(2)
    for (int b = 0; b < batch; ++b)
        for (int o = 0; o < out_f; ++o) {
            float acc = 0.0f;                                                     /* :60 */
            for (int k = 0; k < in_f; ++k) {
                const int64_t e = (int64_t)o * in_f + k;                          /* :73 */
                const int64_t byte = e * n_bits / 8;                              /* :74 */
                const int bit = (int)(e * n_bits % 8);                            /* :75 */
                uint8_t v = (uint8_t)((w_packed[byte] >> bit) & mask);            /* :76 */
                if (bit + n_bits > 8) v |= (uint8_t)((w_packed[byte + 1] << (8 - bit)) & mask);   /* :77-78 */
                v = (uint8_t)(v - offset);                                        /* :81 */
                const float wv = sign ? (float)(int8_t)v : (float)v;              /* :82 */
                const float wf = per_tensor ? (wv - w_zero[0]) * w_scale[0] : (wv - w_zero[o]) * w_scale[o];   /* :83-85 */
                acc = fmaf(x[(int64_t)b * in_f + k], wf, acc);                    /* :95 */
            }
            out[(int64_t)b * out_f + o] = acc + (bias ? bias[o] : 0.0f);          /* :104 */
        }
}

/* ---- engine/kernels/functions/quantconv2d.cu:49-147: packed activations x packed weights ---------------
 * Both operands are tpack streams; each MAC dequantizes both with (q - zero) * scale (:113-115, :127-129) and
 * accumulates in fp32 in the order inc -> keh -> kew, bias first (:95-134; `+= a * b` contracted to FFMA by nvcc). */
static inline uint8_t qo_stream_get(const uint8_t* s, int64_t e, int n_bits)
{
    const uint8_t mask = (uint8_t)((1 << n_bits) - 1);
    const int64_t byte = e * n_bits / 8;
    const int bit = (int)(e * n_bits % 8);
    uint8_t v = (uint8_t)((s[byte] >> bit) & mask);
    if (bit + n_bits > 8) v |= (uint8_t)((s[byte + 1] << (8 - bit)) & mask);
    return v;
}

void qo_quantconv2d(const uint8_t* in_packed, int in_bits, int in_sign, const float* in_scale, const float* in_zero,
                    int in_per_tensor, const uint8_t* w_packed, int w_bits, int w_sign, const float* w_scale,
                    const float* w_zero, int w_per_tensor, const float* bias, float* out,
                    int N, int C, int H, int W, int K, int R, int S, int stride, int pad)
{
    const int P = (H + 2 * pad - R) / stride + 1;                       /* :219 */
    const int Q = (W + 2 * pad - S) / stride + 1;                       /* :220 */
    const uint8_t ioff = in_sign ? (uint8_t)(1 << (in_bits - 1)) : 0;   /* :227 */
    const uint8_t woff = w_sign ? (uint8_t)(1 << (w_bits - 1)) : 0;     /* :228 */
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            for (int p = 0; p < P; ++p)
                for (int q = 0; q < Q; ++q) {
                    float o = bias ? bias[k] : 0.0f;                                              /* :92 */
                    for (int c = 0; c < C; ++c)
                        for (int r = 0; r < R; ++r)
                            for (int s = 0; s < S; ++s) {
                                const int ih = p * stride + r - pad, iw = q * stride + s - pad;   /* :98-99 */
                                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;             /* :101 */
                                uint8_t a = qo_stream_get(in_packed, (((int64_t)n * C + c) * H + ih) * W + iw, in_bits);
                                a = (uint8_t)(a - ioff);                                          /* :111 */
                                const float af = in_sign ? (float)(int8_t)a : (float)a;           /* :112 */
                                const float xf = in_per_tensor ? (af - in_zero[0]) * in_scale[0]
                                                               : (af - in_zero[c]) * in_scale[c]; /* :113-115 */
                                uint8_t w = qo_stream_get(w_packed, (((int64_t)k * C + c) * R + r) * S + s, w_bits);
                                w = (uint8_t)(w - woff);                                          /* :125 */
                                const float wv = w_sign ? (float)(int8_t)w : (float)w;            /* :126 */
                                const float wf = w_per_tensor ? (wv - w_zero[0]) * w_scale[0]
                                                              : (wv - w_zero[k]) * w_scale[k];    /* :127-129 */
                                o = fmaf(xf, wf, o);                                              /* :132 */
                            }
                    out[(((int64_t)n * K + k) * P + p) * Q + q] = o;                              /* :139 */
                }
}

/* integer accumulators of the same op: sum of STORED activation values (u = q + offset) times signed weights over the
 * in-bounds taps — what the tensor-core path accumulates before its epilogue folds -(offset + zero) * sum(qw) in */
void qo_quantconv2d_acc(const uint8_t* in_packed, int in_bits, const uint8_t* w_packed, int w_bits, int w_sign, int32_t* acc,
                        int N, int C, int H, int W, int K, int R, int S, int stride, int pad)
{
    const int P = (H + 2 * pad - R) / stride + 1, Q = (W + 2 * pad - S) / stride + 1;
    const uint8_t woff = w_sign ? (uint8_t)(1 << (w_bits - 1)) : 0;
#pragma omp parallel for collapse(2)
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k)
            for (int p = 0; p < P; ++p)
                for (int q = 0; q < Q; ++q) {
                    int32_t o = 0;
                    for (int c = 0; c < C; ++c)
                        for (int r = 0; r < R; ++r)
                            for (int s = 0; s < S; ++s) {
                                const int ih = p * stride + r - pad, iw = q * stride + s - pad;
                                if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                                const int32_t u = qo_stream_get(in_packed, (((int64_t)n * C + c) * H + ih) * W + iw, in_bits);
                                const uint8_t w = (uint8_t)(qo_stream_get(w_packed, (((int64_t)k * C + c) * R + r) * S + s, w_bits) - woff);
                                o += u * (w_sign ? (int32_t)(int8_t)w : (int32_t)w);
                            }
                    acc[(((int64_t)n * K + k) * P + p) * Q + q] = o;
                }
}

/* ---- engine/kernels/functions/quantlinear.cu:39-133: packed x packed linear ----------------------------
 * tmp += (qi + in_zero[row]) * (qw + w_zero[col]) * (in_scale[row] * w_scale[col]) for k ascending (:110-120: the product
 * of the two operands is rounded, then multiplied by the scale product and added — nvcc contracts that last multiply-add
 * to one FFMA), out = tmp + bias (:127).  Sums exactly input_size terms (the reference's 32-wide tiles multiply stale
 * entries when input_size % 32 != 0, see qo_quantlinear_float_input). */
void qo_quantlinear(const uint8_t* in_packed, int in_bits, int in_sign, const float* in_scale, const float* in_zero,
                    const uint8_t* w_packed, int w_bits, int w_sign, const float* w_scale, const float* w_zero,
                    const float* bias, float* out, int batch, int in_f, int out_f)
{
    const uint8_t ioff = in_sign ? (uint8_t)(1 << (in_bits - 1)) : 0;
    const uint8_t woff = w_sign ? (uint8_t)(1 << (w_bits - 1)) : 0;
#pragma omp parallel for collapse(2)
    for (int b = 0; b < batch; ++b)
        for (int o = 0; o < out_f; ++o) {
            float tmp = 0.0f;                                                               /* :79 */
            const float sc = in_scale[b] * w_scale[o];                                      /* :96 */
            for (int k = 0; k < in_f; ++k) {
                const uint8_t a = (uint8_t)(qo_stream_get(in_packed, (int64_t)b * in_f + k, in_bits) - ioff);   /* :110 */
                const float av = (in_sign ? (float)(int8_t)a : (float)a) + in_zero[b];      /* :111-112 */
                const uint8_t w = (uint8_t)(qo_stream_get(w_packed, (int64_t)o * in_f + k, w_bits) - woff);     /* :115 */
                const float wv = (w_sign ? (float)(int8_t)w : (float)w) + w_zero[o];        /* :116-117 */
                const float prod = av * wv;
                tmp = fmaf(prod, sc, tmp);                                                  /* :120 */
            }
            out[(int64_t)b * out_f + o] = tmp + (bias ? bias[o] : 0.0f);                    /* :127 */
        }
}
