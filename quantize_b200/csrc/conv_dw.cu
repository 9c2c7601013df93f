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
#include <cstdlib>
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
    // shared memory after the patch: the filters [CH][taps] and, for KS kernels, the in-bounds weight sums of the 16
    // window classes [CH][16]: class = (first row missing) | (last row missing) << 1 | (first column missing) << 2 |
    // (last column missing) << 3  (KS <= 3 with pad <= 1: at most one row / column is missing on each side)
    int* wsm = reinterpret_cast<int*>(patch + (((size_t)CH * in_h * in_w_alloc + 15) & ~(size_t)15));
    int* wcls = wsm + CH * taps;
    for (int i = threadIdx.x; i < nch * taps; i += kDwThreads) {
        const uint8_t b = wq[((int64_t)k0 * taps + i) * g.Cgp];
        wsm[i] = kSignedW ? (int)(int8_t)b : (int)b;
    }
    const bool cls_ok = KS == 3 && g.pad <= 1;            // the class table describes every window
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // narrow planes (Q < 32: 14x14, 7x7): lanes run over the flattened (row, column) index of a channel's band instead,
    // so that all 32 lanes work; the division by the row length is a multiply-shift (exact for indices < 1024, divisors < 64)
    const bool flat = g.Q < 32 && in_h * in_w < 1024;
    const uint32_t magic_in = (65536u + (uint32_t)in_w - 1u) / (uint32_t)in_w, magic_q = (65536u + (uint32_t)g.Q - 1u) / (uint32_t)g.Q;
    auto quant1 = [&](float v) -> uint8_t {
        if (p.byte_clamp) return (uint8_t)min(max(quant_int(v, p), p.ilo), p.ihi);
        return (uint8_t)(quant_word_exact(v, 0.f, 0.f, 0.f, p.s, p.z, p.lo, p.hi) & 0xFFu);
    };
    // ---- load + quantize.  All loads of a pass are issued before the first value is used (a load -> quantize -> store
    //      chain per element left one HBM round trip per element on the critical path of every warp). ----
    if (flat) {
        for (int ch = warp; ch < nch; ch += kDwWarps) {
            const int c = (k0 + ch) / mult;
            const float* xc = x + ((int64_t)n * g.C + c) * g.H * g.W;
            uint8_t* pch = patch + (size_t)ch * in_h * in_w_alloc;
            for (int i0 = 0; i0 < in_h * in_w; i0 += 128) {
                float v[4];
                int off[4];
                bool inb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    const int r = (int)(((uint32_t)i * magic_in) >> 16), col = i - r * in_w;
                    const int ih = ih0 + r, iw = iw0 + col;
                    inb[u] = i < in_h * in_w && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
                    off[u] = i < in_h * in_w ? r * in_w_alloc + col : -1;
                    v[u] = inb[u] ? __ldg(xc + ih * g.W + iw) : 0.f;
                }
                // pixels outside the image are 0 (padded taps add nothing); an in-image NaN goes through the quantizer
                // like everywhere else in the engine (-> qmin)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (off[u] >= 0) pch[off[u]] = inb[u] ? quant1(v[u]) : (uint8_t)0;
            }
        }
    } else {
        for (int pr = warp; pr < nch * in_h; pr += kDwWarps) {
            const int ch = pr / in_h, r = pr - ch * in_h;
            const int ih = ih0 + r;
            const int c = (k0 + ch) / mult;
            uint8_t* prow = patch + ((size_t)ch * in_h + r) * in_w_alloc;
            const bool rok = ih >= 0 && ih < g.H;
            const float* xr = x + (((int64_t)n * g.C + c) * g.H + (rok ? ih : 0)) * g.W;
            for (int c0 = 0; c0 < in_w; c0 += 128) {
                float v[4];
                bool in[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col = c0 + u * 32 + lane, iw = iw0 + col;
                    in[u] = rok && col < in_w && iw >= 0 && iw < g.W;
                    v[u] = in[u] ? __ldg(xr + iw) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col = c0 + u * 32 + lane;
                    if (col < in_w) prow[col] = in[u] ? quant1(v[u]) : (uint8_t)0;
                }
            }
        }
    }
    if (cls_ok) {
        for (int i = threadIdx.x; i < nch * 16; i += kDwThreads) {
            const int ch = i >> 4, cls = i & 15;
            const int r_lo = cls & 1, r_hi = 3 - ((cls >> 1) & 1), c_lo = (cls >> 2) & 1, c_hi = 3 - ((cls >> 3) & 1);
            int sum = 0;
            for (int r = r_lo; r < r_hi; ++r)
                for (int s2 = c_lo; s2 < c_hi; ++s2) {
                    const uint8_t b = wq[((int64_t)(k0 + ch) * taps + r * 3 + s2) * g.Cgp];
                    sum += kSignedW ? (int)(int8_t)b : (int)b;
                }
            wcls[i] = sum;
        }
    }
    __syncthreads();

    // ---- stencil + dequant: a warp takes (channel, output row) pairs (or whole channel bands when `flat`) ----
    const EpilogueScalars es = load_epilogue_scalars(ep);
    const bool acc_out = ep.out_kind == QB200_OUT_ACC;
    const int n_pr = flat ? nch : nch * n_rows;
    for (int pr = warp; pr < n_pr; pr += kDwWarps) {
        const int ch = flat ? pr : pr / n_rows;
        const int orow_fixed = flat ? 0 : pr - ch * n_rows;
        const int k = k0 + ch;
        const float scale = __fmul_rn(es.s_a, __ldg(ep.w_scale + (ep.per_tensor_w ? 0 : k)));
        const float bias = ep.bias ? __ldg(ep.bias + k) : 0.f;
        const int* wk = wsm + ch * taps;
        const int* wc = wcls + ch * 16;
        const uint8_t* pch = patch + (size_t)ch * in_h * in_w_alloc;
        // outputs of a channel band are contiguous: idx = base + orow * Q + ow  (= base + item index when flat)
        const int64_t base = (((int64_t)n * g.K + k) * g.P + oh0 + orow_fixed) * g.Q;
        int w9[KS ? KS * KS : 1];
        uint32_t wpack[KS ? KS : 1];   // a filter row as the 4 bytes of a dp4a operand (KS <= 4)
        if (KS) {
#pragma unroll
            for (int i = 0; i < KS * KS; ++i) w9[i] = wk[i];
#pragma unroll
            for (int r = 0; r < KS; ++r) {
                wpack[r] = 0;
#pragma unroll
                for (int s2 = 0; s2 < KS; ++s2) wpack[r] |= ((uint32_t)w9[r * KS + s2] & 0xFFu) << (8 * s2);
            }
        }
        const int n_items = flat ? n_rows * g.Q : g.Q;
        for (int it = lane; it < n_items; it += 32) {
            int orow = orow_fixed, ow = it;
            if (flat) {
                orow = (int)(((uint32_t)it * magic_q) >> 16);
                ow = it - orow * g.Q;
            }
            const int h_in = (oh0 + orow) * g.stride - g.pad, w_in = ow * g.stride - g.pad;
            int acc = 0, ws = 0;
            if (KS) {
                // a row's KS taps = one dp4a on the 4 bytes starting at the output's column (two aligned words + a funnel
                // shift; the byte past the filter meets a zero weight)
                const int col0 = ow * g.stride;
                const uint32_t* row32 = reinterpret_cast<const uint32_t*>(pch + (size_t)(orow * g.stride) * in_w_alloc) + (col0 >> 2);
                const int sh = (col0 & 3) * 8;
#pragma unroll
                for (int r = 0; r < KS; ++r) {
                    const uint32_t* rr = row32 + r * (in_w_alloc >> 2);
                    const uint32_t a = __funnelshift_r(rr[0], rr[1], sh);
                    if (kSignedW) asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(wpack[r]));
                    else asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc) : "r"(a), "r"(wpack[r]));
                }
                if (cls_ok) {
                    const int cls = (h_in < 0 ? 1 : 0) | (h_in + KS > g.H ? 2 : 0) | (w_in < 0 ? 4 : 0) | (w_in + KS > g.W ? 8 : 0);
                    ws = wc[cls];
                } else {
#pragma unroll
                    for (int r = 0; r < KS; ++r)
#pragma unroll
                        for (int s2 = 0; s2 < KS; ++s2)
                            ws += (h_in + r >= 0 && h_in + r < g.H && w_in + s2 >= 0 && w_in + s2 < g.W) ? w9[r * KS + s2] : 0;
                }
            } else {
                const uint8_t* pp = pch + (size_t)(orow * g.stride) * in_w_alloc + ow * g.stride;
                for (int r = 0; r < g.R; ++r) {
                    const bool rok = h_in + r >= 0 && h_in + r < g.H;
                    for (int s2 = 0; s2 < g.S; ++s2) {
                        const int w = wk[r * g.S + s2];
                        acc += (int)pp[r * in_w_alloc + s2] * w;           // out-of-image patch entries are 0
                        ws += (rok && w_in + s2 >= 0 && w_in + s2 < g.W) ? w : 0;
                    }
                }
            }
            const int64_t idx = base + it;
            if (acc_out) {
                static_cast<int32_t*>(out)[idx] = acc;
            } else {
                float t = (float)acc;
                if (es.z_a != 0.f) t = __fmaf_rn(es.z_a, (float)ws, t);
                static_cast<float*>(out)[idx] = epilogue_tail(__fmaf_rn(scale, t, bias), idx, ep);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Streaming form for 3x3 / pad 1 / stride 1 or 2 depthwise layers with one filter per channel (MobileNetV2's): nothing is
// staged in shared memory.  A lane owns one column and walks down the plane; per input row it loads its fp32 value(s),
// quantizes them in registers (the engine's exact quantizer), gets the neighbouring columns by shuffle, and adds the three
// filter rows' dot products to rolling accumulators of the three output rows the input row belongs to — every input
// element is read once, every output written once, the loads of kDwRows input rows are in flight before the first use.
// The band kernel above quantized into a shared patch and read every byte back nine times through funnel shifts: 70 %
// issue-active at 0.8-1.2 TB/s (profiles/r01_ncu_full_prof_dw.csv).
// Lane groups of G lanes (8 / 16 / 32, the shuffle width) cover one (plane, run of columns): G - 2 outputs per group at
// stride 1 (the first and last lane only supply neighbours), G - 1 at stride 2 (lane j owns output column c0 - 1 + j and
// input columns 2 * that and + 1; lane 0 supplies the left neighbour); 32 / G planes per warp for narrow planes.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef QB200_DW_ROWS1
#define QB200_DW_ROWS1 8
#endif
#ifndef QB200_DW_ROWS2
#define QB200_DW_ROWS2 8
#endif

template <bool kSignedW, int STRIDE, int G>
__global__ void __launch_bounds__(256)
conv_dw3_stream_kernel(const float* __restrict__ x, const uint8_t* __restrict__ wq, ConvGeom g, EpilogueParams ep,
                       void* __restrict__ out, int groups_per_plane, int64_t items,
                       const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                       const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int kDwRows = STRIDE == 1 ? QB200_DW_ROWS1 : QB200_DW_ROWS2;   // input rows per batch of loads
    constexpr int kPPW = 32 / G;                       // planes per warp
    constexpr int kOut = STRIDE == 1 ? G - 2 : G - 1;  // output columns per group
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), sub = lane / G;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (set of kPPW planes, column group)
    if (item >= items) return;
    const int64_t planes = (int64_t)g.N * g.K;
    const int64_t pset = item / groups_per_plane;
    const int cg = (int)(item - pset * groups_per_plane);
    const int64_t plane = pset * kPPW + sub;
    const bool plane_ok = plane < planes;
    const int k = plane_ok ? (int)(plane % g.K) : 0;
    const int oc = cg * kOut - 1 + gl;                  // this lane's output column (stride 1: also its input column)
    const bool emit = plane_ok && oc >= 0 && oc < g.Q && (STRIDE == 1 ? (gl >= 1 && gl <= G - 2) : gl >= 1);
    const int ic = STRIDE * oc;                         // first input column of the lane
    const bool in0 = plane_ok && ic >= 0 && ic < g.W, in1 = STRIDE == 2 && plane_ok && ic >= 0 && ic + 1 < g.W;
    const float* xp = x + (plane_ok ? plane : 0) * g.H * g.W + (in0 ? ic : 0);
    // the filter (value = byte 0 of each tap's padded channel group) and the in-bounds column sums of its rows
    int w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint8_t b = __ldg(wq + ((int64_t)k * 9 + i) * g.Cgp);
        w[i] = kSignedW ? (int)(int8_t)b : (int)b;
    }
    const bool has_l = STRIDE * oc - 1 >= 0, has_r = STRIDE * oc + 1 < g.W;   // left / right tap column inside the image
    int cs[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) cs[r] = (has_l ? w[3 * r] : 0) + w[3 * r + 1] + (has_r ? w[3 * r + 2] : 0);
    const EpilogueScalars es = load_epilogue_scalars(ep);
    const float scale = __fmul_rn(es.s_a, __ldg(ep.w_scale + (ep.per_tensor_w ? 0 : k)));
    const float bias = ep.bias ? __ldg(ep.bias + k) : 0.f;
    const bool acc_out = ep.out_kind == QB200_OUT_ACC;
    const int64_t obase = (plane_ok ? plane : 0) * g.P * g.Q + (emit ? oc : 0);
    auto store = [&](int prow, int acc) {     // output row prow of this lane's column
        if (!emit) return;
        const int64_t idx = obase + (int64_t)prow * g.Q;
        if (acc_out) {
            static_cast<int32_t*>(out)[idx] = acc;
        } else {
            float t = (float)acc;
            if (es.z_a != 0.f) {
                const int h0 = prow * STRIDE - 1;
                const int ws = (h0 >= 0 ? cs[0] : 0) + cs[1] + (h0 + 2 < g.H ? cs[2] : 0);
                t = __fmaf_rn(es.z_a, (float)ws, t);
            }
            static_cast<float*>(out)[idx] = epilogue_tail(__fmaf_rn(scale, t, bias), idx, ep);
        }
    };
    auto quant = [&](float v) -> int {
        if (p.byte_clamp) return quant_int(v, p);       // tight input clamp: already within [qmin, qmax]
        return (int)(quant_word_exact(v, 0.f, 0.f, 0.f, p.s, p.z, p.lo, p.hi) & 0xFFu);
    };
    int acc_a = 0, acc_b = 0;   // stride 1: partial sums of output rows ih - 1 and ih; stride 2: acc_b = row 2p - 1's part of output p
    for (int h0 = 0; h0 < g.H; h0 += kDwRows) {
        float v0[kDwRows], v1[STRIDE == 2 ? kDwRows : 1];
#pragma unroll
        for (int u = 0; u < kDwRows; ++u) {
            const bool rok = h0 + u < g.H;
            v0[u] = 0.f;
            if (STRIDE == 2) {
                v1[u] = 0.f;
                if (rok && in1) {
                    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v0[u]), "=f"(v1[u]) : "l"(xp + (int64_t)(h0 + u) * g.W));
                } else if (rok && in0) {
                    v0[u] = __ldg(xp + (int64_t)(h0 + u) * g.W);
                }
            } else if (rok && in0) {
                asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v0[u]) : "l"(xp + (int64_t)(h0 + u) * g.W));
            }
        }
#pragma unroll
        for (int u = 0; u < kDwRows; ++u) {
            const int ih = h0 + u;
            if (ih >= g.H) break;                       // (uniform)
            // quantized taps of this input row: pixels outside the image are 0 (they add nothing, and the zero-point term
            // counts in-bounds taps only)
            if (STRIDE == 1) {
                const int a = in0 ? quant(v0[u]) : 0;
                const int l = __shfl_up_sync(0xffffffffu, a, 1, G), r = __shfl_down_sync(0xffffffffu, a, 1, G);
                const int t0 = w[0] * l + w[1] * a + w[2] * r, t1 = w[3] * l + w[4] * a + w[5] * r, t2 = w[6] * l + w[7] * a + w[8] * r;
                if (ih >= 1) store(ih - 1, acc_a + t2);
                acc_a = acc_b + t1;
                acc_b = t0;
            } else {
                const int a0 = in0 ? quant(v0[u]) : 0, a1 = in1 ? quant(v1[u]) : 0;
                const int l = __shfl_up_sync(0xffffffffu, a1, 1, G);
                if ((ih & 1) == 0) {                    // row 2p: filter row 1 of output p
                    acc_a = acc_b + w[3] * l + w[4] * a0 + w[5] * a1;
                } else {                                // row 2p + 1: filter row 2 of output p, filter row 0 of output p + 1
                    store(ih >> 1, acc_a + w[6] * l + w[7] * a0 + w[8] * a1);
                    acc_b = w[0] * l + w[1] * a0 + w[2] * a1;
                }
            }
        }
    }
    if (STRIDE == 1) store(g.H - 1, acc_a);             // the last output row has no row below it
    else if (g.H & 1) store(g.H >> 1, acc_a);           // odd H: output row (H - 1) / 2 ends on the last input row
}

}  // namespace

bool dw_fused_supported(const ConvGeom& g) { return g.groups == g.C && g.Cg == 1 && g.groups > 1 && g.K <= 65535 && g.N <= 65535; }

int launch_conv_dw_fused(const ConvGeom& g, const float* x, const uint8_t* wq, const EpilogueParams& ep,
                         const qb200_act_quant* aq, void* out, cudaStream_t st) {
    QB_REQUIRE(dw_fused_supported(g), QB200_EUNSUPPORTED, "conv_dw: not a depthwise layer");
    QB_REQUIRE(aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL, "conv_dw: activation quantizer parameters missing");
    QB_REQUIRE(ep.q8_out == nullptr, QB200_EUNSUPPORTED, "conv_dw: no quantized hand-off from depthwise layers");
    // 3x3 / pad 1 / stride 1 or 2 with one filter per channel: the streaming kernel (QB200_DW_STREAM=0: the band kernel)
    static const bool stream_on = [] {
        const char* e = getenv("QB200_DW_STREAM");
        return !(e && e[0] == '0');
    }();
    if (stream_on && g.R == 3 && g.S == 3 && g.pad == 1 && (g.stride == 1 || g.stride == 2) && g.K == g.C &&
        (g.stride == 1 || (g.W % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 8 == 0)) && (int64_t)g.H * g.W < (1ll << 30)) {
        const int need = g.stride == 1 ? g.Q + 2 : g.Q + 1;          // lanes one group needs to cover a whole row
        const int G = need <= 8 ? 8 : (need <= 16 ? 16 : 32);
        const int per_group = g.stride == 1 ? G - 2 : G - 1;
        const int gpp = (g.Q + per_group - 1) / per_group;
        const int64_t planes = (int64_t)g.N * g.K;
        const int64_t items = ((planes + 32 / G - 1) / (32 / G)) * gpp;
        QB_REQUIRE((items + 7) / 8 < (1ll << 31), QB200_EINVAL, "conv_dw: too many blocks");
        const dim3 sgrid((unsigned)((items + 7) / 8));
#define QB_DWS_LAUNCH(SIGNED, ST, GG)                                                                                         \
        QB_CUDA(launch_pdl(conv_dw3_stream_kernel<SIGNED, ST, GG>, sgrid, dim3(256), 0, st, x, wq, g, ep, out, gpp, items,   \
                           aq->scale, aq->zero, aq->qmin, aq->qmax))
#define QB_DWS_G(SIGNED, ST)                                                                                                  \
        do { if (G == 8) QB_DWS_LAUNCH(SIGNED, ST, 8); else if (G == 16) QB_DWS_LAUNCH(SIGNED, ST, 16); else QB_DWS_LAUNCH(SIGNED, ST, 32); } while (0)
        if (g.w_sign && g.stride == 1) QB_DWS_G(true, 1);
        else if (g.w_sign) QB_DWS_G(true, 2);
        else if (g.stride == 1) QB_DWS_G(false, 1);
        else QB_DWS_G(false, 2);
#undef QB_DWS_G
#undef QB_DWS_LAUNCH
        QB_LAUNCH_CHECK();
        return 0;
    }
    // rows per band: up to 16; channels per block: as many as give ~8K outputs (and fit 40 KB of patch)
    const int TH = std::min(16, g.P);
    const int bands = (g.P + TH - 1) / TH;
    const int in_h = (TH - 1) * g.stride + g.R, in_w = (g.Q - 1) * g.stride + g.S;
    const int in_w_alloc = (in_w + 7) & ~3;   // 4-byte rows with one spare word: the stencil reads aligned word pairs
    QB_REQUIRE((size_t)in_h * in_w_alloc <= 40 * 1024, QB200_EUNSUPPORTED, "conv_dw: plane too wide");
    int CH = std::max(1, 8192 / (TH * g.Q));
    CH = std::min(CH, std::min(64, g.K));
    auto smem_of = [&](int ch) {
        return (((size_t)ch * in_h * in_w_alloc + 15) & ~(size_t)15) + (size_t)ch * (g.R * g.S + 16) * sizeof(int);
    };
    while (CH > 1 && smem_of(CH) > 48 * 1024) --CH;   // patch + filter tables within the default dynamic shared memory
    const size_t smem = smem_of(CH);
    QB_REQUIRE(smem <= 48 * 1024, QB200_EUNSUPPORTED, "conv_dw: plane too wide");
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
