// Activation quantizer: fp32 NCHW -> u8 NHWC(Cp), the first half of the fused hot path.
//
// Restates Quantizer.simulate's quantize step (reference modelzoo/modules/quantizer.py:31 Round.forward,
// :215 `self.round(x, scale, zero).clamp(self.qmin, self.qmax)`) with per-tensor scale/zero:
//     q = clamp(rint(x / scale - zero), qmin, qmax)          all in fp32, IEEE divide, round-half-even
// and changes the layout to channel-last so that the implicit-GEMM conv can fetch K-contiguous im2col
// rows with TMA.  Channels C..Cp-1 are written as 0 (they meet zero weights in the MMA).
//
// Bit-exactness without an IEEE division per element and without data-dependent branches: quant_math.cuh computes
// RN(x/s) as x*RN(1/s) followed by two exact-residual FMA corrections (Markstein), then RN(. - z) and a magic-number
// round-to-nearest-even; tests/test_conv_gpu.py checks the integers against the oracle bit for bit, including inputs
// placed on and one ulp around every rounding boundary.
//
// Machine mapping (vector kernel, H*W % 4 == 0): a block is 128 consecutive pixels; a thread owns 4 consecutive
// pixels x 16 channels: 16 coalesced 16-byte loads (a warp reads 512 contiguous bytes of one channel plane),
// quantizes in registers, writes four 16-byte channel vectors into an XOR-swizzled shared tile, and the block
// copies the tile out as whole contiguous NHWC rows.  HBM-bound: bytes per element = 4 (read) + Cp/C (write).
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"
#include "conv_common.cuh"
#include "quant_math.cuh"

namespace qb200 {
namespace {

#ifndef QB200_VEC4_MIN_BLOCKS
#define QB200_VEC4_MIN_BLOCKS 6   // 6 x 128 threads per SM (<= 85 registers): 105 registers left 4 blocks per SM, and the small
#endif                            // layers (784 blocks for 256 ch @14x14) ran as 1.3 latency-bound waves (ncu, round 2)
constexpr int kPix = 128;  // pixels per block
constexpr int kThreads = 128;
constexpr int kCw = 128;   // channel bytes per shared-memory pass

__device__ __forceinline__ void copy_out(const uint4* tile, uint8_t* __restrict__ q, int64_t g0, int n_rows, int chunks, int Cp,
                                         int c_base) {
    // rows of `chunks` 16-byte vectors at stride Cp; consecutive threads take consecutive vectors (full lines)
    const int total = n_rows * chunks;
    for (int i = threadIdx.x; i < total; i += kThreads) {
        const int row = i / chunks, col = i - row * chunks;
        const uint4 v = tile[row * (kCw / 16) + (col ^ (row & 7))];
        *reinterpret_cast<uint4*>(q + (g0 + row) * (int64_t)Cp + c_base + col * 16) = v;
    }
}

// Zero-padded output for the halo path of the conv kernel: pixel (n,h,w) goes to row ((n*Hp + h + pad)*Wp + w + pad) of a
// [N*Hp*Wp][Cp] matrix (Hp = H + 2*pad, Wp = W + 2*pad) and the pad pixels are written as zeros by the blocks that own
// the neighbouring image-border pixels, so that the buffer never needs a separate memset.
struct PadSpec {
    int pad, H, W;
};

__device__ __forceinline__ void copy_out_padded(const uint4* tile, uint8_t* __restrict__ q, int64_t g0, int n_rows, int chunks,
                                                int Cp, int c_base, const PadSpec ps, int* rowoff) {
    const int Hp = ps.H + 2 * ps.pad, Wp = ps.W + 2 * ps.pad, HW = ps.H * ps.W;
    const int t = threadIdx.x;
    int n = 0, h = 0, w = 0;
    if (t < n_rows) {
        const int64_t g = g0 + t;
        n = (int)(g / HW);
        const int rem = (int)(g - (int64_t)n * HW);
        h = rem / ps.W;
        w = rem - h * ps.W;
        rowoff[t] = (n * Hp + h + ps.pad) * Wp + w + ps.pad;
    }
    __syncthreads();
    const int total = n_rows * chunks;
    for (int i = t; i < total; i += kThreads) {
        const int row = i / chunks, col = i - row * chunks;
        const uint4 v = tile[row * (kCw / 16) + (col ^ (row & 7))];
        *reinterpret_cast<uint4*>(q + (int64_t)rowoff[row] * Cp + c_base + col * 16) = v;
    }
    if (t < n_rows) {
        // pad pixels adjacent to this pixel: left / right of its row; above / below its column (corners with the
        // first / last pixel of the first / last row)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        auto zero_px = [&](int hp, int wp) {
            uint8_t* d = q + (int64_t)((n * Hp + hp) * Wp + wp) * Cp + c_base;
            for (int c = 0; c < chunks; ++c) reinterpret_cast<uint4*>(d)[c] = z;
        };
        const int wlo = (w == 0) ? 0 : w + ps.pad, whi = (w == ps.W - 1) ? Wp : w + ps.pad + 1;  // column span incl. corners
        if (w == 0)
            for (int i = 0; i < ps.pad; ++i) zero_px(h + ps.pad, i);
        if (w == ps.W - 1)
            for (int i = 0; i < ps.pad; ++i) zero_px(h + ps.pad, ps.W + ps.pad + i);
        if (h == 0)
            for (int r = 0; r < ps.pad; ++r)
                for (int c = wlo; c < whi; ++c) zero_px(r, c);
        if (h == ps.H - 1)
            for (int r = 0; r < ps.pad; ++r)
                for (int c = wlo; c < whi; ++c) zero_px(ps.H + ps.pad + r, c);
    }
}

// ---------------------------------------------------------------------------------------------
// vector kernel: H*W % 4 == 0, x 16-byte aligned.  thread = (pixel quad, 16-channel group)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, QB200_VEC4_MIN_BLOCKS)
act_quantize_nhwc_vec4_kernel(const float* __restrict__ x, uint8_t* __restrict__ q, int64_t total_pix, int C, int Cp, int HW,
                              const PadSpec ps, const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                              const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ uint4 tile[kPix * (kCw / 16)];
    __shared__ int rowoff[kPix];
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int pq = threadIdx.x & 31;   // pixel quad inside the block
    const int cg = threadIdx.x >> 5;   // 16-channel group inside a 64-channel sub-pass (warp-uniform)
    const int64_t g0 = (int64_t)blockIdx.x * kPix;
    const int64_t g = g0 + pq * 4;
    const int c_base = blockIdx.y * kCw;
    const int cw = min(kCw, Cp - c_base);  // multiple of 32
    const int chunks = cw >> 4;

    const bool active = g < total_pix;  // total_pix % 4 == 0: a quad is all-in or all-out, and never straddles images
    const int64_t n = active ? g / HW : 0;
    const int pix = (int)(g - n * HW);
    for (int sub = 0; sub < chunks; sub += 4) {     // 64 channels per sub-pass
        const int j = sub + cg;                      // this warp's 16-channel chunk
        if (active && j < chunks) {
            const int c0 = c_base + j * 16;
            const float* xp = x + (n * C + c0) * (int64_t)HW + pix;
            float4 v[16];
            if (c0 + 16 <= C) {   // warp-uniform: all 16 channel planes exist; walk the plane stride with pointer adds
                const float* pi = xp;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] = ldg_stream4(pi);
                    pi += HW;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] = (c0 + i < C) ? ldg_stream4(xp + (int64_t)i * HW) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t w[4][4];  // [pixel][word]
            quant_tile<4>(v, w, p);
            // padded channels must be exactly 0 even when qmin > 0
            if (c0 + 16 > C) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t keep = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (c0 + k * 4 + b < C) keep |= 0xFFu << (8 * b);
#pragma unroll
                    for (int px = 0; px < 4; ++px) w[px][k] &= keep;
                }
            }
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const int row = pq * 4 + px;
                tile[row * (kCw / 16) + (j ^ (row & 7))] = make_uint4(w[px][0], w[px][1], w[px][2], w[px][3]);
            }
        }
    }
    __syncthreads();
    const int n_rows = (int)min((int64_t)kPix, total_pix - g0);
    if (ps.pad > 0) copy_out_padded(tile, q, g0, n_rows, chunks, Cp, c_base, ps, rowoff);
    else copy_out(tile, q, g0, n_rows, chunks, Cp, c_base);
}

// ---------------------------------------------------------------------------------------------
// stride-2 sub-sampling kernel (the input of a 1x1 / stride-2 / pad-0 layer: ResNet's down-sampling shortcuts).
// Same machine mapping as the vector kernel, over the SAMPLED rows only: a thread owns 4 consecutive input pixels of a
// sampled row (one 16-byte load per channel: a warp reads whole 512-byte runs) x 16 channels, and keeps pixels 0 and 2.
// The scalar kernel reads every other 4-byte word with 16 loads in flight per thread: 2.3-3.5 TB/s of the bytes it
// touches; this one reads the same sectors with 16-byte loads.  Needs W % 4 == 0 and a 16-byte aligned input.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
act_quantize_nhwc_sub2_kernel(const float* __restrict__ x, uint8_t* __restrict__ q, int64_t total_quads, int64_t total_out, int C,
                              int Cp, int H, int W, int P, const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                              const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ uint4 tile[(kPix / 2) * (kCw / 16)];
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int pq = threadIdx.x & 31;   // quad inside the block
    const int cg = threadIdx.x >> 5;   // 16-channel group inside a 64-channel sub-pass (warp-uniform)
    const int W4 = W >> 2;
    const int64_t gq = (int64_t)blockIdx.x * 32 + pq;
    const int c_base = blockIdx.y * kCw;
    const int cw = min(kCw, Cp - c_base);
    const int chunks = cw >> 4;
    const bool active = gq < total_quads;
    int64_t n = 0;
    int prow = 0, w4 = 0;
    if (active) {
        const int64_t per_img = (int64_t)P * W4;
        n = gq / per_img;
        const int rem = (int)(gq - n * per_img);
        prow = rem / W4;
        w4 = rem - prow * W4;
    }
    const int64_t HW = (int64_t)H * W;
    for (int sub = 0; sub < chunks; sub += 4) {
        const int j = sub + cg;
        if (active && j < chunks) {
            const int c0 = c_base + j * 16;
            const float* xp = x + (n * C + c0) * HW + (int64_t)prow * 2 * W + w4 * 4;
            float4 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (c0 + i < C) ? ldg_stream4(xp + (int64_t)i * HW) : make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t w0[4], w1[4];   // output pixels 2*w4 and 2*w4 + 1 (input columns 4*w4 and 4*w4 + 2)
            if (p.byte_clamp) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    int a[4], b[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) quant_int2(v[4 * k + c].x, v[4 * k + c].z, p, a[c], b[c]);
                    w0[k] = pack_clamp4<false>(a[0], a[1], a[2], a[3], p);
                    w1[k] = pack_clamp4<false>(b[0], b[1], b[2], b[3], p);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    w0[k] = quant_word_exact(v[4 * k].x, v[4 * k + 1].x, v[4 * k + 2].x, v[4 * k + 3].x, p.s, p.z, p.lo, p.hi);
                    w1[k] = quant_word_exact(v[4 * k].z, v[4 * k + 1].z, v[4 * k + 2].z, v[4 * k + 3].z, p.s, p.z, p.lo, p.hi);
                }
            }
            if (c0 + 16 > C) {   // padded channels must be exactly 0 even when qmin > 0
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t keep = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (c0 + k * 4 + b < C) keep |= 0xFFu << (8 * b);
                    w0[k] &= keep;
                    w1[k] &= keep;
                }
            }
            const int r0 = pq * 2, r1 = pq * 2 + 1;
            tile[r0 * (kCw / 16) + (j ^ (r0 & 7))] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
            tile[r1 * (kCw / 16) + (j ^ (r1 & 7))] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
        }
    }
    __syncthreads();
    const int64_t g0 = (int64_t)blockIdx.x * (kPix / 2);
    const int n_rows = (int)min((int64_t)(kPix / 2), total_out - g0);
    copy_out(tile, q, g0, n_rows, chunks, Cp, c_base);
}

// ---------------------------------------------------------------------------------------------
// generic kernel: any H*W / alignment.  thread = one pixel, scalar (still warp-coalesced) loads
// ---------------------------------------------------------------------------------------------
// sub > 1: only the pixels (p*sub, q*sub) of each image are quantized, into a compact [N, P, Q, Cp] buffer — what a
// 1x1 convolution with stride `sub` and no padding reads (the other 1 - 1/sub^2 of the input is never touched).
__global__ void __launch_bounds__(kThreads)
act_quantize_nhwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ q, int64_t total_pix, int C,
                         int Cp, int HW, int sub, int W_in, int Q_out, int PQ_out, const PadSpec ps,
                         const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                         const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ uint4 tile[kPix * (kCw / 16)];
    __shared__ int rowoff[kPix];
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int t = threadIdx.x;
    const int64_t g0 = (int64_t)blockIdx.x * kPix;
    const int64_t g = g0 + t;
    const int c_base = blockIdx.y * kCw;
    const int cw = min(kCw, Cp - c_base);  // multiple of 32
    const int chunks = cw >> 4;

    if (g < total_pix) {
        int64_t n;
        int pix;
        if (sub > 1) {
            n = g / PQ_out;
            const int po = (int)(g - n * PQ_out);
            const int p = po / Q_out, qq = po - p * Q_out;
            pix = p * sub * W_in + qq * sub;
        } else {
            n = g / HW;
            pix = (int)(g - n * HW);
        }
        const float* xp = x + (n * C + c_base) * (int64_t)HW + pix;
        for (int j = 0; j < chunks; ++j) {
            float v[16];
            const int c0 = c_base + j * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (c0 + i < C) ? __ldg(xp + (int64_t)(j * 16 + i) * HW) : 0.f;
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t keep = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (c0 + k * 4 + b < C) keep |= 0xFFu << (8 * b);
                w[k] = quant_word(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3], p) & keep;
            }
            tile[t * (kCw / 16) + (j ^ (t & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    __syncthreads();
    const int n_rows = (int)min((int64_t)kPix, total_pix - g0);
    if (ps.pad > 0) copy_out_padded(tile, q, g0, n_rows, chunks, Cp, c_base, ps, rowoff);
    else copy_out(tile, q, g0, n_rows, chunks, Cp, c_base);
}

// ---------------------------------------------------------------------------------------------
// band kernel: planes that the vector kernel cannot take (H*W % 4 != 0: 7x7, 9x9, ...) and sub-sampled inputs (1x1 / stride s).
// The scalar kernel above keeps 16 four-byte loads in flight per thread and serialises one HBM round trip per 16-channel
// chunk: 1.7 TB/s on 7x7 planes, 2.3-2.6 TB/s sub-sampled (profiles/README.md).  Here a block = (image, 32-channel slab,
// band of the input rows that are actually read): phase 1 streams the band of every channel into shared memory with
// cp.async (16 / 8 / 4 bytes per copy, whatever the row alignment allows) — the whole block's input is in flight at once —
// phase 2 turns it: a thread owns one output pixel, reads its 32 channels from shared memory (lanes = consecutive pixels:
// conflict-free), quantizes them and writes one full 32-byte sector of the NHWC row.
// ---------------------------------------------------------------------------------------------
constexpr int kBandCh = 32;            // channels per block
constexpr int kBandFloats = 384;       // floats per channel and band in shared memory (48 KB per block)

__device__ __forceinline__ void cp_async_n(void* dst_smem, const void* src, int bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
    if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

// rows_band input rows per band (each `sub`-th row of the image), `seg` floats per row (the whole row), vec = bytes per copy.
// dense (sub == 1, band = whole plane, vec chosen for the CONTIGUOUS slab of channels): one long segment per block.
__global__ void __launch_bounds__(256)
act_quantize_band_kernel(const float* __restrict__ x, uint8_t* __restrict__ q, int C, int Cp, int H, int W, int sub, int P_out,
                         int Q_out, int rows_band, int n_bands, int vec, const float* __restrict__ p_scale,
                         const float* __restrict__ p_zero, const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) float band[];   // [kBandCh][plane], plane = rows_band * W floats
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int n_slabs = Cp / kBandCh;
    int b = blockIdx.x;
    const int bi = b % n_bands; b /= n_bands;
    const int slab = b % n_slabs;
    const int n = b / n_slabs;
    const int c0 = slab * kBandCh;
    const int nch = min(kBandCh, C - c0);            // <= 0: a slab of padded channels only (all zeros)
    const int r0 = bi * rows_band;                   // first output row of the band
    const int rows = min(rows_band, P_out - r0);
    const int plane = rows_band * W;
    const int HW = H * W;
    // ---- phase 1: global -> shared ----
    if (nch > 0) {
        const float* src0 = x + ((int64_t)n * C + c0) * HW;
        const int fpc = vec >> 2;                    // floats per copy
        if (sub == 1 && n_bands == 1) {
            // the slab's planes are one contiguous run of nch * HW floats (plane == HW)
            const int total = nch * HW;
            const int n_copies = total / fpc;
            for (int i = threadIdx.x; i < n_copies; i += blockDim.x) cp_async_n(band + i * fpc, src0 + i * fpc, vec);
            for (int i = n_copies * fpc + threadIdx.x; i < total; i += blockDim.x) cp_async_n(band + i, src0 + i, 4);
        } else {
            const int per_row = W / fpc;             // W % fpc == 0 by construction of vec
            const int n_copies = nch * rows * per_row;
            for (int i = threadIdx.x; i < n_copies; i += blockDim.x) {
                const int seg = i / per_row, o = (i - seg * per_row) * fpc;
                const int c = seg / rows, r = seg - c * rows;
                cp_async_n(band + c * plane + r * W + o, src0 + (int64_t)c * HW + (int64_t)(r0 + r) * sub * W + o, vec);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // ---- phase 2: one thread = one output pixel x 32 channels ----
    const int npix = rows * Q_out;
    for (int i = threadIdx.x; i < npix; i += blockDim.x) {
        const int r = i / Q_out, qq = i - r * Q_out;
        const float* s = band + r * W + qq * sub;
        uint32_t w[8];
        if (nch == kBandCh) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = s[j * plane];
            quant_row<8>(v, w, p);
        } else {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = j < nch ? s[j * plane] : 0.f;
            quant_row<8>(v, w, p);
#pragma unroll
            for (int k = 0; k < 8; ++k) {                  // padded channels must be exactly 0 even when qmin > 0
                uint32_t keep = 0;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb)
                    if (k * 4 + bb < nch) keep |= 0xFFu << (8 * bb);
                w[k] &= keep;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(q + (((int64_t)n * P_out + r0 + r) * Q_out + qq) * Cp + c0);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// ---------------------------------------------------------------------------------------------
// few-channel layers (RGB stem): quantize straight into im2col rows.
// One block = 8 output rows of one image.  The input rows they touch are quantized ONCE into a shared patch of packed
// words (one word = the <=4 channels of a pixel); pixels outside the image are 0, so that padded taps add nothing to
// sum(qa*qw) — the reference skips them (quantconv2d_float_input.cu:92) and their zero-point term is excluded by the
// border tables.  Then every output pixel's row of Kcol bytes (word r*S+s = patch[r][q*stride + s], remaining words 0)
// is written with fully coalesced stores.
// ---------------------------------------------------------------------------------------------
constexpr int kIm2colRows = 8;  // output rows per block: (8-1)*stride + R input rows are quantized once and shared

__global__ void __launch_bounds__(256)
act_quantize_im2col_kernel(const float* __restrict__ x, uint32_t* __restrict__ a_col, ConvGeom g, int Kwords,
                           const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                           const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ uint32_t patch[];  // [rows_in][Wp], Wp = W + 2*pad
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int Wp = g.W + 2 * g.pad;
    const int pblocks = (g.P + kIm2colRows - 1) / kIm2colRows;
    const int n = blockIdx.x / pblocks, p0 = (blockIdx.x - n * pblocks) * kIm2colRows;
    const int n_out = min(kIm2colRows, g.P - p0);
    const int rows_in = (n_out - 1) * g.stride + g.R;
    const int h0 = p0 * g.stride - g.pad;
    const int HW = g.H * g.W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t keep = g.C >= 4 ? 0xFFFFFFFFu : (1u << (8 * g.C)) - 1u;
    for (int r = warp; r < rows_in; r += 8) {
        const int ih = h0 + r;
        const bool row_in = ih >= 0 && ih < g.H;
        const float* xr = x + (int64_t)n * g.C * HW + (int64_t)ih * g.W;
        for (int j = lane; j < Wp; j += 32) {
            const int iw = j - g.pad;
            uint32_t word = 0;
            if (row_in && iw >= 0 && iw < g.W) {
                float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < g.C) a[c] = __ldg(xr + (int64_t)c * HW + iw);
                word = quant_word(a[0], a[1], a[2], a[3], p) & keep;
            }
            patch[r * Wp + j] = word;
        }
    }
    __syncthreads();
    const int taps = g.R * g.S;
    for (int t = 0; t < n_out; ++t) {
        uint32_t* orow = a_col + ((int64_t)n * g.P + p0 + t) * g.Q * Kwords;
        const uint32_t* prow = patch + t * g.stride * Wp;
        if (blockDim.x % Kwords == 0) {
            // every thread owns one word position of the row for all the pixels it visits: no per-word division
            const int wd = threadIdx.x % Kwords, q0 = threadIdx.x / Kwords, qstep = blockDim.x / Kwords;
            const int r = wd / g.S;
            const int off = (wd < taps) ? r * Wp + (wd - r * g.S) : -1;
            for (int q = q0; q < g.Q; q += qstep) orow[q * Kwords + wd] = off >= 0 ? prow[off + q * g.stride] : 0u;
        } else {
            for (int i = threadIdx.x; i < g.Q * Kwords; i += blockDim.x) {
                const int q = i / Kwords, wd = i - q * Kwords;
                uint32_t v = 0;
                if (wd < taps) {
                    const int r = wd / g.S, s = wd - r * g.S;
                    v = prow[r * Wp + q * g.stride + s];
                }
                orow[i] = v;
            }
        }
    }
}

// Grouped rows (im2col_grouped: the 7x7 stem).  Same block shape; the input rows are quantized once into per-channel
// byte planes plane[c][row][Wq] (column j = input column j - pad; zeros outside the image), and the 8 bytes of group
// (c, r) of output pixel (p, q) are the window plane[c][p*stride + r][q*stride .. q*stride + 7]: three aligned words and
// two funnel shifts.  Bytes C*R*8 .. Kcol-1 of a row are never written: their weights are zero and any byte times zero
// is zero in integer arithmetic.
__global__ void __launch_bounds__(256)
act_quantize_im2col8_kernel(const float* __restrict__ x, uint8_t* __restrict__ a_col, ConvGeom g, int Kcol, int Wq,
                            const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                            const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ uint32_t plane32[];  // [C][rows_in][Wq / 4] words
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int pblocks = (g.P + kIm2colRows - 1) / kIm2colRows;
    const int n = blockIdx.x / pblocks, p0 = (blockIdx.x - n * pblocks) * kIm2colRows;
    const int n_out = min(kIm2colRows, g.P - p0);
    const int rows_max = (kIm2colRows - 1) * g.stride + g.R;
    const int rows_in = (n_out - 1) * g.stride + g.R;
    const int h0 = p0 * g.stride - g.pad;
    const int HW = g.H * g.W, Ww = Wq >> 2;
    // ---- quantize: one thread = 4 consecutive columns of one (channel, row) ----
    for (int i = threadIdx.x; i < g.C * rows_in * Ww; i += blockDim.x) {
        const int cr = i / Ww, jw = i - cr * Ww;
        const int c = cr / rows_in, r = cr - c * rows_in;
        const int ih = h0 + r, iw0 = jw * 4 - g.pad;
        uint32_t word = 0;
        if (ih >= 0 && ih < g.H && iw0 + 3 >= 0 && iw0 < g.W) {
            const float* xr = x + ((int64_t)n * g.C + c) * HW + (int64_t)ih * g.W;
            float a[4];
            uint32_t keep = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const bool in = iw0 + b >= 0 && iw0 + b < g.W;
                a[b] = in ? __ldg(xr + iw0 + b) : 0.f;
                if (in) keep |= 0xFFu << (8 * b);
            }
            word = quant_word(a[0], a[1], a[2], a[3], p) & keep;
        }
        plane32[(c * rows_max + r) * Ww + jw] = word;
    }
    __syncthreads();
    // ---- write: thread = one group index for a strided set of pixels ----
    const int G = g.C * g.R, Gall = Kcol >> 3;           // real groups / 8-byte slots per row (the rest are written as 0:
    const int per_iter = blockDim.x / Gall;              // whole 32-byte sectors, no read-modify-write in DRAM)
    const int gi = threadIdx.x % Gall, qlane = threadIdx.x / Gall;
    if (qlane < per_iter) {
        const int c = gi / g.R, r = gi - c * g.R;
        const bool real = gi < G;
        for (int t = 0; t < n_out; ++t) {
            uint8_t* orow = a_col + ((int64_t)n * g.P + p0 + t) * g.Q * Kcol + gi * 8;
            const uint32_t* prow = plane32 + (real ? (c * rows_max + t * g.stride + r) * Ww : 0);
#pragma unroll 4
            for (int q = qlane; q < g.Q; q += per_iter) {
                const int col = q * g.stride;
                const uint32_t* w = prow + (col >> 2);
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
                const int sh = (col & 3) * 8;
                uint2 v;
                v.x = real ? __funnelshift_r(w0, w1, sh) : 0u;
                v.y = real ? __funnelshift_r(w1, w2, sh) : 0u;
                *reinterpret_cast<uint2*>(orow + (int64_t)q * Kcol) = v;
            }
        }
    }
}

// H*W == 1 (linear layers as 1x1 convolutions): NCHW is already NHWC; one thread = one word of four channels
__global__ void __launch_bounds__(256)
act_quantize_rows_kernel(const float* __restrict__ x, uint32_t* __restrict__ q, int64_t words, int C, int Cp,
                         const float* __restrict__ p_scale, const float* __restrict__ p_zero,
                         const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int wpr = Cp >> 2;   // words per row
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i / wpr;
        const int c = (int)(i - n * wpr) * 4;
        const float* xp = x + n * C + c;
        float a[4];
        uint32_t keep = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const bool in = c + b < C;
            a[b] = in ? __ldg(xp + b) : 0.f;
            if (in) keep |= 0xFFu << (8 * b);
        }
        q[i] = quant_word(a[0], a[1], a[2], a[3], p) & keep;
    }
}

}  // namespace

int launch_act_quantize_im2col(const float* x, const ConvGeom& g, int Kcol, const qb200_act_quant* aq, uint8_t* a_col,
                               cudaStream_t st) {
    QB_REQUIRE(aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL,
               "act_quantize: activation quantizer parameters missing");
    QB_REQUIRE(g.C <= 4 && g.groups == 1, QB200_EINVAL, "act_quantize_im2col: not a few-channel layer");
    if (im2col_grouped(g.C, g.R, g.S)) {
        QB_REQUIRE(Kcol == im2col8_row_bytes(g.C, g.R) && g.C * g.R <= 256, QB200_EINVAL, "act_quantize_im2col: bad row size");
        // plane width: the last window starts at (Q-1)*stride and spans 8 bytes + 4 of slack for the third word
        const int Wq = round_up_int(std::max(g.W + 2 * g.pad, (g.Q - 1) * g.stride + 12), 4);
        const size_t smem8 = (size_t)g.C * ((kIm2colRows - 1) * g.stride + g.R) * Wq;
        QB_REQUIRE(smem8 <= 48 * 1024, QB200_EUNSUPPORTED, "act_quantize_im2col: input row too wide");
        const int pb = (g.P + kIm2colRows - 1) / kIm2colRows;
        QB_CUDA(launch_pdl(act_quantize_im2col8_kernel, dim3((unsigned)(g.N * pb)), dim3(256), smem8, st, x, a_col, g, Kcol, Wq,
                           aq->scale, aq->zero, aq->qmin, aq->qmax));
        QB_LAUNCH_CHECK();
        return 0;
    }
    QB_REQUIRE(Kcol >= g.R * g.S * 4, QB200_EINVAL, "act_quantize_im2col: bad row size");
    const size_t smem = (size_t)((kIm2colRows - 1) * g.stride + g.R) * (g.W + 2 * g.pad) * sizeof(uint32_t);
    QB_REQUIRE(smem <= 48 * 1024, QB200_EUNSUPPORTED, "act_quantize_im2col: input row too wide");
    const int pblocks = (g.P + kIm2colRows - 1) / kIm2colRows;
    QB_CUDA(launch_pdl(act_quantize_im2col_kernel, dim3((unsigned)(g.N * pblocks)), dim3(256), smem, st, x, reinterpret_cast<uint32_t*>(a_col), g, Kcol / 4,
                                                                         aq->scale, aq->zero, aq->qmin, aq->qmax));
    QB_LAUNCH_CHECK();
    return 0;
}

}  // namespace qb200


namespace qb200 {
namespace {
// Launch the band kernel for (optionally sub-sampled) fp32 NCHW -> compact u8 [N, P_out, Q_out, Cp].  Returns -100 when the
// shape does not fit its shared-memory band (the caller then uses the scalar kernel).
int launch_band(const float* x, uint8_t* q, int N, int C, int Cp, int H, int W, int sub, const qb200_act_quant* aq, cudaStream_t st) {
    const int P_out = (H + sub - 1) / sub, Q_out = (W + sub - 1) / sub;
    const int HW = H * W;
    int rows_band, n_bands, vec;
    const uintptr_t a = reinterpret_cast<uintptr_t>(x);
    if (sub == 1 && HW <= kBandFloats) {
        rows_band = H;
        n_bands = 1;
        // a slab starts (n * C + 32 * k) * HW floats from x: 16-byte copies need C * HW to be a multiple of 4 floats
        const int64_t img = (int64_t)C * HW;
        vec = (a % 16 == 0 && img % 4 == 0) ? 16 : ((a % 8 == 0 && img % 2 == 0) ? 8 : 4);
    } else {
        if (W > kBandFloats) return -100;
        rows_band = kBandFloats / W;
        if (rows_band > P_out) rows_band = P_out;
        n_bands = (P_out + rows_band - 1) / rows_band;
        vec = (a % 16 == 0 && W % 4 == 0) ? 16 : ((a % 8 == 0 && W % 2 == 0) ? 8 : 4);
    }
    const int plane = rows_band * W;
    const size_t smem = (size_t)kBandCh * plane * sizeof(float);
    if (smem > 48 * 1024) return -100;
    const int64_t blocks = (int64_t)N * (Cp / kBandCh) * n_bands;
    if (blocks >= (1ll << 31)) return -100;
    const int npix = rows_band * Q_out;
    const int threads = npix >= 192 ? 256 : (npix >= 96 ? 128 : 64);
    QB_CUDA(launch_pdl(act_quantize_band_kernel, dim3((unsigned)blocks), dim3(threads), smem, st, x, q, C, Cp, H, W, sub, P_out, Q_out,
                       rows_band, n_bands, vec, aq->scale, aq->zero, aq->qmin, aq->qmax));
    QB_LAUNCH_CHECK();
    return 0;
}
// QB200_BAND_QUANT: 0 = never, 1 = planes the vector kernel cannot take + sub-sampled inputs, 2 = also every dense plane of
// at most kBandFloats pixels (14x14, 7x7 ...) — A/B switch
int band_mode() {
    static const int mode = [] {
        const char* e = getenv("QB200_BAND_QUANT");
        return e ? atoi(e) : 1;
    }();
    return mode;
}
bool band_enabled() { return band_mode() != 0; }
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
    const int HW = H * W;
    const int64_t total = (int64_t)N * HW;
    dim3 grid((unsigned)ceil_div64(total, kPix), (unsigned)((Cp + kCw - 1) / kCw));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PadSpec ps{0, H, W};
    if (HW == 1) {
        const int64_t words = (int64_t)N * (Cp >> 2);
        const int blocks = (int)std::min<int64_t>(ceil_div64(words, 256), (int64_t)kNumSMs * 8);
        QB_CUDA(launch_pdl(act_quantize_rows_kernel, dim3(blocks), dim3(256), 0, st, x, reinterpret_cast<uint32_t*>(q_nhwc), words, C, Cp,
                           aq->scale, aq->zero, aq->qmin, aq->qmax));
    } else if (band_mode() == 2 && HW <= kBandFloats) {
        const int rc = launch_band(x, q_nhwc, N, C, Cp, H, W, 1, aq, st);
        if (rc != -100) return rc;
        QB_REQUIRE(false, QB200_EUNSUPPORTED, "act_quantize: band kernel refused a small plane");
    } else if (HW % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0)
        QB_CUDA(launch_pdl(act_quantize_nhwc_vec4_kernel, dim3(grid), dim3(kThreads), 0, st, x, q_nhwc, total, C, Cp, HW, ps, aq->scale, aq->zero, aq->qmin,
                                                                  aq->qmax));
    else {
        if (band_enabled()) {
            const int rc = launch_band(x, q_nhwc, N, C, Cp, H, W, 1, aq, st);
            if (rc != -100) return rc;
        }
        QB_CUDA(launch_pdl(act_quantize_nhwc_kernel, dim3(grid), dim3(kThreads), 0, st, x, q_nhwc, total, C, Cp, HW, 1, W, W, HW, ps, aq->scale, aq->zero,
                                                             aq->qmin, aq->qmax));
    }
    QB_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

namespace qb200 {

int launch_act_quantize_subsampled(const float* x, const ConvGeom& g, const qb200_act_quant* aq, uint8_t* q, cudaStream_t st) {
    QB_REQUIRE(aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL,
               "act_quantize: activation quantizer parameters missing");
    QB_REQUIRE(g.R == 1 && g.S == 1 && g.pad == 0 && g.stride > 1, QB200_EINVAL, "act_quantize_subsampled: not a strided 1x1 layer");
    const int64_t total = (int64_t)g.N * g.P * g.Q;
    static const bool sub2_ok = [] {
        const char* e = getenv("QB200_SUB2_QUANT");
        return !(e && e[0] == '0');
    }();
    if (sub2_ok && g.stride == 2 && g.W % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && g.Q * 2 == g.W) {
        const int64_t quads = (int64_t)g.N * g.P * (g.W / 4);
        dim3 grid2((unsigned)ceil_div64(quads, 32), (unsigned)((g.Cp + kCw - 1) / kCw));
        QB_CUDA(launch_pdl(act_quantize_nhwc_sub2_kernel, grid2, dim3(kThreads), 0, st, x, q, quads, total, g.C, g.Cp, g.H, g.W, g.P,
                           aq->scale, aq->zero, aq->qmin, aq->qmax));
        QB_LAUNCH_CHECK();
        return 0;
    }
    // (pad == 0 and R == 1: P = (H - 1) / stride + 1 = ceil(H / stride), what the band kernel derives)
    // measured (profiles/README.md, round 2): the band kernel LOSES on sub-sampled inputs (256ch @56 s2: 132 vs 116 us;
    // 1024ch @14 s2: 58 vs 49 us — 8-byte copies of half-used rows), so it is opt-in here
    if (band_mode() == 3) {
        const int rc = launch_band(x, q, g.N, g.C, g.Cp, g.H, g.W, g.stride, aq, st);
        if (rc != -100) return rc;
    }
    dim3 grid((unsigned)ceil_div64(total, kPix), (unsigned)((g.Cp + kCw - 1) / kCw));
    QB_CUDA(launch_pdl(act_quantize_nhwc_kernel, dim3(grid), dim3(kThreads), 0, st, x, q, total, g.C, g.Cp, g.H * g.W, g.stride, g.W, g.Q, g.P * g.Q,
                                                         PadSpec{0, g.H, g.W}, aq->scale, aq->zero, aq->qmin, aq->qmax));
    QB_LAUNCH_CHECK();
    return 0;
}

// zero-padded NHWC for the halo path: [N][H + 2*pad][W + 2*pad][Cp]
int launch_act_quantize_padded(const float* x, const ConvGeom& g, const qb200_act_quant* aq, uint8_t* q, cudaStream_t st) {
    QB_REQUIRE(aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL,
               "act_quantize: activation quantizer parameters missing");
    const int HW = g.H * g.W;
    const int64_t total = (int64_t)g.N * HW;
    dim3 grid((unsigned)ceil_div64(total, kPix), (unsigned)((g.Cp + kCw - 1) / kCw));
    const PadSpec ps{g.pad, g.H, g.W};
    if (HW % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0)
        QB_CUDA(launch_pdl(act_quantize_nhwc_vec4_kernel, dim3(grid), dim3(kThreads), 0, st, x, q, total, g.C, g.Cp, HW, ps, aq->scale, aq->zero, aq->qmin,
                                                                  aq->qmax));
    else
        QB_CUDA(launch_pdl(act_quantize_nhwc_kernel, dim3(grid), dim3(kThreads), 0, st, x, q, total, g.C, g.Cp, HW, 1, g.W, g.W, HW, ps, aq->scale, aq->zero,
                                                             aq->qmin, aq->qmax));
    QB_LAUNCH_CHECK();
    return 0;
}


namespace {
// One 16-byte chunk per thread over the pad pixels only: for each image the top and bottom `pad` rows (full width) and the
// left / right `pad` columns of the rows between.
__global__ void __launch_bounds__(256)
zero_pad_borders_kernel(uint4* __restrict__ q, int N, int H, int W, int pad, int chunks) {
    pdl_launch_dependents();
    pdl_wait();
    const int Hp = H + 2 * pad, Wp = W + 2 * pad;
    const int border = 2 * pad * Wp + 2 * pad * H;  // pad pixels per image
    const int64_t total = (int64_t)N * border * chunks;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % chunks);
        const int64_t bp = i / chunks;
        const int n = (int)(bp / border);
        int b = (int)(bp - (int64_t)n * border);
        int h, w;
        if (b < pad * Wp) {                       // top rows
            h = b / Wp; w = b - h * Wp;
        } else if (b < 2 * pad * Wp) {            // bottom rows
            b -= pad * Wp;
            h = b / Wp; w = b - h * Wp; h += pad + H;
        } else {                                  // side columns of the H interior rows
            b -= 2 * pad * Wp;
            h = b / (2 * pad);
            const int j = b - h * 2 * pad;
            w = j < pad ? j : W + j;
            h += pad;
        }
        q[((int64_t)(n * Hp + h) * Wp + w) * chunks + c] = make_uint4(0u, 0u, 0u, 0u);
    }
}
}  // namespace

int launch_zero_pad_borders(uint8_t* q, int N, int H, int W, int pad, int Cp, cudaStream_t st) {
    QB_REQUIRE(q && Cp % 16 == 0 && reinterpret_cast<uintptr_t>(q) % 16 == 0, QB200_EINVAL, "zero_pad_borders: bad buffer");
    if (pad == 0 || N == 0) return 0;
    const int64_t total = (int64_t)N * (2 * pad * (W + 2 * pad) + 2 * pad * H) * (Cp / 16);
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), 148 * 8);
    QB_CUDA(launch_pdl(zero_pad_borders_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<uint4*>(q), N, H, W, pad, Cp / 16));
    QB_LAUNCH_CHECK();
    return 0;
}


namespace {
// Quantizer.simulate for a per-tensor quantizer as ONE elementwise pass (the reference runs five torch kernels:
// div, sub, round, clamp, add*mul — quantizer.py:194, :215-218):  out = (clamp(rint(x / s - z), qmin, qmax) + z) * s,
// every step rounded like the separate fp32 ops.
__device__ __forceinline__ float fake_quant1(float x, const QuantParams& p) {
    if (x != x) return x;   // torch's clamp propagates NaN
    float q;
    if (p.byte_clamp) {
        q = (float)min(max(quant_int(x, p), p.ilo), p.ihi);
    } else {
        q = rintf(__fsub_rn(__fdiv_rn(x, p.s), p.z));
        q = fminf(fmaxf(q, p.lo), p.hi);
    }
    return __fmul_rn(__fadd_rn(q, p.z), p.s);
}

__global__ void __launch_bounds__(256)
fake_quantize_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, const float* __restrict__ p_scale,
                     const float* __restrict__ p_zero, const float* __restrict__ p_qmin, const float* __restrict__ p_qmax) {
    pdl_launch_dependents();
    pdl_wait();
    const QuantParams p = load_params(p_scale, p_zero, p_qmin, p_qmax);
    const int64_t n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        for (int64_t i = t0; i < n4; i += stride) {
            const float4 v = ldg_stream4(x + 4 * i);
            float4 o;
            o.x = fake_quant1(v.x, p); o.y = fake_quant1(v.y, p); o.z = fake_quant1(v.z, p); o.w = fake_quant1(v.w, p);
            reinterpret_cast<float4*>(out)[i] = o;
        }
        for (int64_t i = 4 * n4 + t0; i < n; i += stride) out[i] = fake_quant1(__ldg(x + i), p);
    } else {
        for (int64_t i = t0; i < n; i += stride) out[i] = fake_quant1(__ldg(x + i), p);
    }
}
}  // namespace

}  // namespace qb200

extern "C" int qb200_fake_quantize_f32(const float* x, int64_t n, const qb200_act_quant* aq, float* out, void* stream) {
    using namespace qb200;
    QB_REQUIRE(n >= 0, QB200_EINVAL, "fake_quantize: negative size");
    if (n == 0) return 0;
    QB_REQUIRE(x && out && aq && aq->scale && aq->zero && aq->qmin && aq->qmax, QB200_EINVAL, "fake_quantize: null pointer");
    const int blocks = (int)std::min<int64_t>(ceil_div64(ceil_div64(n, 4), 256), (int64_t)kNumSMs * 16);
    QB_CUDA(launch_pdl(fake_quantize_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, n, aq->scale,
                       aq->zero, aq->qmin, aq->qmax));
    QB_LAUNCH_CHECK();
    return 0;
}
