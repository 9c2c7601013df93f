// Weight preparation: the reference's packed OIHW bit stream -> MMA-ready channel-last bytes + border tables.
//
// The checkpointed weight stays what QuantConv2d.pack() produced (reference modelzoo/modules/quantconv2d.py:186-191:
// tpack of round(w/s - z) in OIHW order, descriptor [n_bits, sign, K, Cg, R, S]).  This kernel derives, once per
// weight tensor (the caller caches it), the operand the conv kernels consume:
//   wq   [K][R][S][Cgp]   one byte per weight = the value the reference kernel reconstructs per MAC
//                         (quantconv2d_float_input.cu:94-103: field at bit e*n_bits, minus 2^(n-1) when signed)
//   wpre [K][R+1][S+1]    exclusive 2-D prefix sums of T[k][r][s] = sum_c wq[k][r][s][c]; the zero-point term
//                         z_a * sum_{in-bounds taps} qw of a border pixel is 4 lookups.
#include "common.cuh"
#include "conv_common.cuh"

namespace qb200 {

PreparedLayout prepared_layout(const qb200_conv_shape& s) {
    PreparedLayout L;
    const int groups = s.C / s.Cg;
    L.Cgp = (groups == 1) ? qb200_padded_channels(s.Cg) : round_up_int(s.Cg, 4);
    L.wq_bytes = align_up_sz((size_t)s.K * s.R * s.S * L.Cgp, 256);
    L.wpre_off = L.wq_bytes;
    L.total = L.wpre_off + align_up_sz((size_t)s.K * (s.R + 1) * (s.S + 1) * sizeof(int32_t), 256);
    L.Kcol = 0;
    L.wcol_off = L.total;
    if (uses_im2col_rows(s)) {
        L.Kcol = im2col_kcol(s.C, s.R, s.S);
        L.total = L.wcol_off + align_up_sz((size_t)s.K * L.Kcol, 256);
    }
    L.tapKC = 0;
    L.wtap_off = L.total;
    if (groups == 1 && s.R * s.S > 1 && s.C > 4) {
        L.tapKC = kblock_bytes(L.Cgp);
        L.total = L.wtap_off + align_up_sz((size_t)s.K * s.R * s.S * L.Cgp, 256);
    }
    return L;
}

int validate_shape(const qb200_conv_shape* s) {
    QB_REQUIRE(s != nullptr, QB200_EINVAL, "conv: null shape");
    QB_REQUIRE(s->N >= 0 && s->C > 0 && s->H > 0 && s->W > 0 && s->K > 0 && s->Cg > 0 && s->R > 0 && s->S > 0,
               QB200_EINVAL, "conv: non-positive dimension");
    QB_REQUIRE(s->stride > 0 && s->pad >= 0, QB200_EINVAL, "conv: bad stride/padding");
    QB_REQUIRE(s->w_bits > 0 && s->w_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    QB_REQUIRE(s->C % s->Cg == 0, QB200_EINVAL, "conv: input channels %d not divisible by weight channels %d", s->C, s->Cg);
    QB_REQUIRE(s->K % (s->C / s->Cg) == 0, QB200_EINVAL, "conv: out channels %d not divisible by groups %d", s->K, s->C / s->Cg);
    QB_REQUIRE(s->H + 2 * s->pad >= s->R && s->W + 2 * s->pad >= s->S, QB200_EINVAL, "conv: kernel larger than padded input");
    QB_REQUIRE(s->R <= 15 && s->S <= 15, QB200_EUNSUPPORTED, "conv: kernel sizes above 15 are not supported");
    return 0;
}

namespace {

__global__ void unpack_weights_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ wq, int K, int Cg,
                                      int R, int S, int Cgp, int nb, uint32_t offset, int64_t n_bytes) {
    const int64_t total = (int64_t)K * R * S * Cgp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % Cgp);
    int64_t t = i / Cgp;
    const int s = (int)(t % S);
    t /= S;
    const int r = (int)(t % R);
    const int k = (int)(t / R);
    uint8_t v = 0;
    if (c < Cg) {
        const int64_t e = (((int64_t)k * Cg + c) * R + r) * S + s;  // quantconv2d_float_input.cu:94 (OIHW)
        const int64_t bit = e * nb;
        const int64_t byte = bit >> 3;
        const int off = (int)(bit & 7);
        uint32_t w = packed[byte];
        if (off + nb > 8 && byte + 1 < n_bytes) w |= (uint32_t)packed[byte + 1] << 8;
        v = (uint8_t)(((w >> off) & ((1u << nb) - 1u)) - offset);  // :97-102
    }
    wq[i] = v;
}

// im2col weight rows: byte (r*S + s)*4 + c of row k = wq[k][r][s][c], zero elsewhere
// grouped: byte ((c*R + r)*8 + s) of row k = wq[k][r][s][c], zero elsewhere (im2col_grouped, conv_common.cuh)
__global__ void im2col_weights_kernel(const uint8_t* __restrict__ wq, uint8_t* __restrict__ wcol, int K, int C, int R, int S,
                                      int Cgp, int Kcol, int grouped) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K * Kcol) return;
    const int k = (int)(i / Kcol), b = (int)(i % Kcol);
    uint8_t v = 0;
    if (grouped) {
        const int gi = b >> 3, s = b & 7;
        const int c = gi / R, r = gi - c * R;
        if (c < C && s < S) v = wq[((int64_t)k * R * S + r * S + s) * Cgp + c];
    } else {
        const int tap = b >> 2, c = b & 3;
        if (tap < R * S && c < C) v = wq[((int64_t)k * R * S + tap) * Cgp + c];
    }
    wcol[i] = v;
}

// tap-major copy: wtap[((cb*taps + tap)*K + k)*KC + c] = wq[k][tap][cb*KC + c]
__global__ void tap_major_weights_kernel(const uint8_t* __restrict__ wq, uint8_t* __restrict__ wtap, int K, int taps, int Cgp,
                                         int KC) {
    const int64_t total = (int64_t)K * taps * Cgp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % KC);
    int64_t t = i / KC;
    const int k = (int)(t % K);
    t /= K;
    const int tap = (int)(t % taps);
    const int cb = (int)(t / taps);
    wtap[i] = wq[((int64_t)k * taps + tap) * Cgp + cb * KC + c];
}

// one warp per output channel
__global__ void tap_prefix_kernel(const uint8_t* __restrict__ wq, int32_t* __restrict__ wpre, int K, int R, int S,
                                  int Cgp, int is_signed) {
    const int k = blockIdx.x;
    const int lane = threadIdx.x;
    __shared__ int32_t T[16 * 16];
    for (int tap = 0; tap < R * S; ++tap) {
        const uint8_t* p = wq + ((int64_t)k * R * S + tap) * Cgp;
        int32_t acc = 0;
        for (int c = lane; c < Cgp; c += 32) acc += is_signed ? (int32_t)(int8_t)p[c] : (int32_t)p[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) T[tap] = acc;
    }
    __syncwarp();
    // exclusive prefix: Pre[r][s] = sum_{r'<r, s'<s} T[r'][s']
    int32_t* out = wpre + (int64_t)k * (R + 1) * (S + 1);
    for (int i = lane; i < (R + 1) * (S + 1); i += 32) {
        const int r = i / (S + 1), s = i % (S + 1);
        int32_t acc = 0;
        for (int rr = 0; rr < r; ++rr)
            for (int ss = 0; ss < s; ++ss) acc += T[rr * S + ss];
        out[i] = acc;
    }
}

}  // namespace
}  // namespace qb200

extern "C" {

int qb200_conv_out_hw(const qb200_conv_shape* s, int32_t* P, int32_t* Q) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    if (P) *P = (s->H + 2 * s->pad - s->R) / s->stride + 1;  // quantconv2d_float_input.cu:178
    if (Q) *Q = (s->W + 2 * s->pad - s->S) / s->stride + 1;  // :179
    return 0;
}

size_t qb200_conv_prepared_bytes(const qb200_conv_shape* s) {
    using namespace qb200;
    if (validate_shape(s)) return 0;
    return prepared_layout(*s).total;
}

size_t qb200_conv_workspace_bytes(const qb200_conv_shape* s) {
    using namespace qb200;
    if (validate_shape(s)) return 0;
    // the larger of: NHWC(Cp), zero-padded NHWC (halo path of stride-1 spatial kernels), im2col rows (few channels)
    size_t bytes = (size_t)s->N * (s->H + 2 * s->pad) * (s->W + 2 * s->pad) * qb200_padded_channels(s->C);
    if (uses_im2col_rows(*s)) {
        const size_t P = (s->H + 2 * s->pad - s->R) / s->stride + 1, Q = (s->W + 2 * s->pad - s->S) / s->stride + 1;
        const size_t col = (size_t)s->N * P * Q * im2col_kcol(s->C, s->R, s->S);
        if (col > bytes) bytes = col;
    }
    return align_up_sz(bytes, 256);
}

int qb200_conv_prepare_weights(const qb200_conv_shape* s, const uint8_t* w_packed, void* prepared, void* stream) {
    using namespace qb200;
    if (int rc = validate_shape(s)) return rc;
    QB_REQUIRE(w_packed && prepared, QB200_EINVAL, "prepare_weights: null pointer");
    QB_REQUIRE(reinterpret_cast<uintptr_t>(prepared) % 256 == 0, QB200_EINVAL, "prepare_weights: buffer must be 256-B aligned");
    const PreparedLayout L = prepared_layout(*s);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* wq = static_cast<uint8_t*>(prepared);
    int32_t* wpre = reinterpret_cast<int32_t*>(wq + L.wpre_off);
    const int64_t total = (int64_t)s->K * s->R * s->S * L.Cgp;
    const uint32_t offset = s->w_sign ? (1u << (s->w_bits - 1)) : 0u;  // quantconv2d_float_input.cu:185
    const int64_t n_bytes = qb200_packed_bytes((int64_t)s->K * s->Cg * s->R * s->S, s->w_bits);
    unpack_weights_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(w_packed, wq, s->K, s->Cg, s->R, s->S,
                                                                           L.Cgp, s->w_bits, offset, n_bytes);
    QB_LAUNCH_CHECK();
    tap_prefix_kernel<<<s->K, 32, 0, st>>>(wq, wpre, s->K, s->R, s->S, L.Cgp, s->w_sign ? 1 : 0);
    QB_LAUNCH_CHECK();
    if (L.tapKC) {
        const int64_t n = (int64_t)s->K * s->R * s->S * L.Cgp;
        tap_major_weights_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(wq, wq + L.wtap_off, s->K, s->R * s->S, L.Cgp, L.tapKC);
        QB_LAUNCH_CHECK();
    }
    if (L.Kcol) {
        im2col_weights_kernel<<<(unsigned)ceil_div64((int64_t)s->K * L.Kcol, 256), 256, 0, st>>>(
            wq, wq + L.wcol_off, s->K, s->C, s->R, s->S, L.Cgp, L.Kcol, im2col_grouped(s->C, s->R, s->S) ? 1 : 0);
        QB_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
