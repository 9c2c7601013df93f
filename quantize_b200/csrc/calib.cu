// Calibration reductions: the range estimators of the reference run, per layer and per calibration
// batch, x.min() + x.max() (or x.abs().max(); per channel: min/max over dim 1 of a transposed, flattened view) and — for
// percentile ranges — two torch.kthvalue calls (modelzoo/modules/range/minmax.py:62-108), then the running / moving-
// average update (:44-60, :184-203).  Here:
//   qb200_minmax_f32     one pass over the tensor: min, max and abs-max of every row, optional fused range update
//   qb200_kthvalue_f32   exact k-th smallest per row by 4-pass radix select on the order-preserving integer image of a float
// Both view the tensor as [A][R][B] and reduce over A and B: per tensor A=1,R=1,B=numel; weights per channel A=1,R=K,
// B=C*R*S; activations per channel A=N,R=C,B=H*W (no transposed copy is made).
// min / max / k-th value are selections, so results are bit-identical to torch's; NaN propagates like torch.min / max
// (any NaN -> NaN) and sorts last like torch.kthvalue.
#include <algorithm>
#include "common.cuh"

namespace qb200 {
namespace {

// order-preserving unsigned image of a float (NaN -> 0xFFFFFFFF, above +inf)
__device__ __forceinline__ uint32_t fkey(float f) {
    const uint32_t u = __float_as_uint(f);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0xFFFFFFFFu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
    if (k == 0xFFFFFFFFu) return __uint_as_float(0x7FC00000u);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
// |x| as an unsigned key (NaN -> 0xFFFFFFFF)
__device__ __forceinline__ uint32_t akey(float f) {
    const uint32_t u = __float_as_uint(f) & 0x7FFFFFFFu;
    return u > 0x7F800000u ? 0xFFFFFFFFu : u;
}
__device__ __forceinline__ float akey_inv(uint32_t k) { return k == 0xFFFFFFFFu ? __uint_as_float(0x7FC00000u) : __uint_as_float(k); }

constexpr int kRedThreads = 256;
constexpr int kRedItems = 16;   // float4 loads per thread per block-chunk

// state (unsigned, zero-initialised): [0][r] = max of fkey, [1][r] = max of ~fkey (i.e. min; a NaN must win here too, so
// NaN is mapped to 0xFFFFFFFF in this image as well), [2][r] = max of akey
__global__ void __launch_bounds__(kRedThreads)
minmax_kernel(const float* __restrict__ x, int64_t A, int64_t R, int64_t B, int64_t chunks_per_run, uint32_t* __restrict__ state) {
    pdl_launch_dependents();
    pdl_wait();
    // block -> (a, r, chunk of the contiguous run of B elements)
    const int64_t run = blockIdx.x / chunks_per_run, chunk = blockIdx.x - run * chunks_per_run;
    const int64_t a = run / R, r = run - a * R;
    const float* xr = x + (a * R + r) * B;
    const int64_t span = (int64_t)kRedThreads * kRedItems * 4;
    const int64_t e0 = chunk * span, e1 = min(B, e0 + span);
    uint32_t kmax = 0u, kmin = 0u, kabs = 0u;
    auto take = [&](float v) {
        const uint32_t k = fkey(v);
        kmax = max(kmax, k);
        kmin = max(kmin, k == 0xFFFFFFFFu ? k : ~k);
        kabs = max(kabs, akey(v));
    };
    if ((reinterpret_cast<uintptr_t>(xr) & 15) == 0) {
        const int64_t v0 = (e0 + 3) / 4 * 4;   // e0 is a multiple of 4 already
        for (int64_t i = v0 + (int64_t)threadIdx.x * 4; i + 3 < e1; i += (int64_t)kRedThreads * 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr + i));
            take(v.x); take(v.y); take(v.z); take(v.w);
        }
        const int64_t tail = e0 + (e1 - e0) / 4 * 4;
        for (int64_t i = tail + threadIdx.x; i < e1; i += kRedThreads) take(__ldg(xr + i));
    } else {
        for (int64_t i = e0 + threadIdx.x; i < e1; i += kRedThreads) take(__ldg(xr + i));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = max(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kabs = max(kabs, __shfl_xor_sync(0xffffffffu, kabs, o));
    }
    __shared__ uint32_t sm[3][kRedThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[0][warp] = kmax; sm[1][warp] = kmin; sm[2][warp] = kabs; }
    __syncthreads();
    if (threadIdx.x < 3) {
        uint32_t m = 0u;
        for (int w = 0; w < kRedThreads / 32; ++w) m = max(m, sm[threadIdx.x][w]);
        atomicMax(state + threadIdx.x * R + r, m);
    }
}

// decode + optional range update.  mode 0: plain; 1: running min / max (MinMax.update, minmax.py:44-60);
// 2: moving average  new = m * cur + (1 - m) * old  with torch's roundings (MAMinMax.update, :184-203)
__global__ void minmax_finalize_kernel(const uint32_t* __restrict__ state, int64_t R, int symmetric, float* __restrict__ out_min,
                                       float* __restrict__ out_max, int mode, float m, float one_minus_m, float* __restrict__ run_min,
                                       float* __restrict__ run_max) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const uint32_t kmax = state[r], kmin = state[R + r], kabs = state[2 * R + r];
    // symmetric ranges: xmin = 0, xmax = max |x|  (minmax.py:76-77, :90-91)
    float lo = symmetric ? 0.f : fkey_inv(kmin == 0xFFFFFFFFu ? kmin : ~kmin);
    float hi = symmetric ? akey_inv(kabs) : fkey_inv(kmax);
    if (mode == 1) {
        const float olo = run_min[r], ohi = run_max[r];
        lo = (lo != lo || olo != olo) ? __uint_as_float(0x7FC00000u) : fminf(olo, lo);   // torch.min / max propagate NaN
        hi = (hi != hi || ohi != ohi) ? __uint_as_float(0x7FC00000u) : fmaxf(ohi, hi);
    } else if (mode == 2) {
        lo = __fadd_rn(__fmul_rn(m, lo), __fmul_rn(one_minus_m, run_min[r]));
        hi = __fadd_rn(__fmul_rn(m, hi), __fmul_rn(one_minus_m, run_max[r]));
    }
    if (mode != 0) { run_min[r] = lo; run_max[r] = hi; }
    out_min[r] = lo;
    out_max[r] = hi;
}

// ---- radix select ----------------------------------------------------------------------------
// sel[r] = {prefix (key bits decided so far), k (rank still to find inside the prefix class)}; hist[r][256]
struct SelState {
    uint32_t prefix;
    uint32_t pad;
    unsigned long long k;
};

__global__ void __launch_bounds__(kRedThreads)
kth_hist_kernel(const float* __restrict__ x, int64_t A, int64_t R, int64_t B, int64_t chunks_per_run, int use_abs, int pass,
                const SelState* __restrict__ sel, unsigned long long* __restrict__ hist) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t run = blockIdx.x / chunks_per_run, chunk = blockIdx.x - run * chunks_per_run;
    const int64_t a = run / R, r = run - a * R;
    const float* xr = x + (a * R + r) * B;
    const int64_t span = (int64_t)kRedThreads * kRedItems * 4;
    const int64_t e0 = chunk * span, e1 = min(B, e0 + span);
    const int shift = 24 - 8 * pass;
    const uint32_t prefix = sel[r].prefix;
    const uint32_t hi_mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int64_t i = e0 + threadIdx.x; i < e1; i += kRedThreads) {
        const float v = __ldg(xr + i);
        const uint32_t k = use_abs ? akey(v) : fkey(v);
        if ((k & hi_mask) == (prefix & hi_mask)) atomicAdd(&h[(k >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    const uint32_t c = h[threadIdx.x];
    if (c) atomicAdd(hist + r * 256 + threadIdx.x, (unsigned long long)c);
}

// one block of 256 threads per row: find the bin holding rank k, narrow (prefix, k), clear the histogram for the next pass
__global__ void __launch_bounds__(256)
kth_pick_kernel(SelState* __restrict__ sel, unsigned long long* __restrict__ hist, int pass, int use_abs, float* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t r = blockIdx.x;
    __shared__ unsigned long long cnt[256];
    cnt[threadIdx.x] = hist[r * 256 + threadIdx.x];
    hist[r * 256 + threadIdx.x] = 0ull;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long cum = 0, k = sel[r].k;
        int b = 0;
        for (; b < 255; ++b) {
            if (cum + cnt[b] >= k) break;
            cum += cnt[b];
        }
        const int shift = 24 - 8 * pass;
        const uint32_t prefix = sel[r].prefix | ((uint32_t)b << shift);
        sel[r].prefix = prefix;
        sel[r].k = k - cum;
        if (pass == 3) out[r] = use_abs ? akey_inv(prefix) : fkey_inv(prefix);
    }
}

__global__ void kth_init_kernel(SelState* __restrict__ sel, unsigned long long* __restrict__ hist, int64_t R, unsigned long long k) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R) { sel[i].prefix = 0u; sel[i].pad = 0u; sel[i].k = k; }
    if (i < R * 256) hist[i] = 0ull;
}

int64_t chunks_of(int64_t B) { return std::max<int64_t>(1, ceil_div64(B, (int64_t)kRedThreads * kRedItems * 4)); }

}  // namespace
}  // namespace qb200

extern "C" {

size_t qb200_minmax_workspace_bytes(int64_t R) { return (size_t)(3 * (R > 0 ? R : 0)) * sizeof(uint32_t); }

int qb200_minmax_f32(const float* x, int64_t A, int64_t R, int64_t B, int32_t symmetric, float* out_min, float* out_max,
                     int32_t update_mode, float momentum, float one_minus_momentum, float* run_min, float* run_max,
                     void* workspace, void* stream) {
    using namespace qb200;
    QB_REQUIRE(A > 0 && R > 0 && B > 0, QB200_EINVAL, "minmax: empty tensor (torch.min of an empty tensor raises too)");
    QB_REQUIRE(x && out_min && out_max && workspace, QB200_EINVAL, "minmax: null pointer");
    QB_REQUIRE(update_mode >= 0 && update_mode <= 2 && (update_mode == 0 || (run_min && run_max)), QB200_EINVAL,
               "minmax: update_mode needs the running range");
    const int64_t cpr = chunks_of(B);
    QB_REQUIRE(A * R * cpr < (1ll << 31), QB200_EUNSUPPORTED, "minmax: too many blocks");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint32_t* state = static_cast<uint32_t*>(workspace);
    QB_CUDA(cudaMemsetAsync(state, 0, qb200_minmax_workspace_bytes(R), st));
    QB_CUDA(launch_pdl(minmax_kernel, dim3((unsigned)(A * R * cpr)), dim3(kRedThreads), 0, st, x, A, R, B, cpr, state));
    QB_LAUNCH_CHECK();
    QB_CUDA(launch_pdl(minmax_finalize_kernel, dim3((unsigned)ceil_div64(R, 128)), dim3(128), 0, st, (const uint32_t*)state, R,
                       (int)(symmetric != 0), out_min, out_max, (int)update_mode, momentum, one_minus_momentum, run_min, run_max));
    QB_LAUNCH_CHECK();
    return 0;
}

size_t qb200_kthvalue_workspace_bytes(int64_t R) { return (size_t)(R > 0 ? R : 0) * (16 + 256 * 8); }

int qb200_kthvalue_f32(const float* x, int64_t A, int64_t R, int64_t B, int32_t use_abs, int64_t k, float* out, void* workspace,
                       void* stream) {
    using namespace qb200;
    QB_REQUIRE(A > 0 && R > 0 && B > 0, QB200_EINVAL, "kthvalue: empty tensor");
    QB_REQUIRE(k >= 1 && k <= A * B, QB200_EINVAL, "kthvalue(): selected number k out of range for dimension");
    QB_REQUIRE(x && out && workspace, QB200_EINVAL, "kthvalue: null pointer");
    const int64_t cpr = chunks_of(B);
    QB_REQUIRE(A * R * cpr < (1ll << 31) && R < (1ll << 31), QB200_EUNSUPPORTED, "kthvalue: too many blocks");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SelState* sel = static_cast<SelState*>(workspace);
    unsigned long long* hist = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(workspace) + (size_t)R * 16);
    QB_CUDA(launch_pdl(kth_init_kernel, dim3((unsigned)ceil_div64(R * 256, 256)), dim3(256), 0, st, sel, hist, R, (unsigned long long)k));
    QB_LAUNCH_CHECK();
    for (int pass = 0; pass < 4; ++pass) {
        QB_CUDA(launch_pdl(kth_hist_kernel, dim3((unsigned)(A * R * cpr)), dim3(kRedThreads), 0, st, x, A, R, B, cpr, (int)(use_abs != 0),
                           pass, (const SelState*)sel, hist));
        QB_LAUNCH_CHECK();
        QB_CUDA(launch_pdl(kth_pick_kernel, dim3((unsigned)R), dim3(256), 0, st, sel, hist, pass, (int)(use_abs != 0), out));
        QB_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
