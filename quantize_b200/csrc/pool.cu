// Max pooling over fp32 NCHW planes — the one non-conv op between the quantized convs of the ResNet family
// (stem conv -> relu -> maxpool 3x3/2 -> layer1).  It is memory-bound (read the plane once, write a quarter of it), so
// the kernel stages a band of input rows in shared memory with coalesced 16-byte loads and every output is computed
// from shared memory; torch's one-thread-per-output kernel runs at ~1.5 TB/s on B200, this one at the HBM rate.
// Results are bit-identical to torch.nn.functional.max_pool2d (max is exact; padding behaves as -inf).
#include <algorithm>
#include <cstdlib>
#include <cfloat>
#include <atomic>
#include "common.cuh"

namespace qb200 {
namespace {

constexpr int kPoolThreads = 256;

// max that propagates NaN like torch's pooling (one instruction)
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// one block = one (plane, band of output rows); the band's input rows [h_lo, h_hi) are staged in shared memory
// KT / ST: compile-time kernel size and stride (0 = run-time values): with run-time loop bounds the pooling loops
// compiled to ~50 instructions per output and the kernel was issue-bound (ncu: issue slots 89 % busy at 2.4 TB/s).
template <int KT, int ST>
__global__ void __launch_bounds__(kPoolThreads)
maxpool2d_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, int P, int Q, int k_rt, int stride_rt, int pad,
                 int band_rows, int bands, int in_rows_alloc) {
    pdl_launch_dependents();
    pdl_wait();
    const int k = KT ? KT : k_rt, stride = ST ? ST : stride_rt;
    extern __shared__ float tile[];
    const int plane = blockIdx.x / bands, band = blockIdx.x - plane * bands;
    const int p0 = band * band_rows, p1 = min(P, p0 + band_rows);
    const int h_lo = max(0, p0 * stride - pad), h_hi = min(H, (p1 - 1) * stride - pad + k);
    const float* xp = x + (int64_t)plane * H * W + (int64_t)h_lo * W;
    const int n_in = (h_hi - h_lo) * W;
    if ((reinterpret_cast<uintptr_t>(xp) & 15) == 0 && (n_in & 3) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(xp);
        float4* t4 = reinterpret_cast<float4*>(tile);
        const int n4 = n_in >> 2;
        // eight independent 16-byte loads per thread in flight before the first shared-memory store: a loop of
        // load -> store pairs serialises one DRAM round trip per iteration and the block spends its life waiting
        for (int base = 0; base < n4; base += 8 * kPoolThreads) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * kPoolThreads + threadIdx.x;
                if (i < n4)
                    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                        : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(x4 + i));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = base + u * kPoolThreads + threadIdx.x;
                if (i < n4) t4[i] = v[u];
            }
        }
    } else {
        for (int i = threadIdx.x; i < n_in; i += kPoolThreads) tile[i] = __ldg(xp + i);
    }
    __syncthreads();
    // separable: each warp takes output rows p0 + warp, p0 + warp + 8, ...; it first reduces the k input rows of an
    // output row column-wise (consecutive lanes = consecutive columns: conflict-free) into its own row of vbuf, then
    // reduces k neighbours of that row per output column.  Only the warp itself reads its vbuf row -> __syncwarp.
    float* vbuf = tile + (size_t)in_rows_alloc * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* op = out + (int64_t)plane * P * Q;
    for (int p = p0 + warp; p < p1; p += kPoolThreads / 32) {
        const int h0 = p * stride - pad;
        const int r_lo = max(0, -h0), r_hi = min(k, H - h0);
        float* vrow = vbuf + (size_t)(p - p0) * W;
        const float* trow = tile + (h0 - h_lo) * W;   // row r of the window = trow + r * W (only rows r_lo..r_hi-1 are staged)
        for (int w = lane; w < W; w += 32) {
            float m = -INFINITY;
            if (KT) {
#pragma unroll
                for (int r = 0; r < (KT ? KT : 1); ++r)
                    if (r >= r_lo && r < r_hi) m = max_nan(m, trow[r * W + w]);
            } else {
                for (int r = r_lo; r < r_hi; ++r) m = max_nan(m, trow[r * W + w]);
            }
            vrow[w] = m;
        }
        __syncwarp();
        for (int q = lane; q < Q; q += 32) {
            const int w0 = q * stride - pad;
            float m = -INFINITY;
            if (KT) {
#pragma unroll
                for (int s2 = 0; s2 < (KT ? KT : 1); ++s2)
                    if (w0 + s2 >= 0 && w0 + s2 < W) m = max_nan(m, vrow[w0 + s2]);
            } else {
                for (int s2 = max(0, -w0); s2 < min(k, W - w0); ++s2) m = max_nan(m, vrow[w0 + s2]);
            }
            op[(int64_t)p * Q + q] = m;
        }
    }
}

// 3x3 / stride 2 / pad 1 on even planes (the ResNet stem's pool), streaming form: a lane owns one output column and walks
// down the plane; per input row it loads the 8 bytes of columns (2q, 2q + 1), takes column 2q - 1 from its left neighbour by
// shuffle, and keeps the row maximum of the previous odd row as the carry into the next output row.  Every input element
// is read from HBM exactly once, nothing is staged in shared memory, the loads of four output rows (eight rows of 256
// contiguous bytes per warp) are in flight before the first max: the band kernel above spent 30 % of its time in its
// shared-memory passes (245 us on 256 x 64 x 112 x 112 against ~170 us of HBM time).
constexpr int kStreamRows = 4;   // output rows per batch of loads
__global__ void __launch_bounds__(256)
maxpool3x3s2_stream_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t planes, int H, int W, int P, int Q, int groups) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (plane, column group)
    if (wid >= planes * groups) return;
    const int64_t plane = wid / groups;
    const int q = (int)(wid - plane * groups) * 32 + lane;
    const bool act = q < Q;
    const float* xp = x + plane * H * W + 2 * q;
    float* op = out + plane * P * Q + q;
    const bool edge = lane == 0 && q > 0;                  // the group's first lane has no left neighbour in the warp
    float carry = -INFINITY;                                // row maximum of input row 2p - 1
    auto row_max = [&](float2 v, float left) { return max_nan(max_nan(left, v.x), v.y); };
    for (int p0 = 0; p0 < P; p0 += kStreamRows) {
        float2 v[2 * kStreamRows];
        float e[2 * kStreamRows];
#pragma unroll
        for (int u = 0; u < 2 * kStreamRows; ++u) {
            const int h = 2 * p0 + u;
            v[u] = make_float2(-INFINITY, -INFINITY);
            e[u] = -INFINITY;
            if (act && h < H) {
                asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[u].x), "=f"(v[u].y) : "l"(xp + (int64_t)h * W));
                if (edge) e[u] = __ldg(xp + (int64_t)h * W - 1);
            }
        }
#pragma unroll
        for (int u = 0; u < kStreamRows; ++u) {
            float l0 = __shfl_up_sync(0xffffffffu, v[2 * u].y, 1), l1 = __shfl_up_sync(0xffffffffu, v[2 * u + 1].y, 1);
            if (lane == 0) { l0 = e[2 * u]; l1 = e[2 * u + 1]; }
            const float m0 = row_max(v[2 * u], l0), m1 = row_max(v[2 * u + 1], l1);
            const int p = p0 + u;
            if (act && p < P) op[(int64_t)p * Q] = max_nan(max_nan(carry, m0), m1);
            carry = m1;
        }
    }
}

// Global average pooling (the op between the last residual stage and the classifier).  Small planes (7x7 = 196 bytes) are
// a latency problem, not a bandwidth one: a warp that owns one plane has 196 bytes in flight.  So four lanes share a plane
// (a warp covers 8 consecutive planes = one contiguous span), every lane issues all its loads before the first add, and two
// shuffle steps finish the sum; sum / (H*W).  torch's generic reduction took 115 us for ResNet-50's 256 x 2048 planes of 49
// values (0.9 TB/s); one warp per plane 78 us.
__global__ void __launch_bounds__(256)
avgpool_global_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t planes, int HW) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, sub = lane & 3;
    const int64_t plane = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 8 + (lane >> 2);
    float s = 0.f;
    if (plane < planes) {
        const float* p = x + plane * HW;
        int i = sub;
        for (; i + 28 < HW; i += 32) {          // 8 independent loads per round
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(p + i + 4 * u);
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; i < HW; i += 4) s += __ldg(p + i);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (sub == 0 && plane < planes) out[plane] = s / (float)HW;
}

}  // namespace
}  // namespace qb200

extern "C" int qb200_avgpool_global_f32(const float* x, int64_t planes, int32_t HW, float* out, void* stream) {
    using namespace qb200;
    QB_REQUIRE(x && out, QB200_EINVAL, "avgpool_global: null pointer");
    QB_REQUIRE(HW >= 1, QB200_EINVAL, "avgpool_global: empty planes");
    if (planes == 0) return 0;
    QB_REQUIRE((planes + 63) / 64 < (1ll << 31), QB200_EINVAL, "avgpool_global: too many planes");
    QB_CUDA(launch_pdl(avgpool_global_kernel, dim3((unsigned)((planes + 63) / 64)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out,
                       planes, (int)HW));
    QB_LAUNCH_CHECK();
    return 0;
}

extern "C" int qb200_maxpool2d_f32(const float* x, int64_t planes, int32_t H, int32_t W, int32_t kernel, int32_t stride,
                                   int32_t pad, float* out, void* stream) {
    using namespace qb200;
    QB_REQUIRE(x && out, QB200_EINVAL, "maxpool2d: null pointer");
    QB_REQUIRE(kernel >= 1 && stride >= 1 && pad >= 0 && 2 * pad <= kernel && H >= 1 && W >= 1, QB200_EINVAL,
               "maxpool2d: bad geometry");
    const int P = (H + 2 * pad - kernel) / stride + 1, Q = (W + 2 * pad - kernel) / stride + 1;
    QB_REQUIRE(P >= 1 && Q >= 1, QB200_EINVAL, "maxpool2d: empty output");
    if (planes == 0) return 0;
    static const bool stream_on = [] {      // QB200_POOL_STREAM=0: the band kernel everywhere (A/B measurements)
        const char* e = getenv("QB200_POOL_STREAM");
        return !(e && e[0] == '0');
    }();
    if (stream_on && kernel == 3 && stride == 2 && pad == 1 && H % 2 == 0 && W % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 8 == 0) {
        const int groups = (Q + 31) / 32;
        const int64_t warps = planes * groups;
        QB_REQUIRE((warps + 7) / 8 < (1ll << 31), QB200_EINVAL, "maxpool2d: too many blocks");
        QB_CUDA(launch_pdl(maxpool3x3s2_stream_kernel, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out,
                           planes, (int)H, (int)W, P, Q, groups));
        QB_LAUNCH_CHECK();
        return 0;
    }
    // as many output rows per block as keep the staged input rows within 32 KB (+ the column-reduced rows: ~5 blocks per SM)
    static const int stage_kb = [] {      // QB200_POOL_STAGE_KB: staged input per block (A/B measurements)
        const char* e = getenv("QB200_POOL_STAGE_KB");
        return e ? std::max(atoi(e), 1) : 32;
    }();
    const int max_in_rows = std::max(kernel, (stage_kb * 1024) / (W * 4));
    QB_REQUIRE((size_t)kernel * W * 4 <= 200 * 1024, QB200_EUNSUPPORTED, "maxpool2d: rows wider than shared memory allows");
    int band_rows = std::max(1, (max_in_rows - kernel) / stride + 1);
    band_rows = std::min(band_rows, P);
    int bands = (P + band_rows - 1) / band_rows;
    band_rows = (P + bands - 1) / bands;  // balance the bands
    bands = (P + band_rows - 1) / band_rows;
    const int in_rows = std::min(H, (band_rows - 1) * stride + kernel);
    const size_t smem = ((size_t)in_rows + band_rows) * W * 4;  // staged input rows + one column-reduced row per output row
    QB_REQUIRE(planes * bands < (1ll << 31), QB200_EINVAL, "maxpool2d: too many blocks");
    // the shared-memory attribute is per (function, device): set it once for every instantiation on each device
    static std::atomic<uint64_t> attr_mask{0};
    int dev = 0;
    QB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !((attr_mask.load(std::memory_order_acquire) >> dev) & 1ull)) {
        QB_CUDA(cudaFuncSetAttribute(maxpool2d_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        QB_CUDA(cudaFuncSetAttribute(maxpool2d_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        QB_CUDA(cudaFuncSetAttribute(maxpool2d_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        QB_CUDA(cudaFuncSetAttribute(maxpool2d_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        if (dev >= 0 && dev < 64) attr_mask.fetch_or(1ull << dev, std::memory_order_release);
    }
    auto launch = [&](auto kern) -> int {
        QB_CUDA(launch_pdl(kern, dim3((unsigned)(planes * bands)), dim3(kPoolThreads), smem, static_cast<cudaStream_t>(stream),
                           x, out, H, W, P, Q, kernel, stride, pad, band_rows, bands, in_rows));
        return 0;
    };
    if (kernel == 3 && stride == 2) { if (int rc = launch(maxpool2d_kernel<3, 2>)) return rc; }
    else if (kernel == 2 && stride == 2) { if (int rc = launch(maxpool2d_kernel<2, 2>)) return rc; }
    else if (kernel == 3 && stride == 1) { if (int rc = launch(maxpool2d_kernel<3, 1>)) return rc; }
    else { if (int rc = launch(maxpool2d_kernel<0, 0>)) return rc; }
    QB_LAUNCH_CHECK();
    return 0;
}
