// Hardware probe (test infrastructure): what HBM delivers for the traffic MIXES of the hot path — a quantizer kernel is
// read-dominated (4 B in, 1 B out per element), a conv epilogue write-dominated (1 B in, 4 B out), the measured "copy
// peak" of MEASURED_PEAKS.json is 1:1.  Prints GB/s for read-only, write-only, copy and 4:1 / 1:4 mixes at 2 GB, so that a
// layer's achieved bandwidth can be judged against what its own read/write ratio can reach.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hbm_mix hbm_mix.cu && ./hbm_mix
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// each thread: R float4 loads and W float4 stores per iteration (all loads issued first)
template <int R, int W>
__global__ void __launch_bounds__(256) mix(const float4* __restrict__ in, float4* __restrict__ out, size_t n_iters, float4* sink) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t it = 0; it < n_iters; ++it) {
        float4 v[R > 0 ? R : 1];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4* p = in + (it * R + r) * nthreads + tid;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[r].x), "=f"(v[r].y), "=f"(v[r].z), "=f"(v[r].w) : "l"(p));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
#pragma unroll
        for (int w = 0; w < W; ++w) out[(it * W + w) * nthreads + tid] = (R > 0) ? v[w % (R > 0 ? R : 1)] : acc;
    }
    if (acc.x == 123.456f) *sink = acc;   // keep the loads alive
}

template <int R, int W>
void run(const char* name, float4* a, float4* b, size_t bytes_each, float4* sink) {
    const int blocks = 148 * 16, threads = 256;
    const size_t nthreads = (size_t)blocks * threads;
    const size_t per_iter = nthreads * 16 * (size_t)(R > W ? R : W);
    const size_t n_iters = bytes_each / per_iter;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int t = 0; t < 6; ++t) {
        CK(cudaEventRecord(e0));
        mix<R, W><<<blocks, threads>>>(a, b, n_iters, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (t > 0 && ms < best) best = ms;
    }
    const double moved = (double)n_iters * nthreads * 16 * (R + W);
    printf("%-12s read:write %d:%d  %7.1f GB/s  (%.1f MB in %.3f ms)\n", name, R, W, moved / best / 1e6, moved / 1e6, best);
}

int main() {
    const size_t bytes = (size_t)2 << 30;
    float4 *a, *b, *sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
    run<4, 0>("read-only", a, b, bytes, sink);
    run<0, 4>("write-only", a, b, bytes, sink);
    run<4, 4>("copy", a, b, bytes, sink);
    run<4, 1>("quantizer", a, b, bytes, sink);
    run<1, 4>("epilogue", a, b, bytes, sink);
    run<2, 4>("1x1 in<out", a, b, bytes, sink);
    // cudaMemset / cudaMemcpy for reference
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int t = 0; t < 3; ++t) {
        CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(b, 0, bytes)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (t == 2) printf("cudaMemset   %7.1f GB/s\n", bytes / ms / 1e6);
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (t == 2) printf("cudaMemcpy   %7.1f GB/s (read + write)\n", 2.0 * bytes / ms / 1e6);
    }
    return 0;
}
