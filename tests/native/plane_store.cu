// Hardware probe (test infrastructure): what does the NCHW epilogue store pattern cost when channel planes are small and
// not sector-aligned?  The conv epilogue writes, per warp and output channel, 32 consecutive pixels (128 B) of one plane;
// planes are P*Q*4 bytes apart: 56x56 -> 12544 B, 28x28 -> 3136 B, 14x14 -> 784 B (= 24.5 sectors), 7x7 -> 196 B.
// This probe writes a [N][K][PQ] fp32 tensor exactly that way (tiles of 128 consecutive flat pixels x 32-channel chunks,
// lane = pixel) for several PQ, and — for comparison — the same bytes "image-major": the [32 ch][PQ] block of one image is
// one contiguous run, written linearly with 16-byte vectors.  Prints GB/s of each.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o plane_store plane_store.cu && ./plane_store
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// one warp = 32 consecutive flat pixels (rows m..m+31) x all K channels in chunks of 32, like the epilogue
__global__ void __launch_bounds__(256) epilogue_like(float* __restrict__ out, int N, int K, int PQ) {
    const long long M = (long long)N * PQ;
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long w = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); w * 32 < M; w += (long long)gridDim.x * warps_per_block) {
        const long long m = w * 32 + lane;
        if (m >= M) continue;
        const int img = (int)(m / PQ), pq = (int)(m - (long long)img * PQ);
        float* o = out + ((long long)img * K) * PQ + pq;
        for (int k = 0; k < K; k += 4) {
            o[(long long)k * PQ] = (float)k;
            o[(long long)(k + 1) * PQ] = (float)k;
            o[(long long)(k + 2) * PQ] = (float)k;
            o[(long long)(k + 3) * PQ] = (float)k;
        }
    }
}

// image-major: a block writes the contiguous [32 ch][PQ] run of (image, chunk) with 16-byte vectors
__global__ void __launch_bounds__(256) linear_runs(float4* __restrict__ out, long long n_runs, int run_vec4) {
    for (long long r = blockIdx.x; r < n_runs; r += gridDim.x) {
        float4* o = out + r * run_vec4;
        for (int i = threadIdx.x; i < run_vec4; i += blockDim.x) o[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
}

int main() {
    const int K = 1024;
    float* buf;
    const size_t max_bytes = (size_t)1 << 30;
    CK(cudaMalloc(&buf, max_bytes + 4096));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int hw : {56, 28, 14, 7, 16, 8}) {
        const int PQ = hw * hw;
        const int N = (int)(max_bytes / ((size_t)K * PQ * 4));
        const double bytes = (double)N * K * PQ * 4;
        float best = 1e30f;
        for (int t = 0; t < 5; ++t) {
            CK(cudaEventRecord(e0));
            epilogue_like<<<148 * 8, 256>>>(buf, N, K, PQ);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (t && ms < best) best = ms;
        }
        float best2 = 1e30f;
        const int run_vec4 = 32 * PQ / 4;   // [32 ch][PQ] floats (PQ*32 is a multiple of 4)
        const long long n_runs = (long long)N * (K / 32);
        for (int t = 0; t < 5; ++t) {
            CK(cudaEventRecord(e0));
            linear_runs<<<148 * 8, 256>>>(reinterpret_cast<float4*>(buf), n_runs, run_vec4);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (t && ms < best2) best2 = ms;
        }
        printf("%2dx%-2d planes of %5d B: epilogue-like stores %7.1f GB/s   image-major 16-byte runs %7.1f GB/s   (%.0f MB)\n", hw, hw,
               PQ * 4, bytes / best / 1e6, bytes / best2 / 1e6, bytes / 1e6);
    }
    return 0;
}
