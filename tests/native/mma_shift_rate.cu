// Hardware probe (test infrastructure): RATE of tcgen05.mma kind::i8 when the A descriptor's start address is shifted by
// whole rows inside a TMA-written swizzled region (the halo variant of the conv kernel) vs aligned to the 8-row swizzle
// atom.  Prints cycles per MMA (M=128, N=BN, K=32) for several shifts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_shift_rate mma_shift_rate.cu && ./mma_shift_rate
// (derived from desc_shift.cu:)
// does a tcgen05 K-major shared-memory matrix descriptor whose START ADDRESS is
// shifted by whole rows inside a TMA-written swizzled region read the rows that live there?  (i.e. is the swizzle a
// function of the absolute shared-memory address, as for the +32-byte K advance?)  If yes, one TMA "halo" load can
// serve all 9 taps of a 3x3 convolution.  Prints one line per (swizzle, shift): OK / MISMATCH.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o desc_shift desc_shift.cu && ./desc_shift
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    for (int i = 0; i < 100000000 && !ok; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (!ok) __trap();
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo16, uint32_t layout) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}


constexpr int KC = 128;
template <int BN>
__global__ void __launch_bounds__(128, 1)
rate(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int shift, int reps, long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // 256 * KC
    uint8_t* sb = smem + 256 * KC;      // BN * KC
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + BN * KC);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 256 * KC + BN * KC);
        tma_2d(sa, &ta, bar, 0, 0);
        tma_2d(sa + 128 * KC, &ta, bar, 0, 128);
        for (int r = 0; r < BN; r += 64) tma_2d(sb + r * KC, &tb, bar, 0, 0);
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sbo16 = (8 * KC) >> 4;
        const uint64_t ad0 = make_desc(smem_u32(sa) + shift * KC, sbo16, 2), bd0 = make_desc(smem_u32(sb), sbo16, 2);
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            const uint64_t ad = ad0 + 2 * (i & 3), bd = bd0 + 2 * (i & 3);   // the four K slices of the 128-byte row
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
        mbar_wait(done, 0);
        *cycles = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BN>
void run(EncodeTiledFn enc) {
    std::vector<uint8_t> A(300 * KC, 1);
    std::vector<int8_t> B(64 * KC, 1);
    uint8_t *dA, *dB;
    long long* dC;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, 8));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    alignas(64) CUtensorMap ta, tb;
    cuuint64_t da[2] = {(cuuint64_t)KC, 300}, sa_[1] = {(cuuint64_t)KC}; cuuint32_t ba[2] = {(cuuint32_t)KC, 128}, es[2] = {1, 1};
    cuuint64_t db[2] = {(cuuint64_t)KC, 64}; cuuint32_t bb[2] = {(cuuint32_t)KC, 64};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, da, sa_, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, db, sa_, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("tensor map encode failed\n"); exit(1);
    }
    const size_t smem = 1024 + 256 * KC + BN * KC + 64;
    CK(cudaFuncSetAttribute(rate<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int reps = 4096;
    for (int shift : {0, 8, 16, 1, 2, 3, 4, 7, 30, 31, 58, 59, 60, 116}) {
        long long best = 1ll << 60;
        for (int t = 0; t < 3; ++t) {
            rate<BN><<<1, 128, smem>>>(ta, tb, shift, reps, dC);
            CK(cudaDeviceSynchronize());
            long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
            if (c < best) best = c;
        }
        printf("N=%3d shift %3d rows: %.1f cycles per MMA\n", BN, shift, (double)best / reps);
    }
}

int main() {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
    run<64>(enc);
    run<128>(enc);
    run<256>(enc);
    return 0;
}
