// Hardware probe (test infrastructure): does a tcgen05 K-major shared-memory matrix descriptor whose START ADDRESS is
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

// A region: 256 rows x KC bytes (two TMA boxes of 128 rows), B: 64 rows x KC.  D[128 x 64] = A[shift .. shift+128) * B^T
template <int KC>
__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int shift, int32_t* out, uint32_t layout) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // 256 * KC
    uint8_t* sb = smem + 256 * KC;      // 64 * KC
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * KC);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 256 * KC + 64 * KC);
        tma_2d(sa, &ta, bar, 0, 0);
        tma_2d(sa + 128 * KC, &ta, bar, 0, 128);
        tma_2d(sb, &tb, bar, 0, 0);
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sbo16 = (8 * KC) >> 4;
        for (int k = 0; k < KC / 32; ++k) {
            const uint64_t ad = make_desc(smem_u32(sa) + shift * KC + 32 * k, sbo16, layout);
            const uint64_t bd = make_desc(smem_u32(sb) + 32 * k, sbo16, layout);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)(k != 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = threadIdx.x;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[row * 64 + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int KC>
int run(EncodeTiledFn enc, CUtensorMapSwizzle swz, uint32_t layout, const char* name) {
    std::vector<uint8_t> A(300 * KC);
    std::vector<int8_t> B(64 * KC);
    srand(7);
    for (auto& v : A) v = (uint8_t)(rand() & 0xFF);
    for (auto& v : B) v = (int8_t)((rand() & 0xFF) - 128);
    uint8_t *dA, *dB;
    int32_t* dO;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dO, 128 * 64 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    alignas(64) CUtensorMap ta, tb;
    cuuint64_t da[2] = {(cuuint64_t)KC, 300}, sa_[1] = {(cuuint64_t)KC}; cuuint32_t ba[2] = {(cuuint32_t)KC, 128}, es[2] = {1, 1};
    cuuint64_t db[2] = {(cuuint64_t)KC, 64}; cuuint32_t bb[2] = {(cuuint32_t)KC, 64};
    if (enc(&ta, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dA, da, sa_, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        enc(&tb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dB, db, sa_, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n"); return 1;
    }
    const int smem = 256 * KC + 64 * KC + 1024 + 64;
    CK(cudaFuncSetAttribute(probe<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int bad_total = 0;
    const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 58, 59, 60, 116, 117, 118, 127, 128};
    for (int shift : shifts) {
        probe<KC><<<1, 128, smem>>>(ta, tb, shift, dO, layout);
        CK(cudaDeviceSynchronize());
        std::vector<int32_t> O(128 * 64);
        CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                int32_t acc = 0;
                for (int k = 0; k < KC; ++k) acc += (int32_t)A[(size_t)(m + shift) * KC + k] * (int32_t)B[(size_t)n * KC + k];
                bad += acc != O[m * 64 + n];
            }
        printf("%s shift %3d rows: %s (%d mismatches)\n", name, shift, bad ? "MISMATCH" : "OK", bad);
        bad_total += bad;
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return bad_total != 0;
}

int main() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
    int rc = 0;
    rc |= run<128>(enc, CU_TENSOR_MAP_SWIZZLE_128B, 2, "SW128 KC=128");
    rc |= run<64>(enc, CU_TENSOR_MAP_SWIZZLE_64B, 4, "SW64  KC=64 ");
    rc |= run<32>(enc, CU_TENSOR_MAP_SWIZZLE_32B, 6, "SW32  KC=32 ");
    printf(rc ? "RESULT: row-shifted descriptors do NOT address TMA-swizzled rows\n" : "RESULT: row-shifted descriptors work\n");
    return 0;
}
