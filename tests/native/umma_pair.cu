// Hardware probe + whole-chip int8 tensor peak (test infrastructure / profiles/int8_peak.json):
// a persistent u8 x s8 -> s32 GEMM  D[M][N] = A[M][K] * B[N][K]^T  on tcgen05, written two ways from the same source:
//   CG = 1   one CTA per tile, 128 x 256, tcgen05.mma.cta_group::1 (what conv_umma_kernel issues)
//   CG = 2   a CTA PAIR per 256 x 256 tile, tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of
//            the B tile (128 of the 256 weight rows); the pair's tensor cores read both halves, so the operand bytes
//            per MAC drop by a third (A 128 + B 128 rows per CTA and k-block instead of 128 + 256)
// It checks the result of a small problem against the host and then times 8192^3.  The cta_group::2 protocol proven here
// (cluster-scope TMA completion on the leader's barrier, multicast commits, remote accumulator-empty arrivals) is the one
// conv_umma.cu uses for deep reductions.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_pair umma_pair.cu && ./umma_pair [json-out]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int BM = 128, BN = 256, KC = 128;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 64 + kEpiWarps * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_rank_or_zero() { return cluster_rank(); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");     // (not .aligned: warps may arrive diverged)
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, int code, int* err) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!ok && clock64() - t0 > 3000000000ll) {     // ~1.5 s: record which wait expired and abort instead of hanging
            if (err) atomicExch(err, code + 16 * (int)cluster_rank_or_zero());
            __threadfence_system();
            __trap();
        }
    }
}
// TMA tile load whose completion is signalled on a barrier given by its shared::cluster address (CG = 2: the leader's)
template <int CG>
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int x, int y) {
    if constexpr (CG == 2)
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(x), "r"(y) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo16, uint32_t layout) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of the pair once the MMAs issued so far have completed
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (CG == 2)
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(kThreads, 1)
gemm_i8(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb, int M, int N, int K, int stages,
        int32_t* __restrict__ D, int store, int* err) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t a_bytes = BM * KC, b_rows = BN / CG, b_bytes = b_rows * KC, stage_bytes = a_bytes + b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)stages * stage_bytes);
    uint64_t* full = bars;             // [stages]   (CG = 2: only the leader's are waited on)
    uint64_t* empty = bars + 16;       // [stages]
    uint64_t* acc_full = bars + 32;    // [2]
    uint64_t* acc_empty = bars + 34;   // [2]        (CG = 2: only the leader's are waited on)
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 36);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_rank() : 0;
    const bool leader = rank == 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;

    const int n_tiles = N / BN, m_tiles = M / (BM * CG);      // a "tile" is what one CTA (CG 1) or one pair (CG 2) computes
    const int total = m_tiles * n_tiles, kblocks = K / KC;
    const int unit = blockIdx.x / CG, n_units = gridDim.x / CG;

    if (warp == 0) {
      if (lane == 0) {
        // ---- TMA producer (both CTAs of a pair: own A rows, own half of B) ----
        int stage = 0; uint32_t phase = 0;
        for (int t = unit; t < total; t += n_units) {
            const int mt = t / n_tiles, nt = t - mt * n_tiles;
            const int row0 = (mt * CG + (int)rank) * BM, col0 = nt * BN + (int)rank * (int)b_rows;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1, 1, err);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                uint32_t bar = smem_u32(&full[stage]);
                if constexpr (CG == 2) {
                    bar = map_to_cta(bar, 0);                                          // the leader's barrier
                    if (leader) mbar_expect_tx(&full[stage], 2 * stage_bytes);         // bytes of BOTH CTAs
                } else {
                    mbar_expect_tx(&full[stage], stage_bytes);
                }
                tma_2d<CG>(sa, &ta, bar, kb * KC, row0);
                tma_2d<CG>(sa + a_bytes, &tb, bar, kb * KC, col0);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issuer (leader CTA only) ----
        if (leader && lane == 0) {
            const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
            const uint32_t sbo16 = (8 * KC) >> 4;
            int stage = 0, buf = 0; uint32_t phase = 0, acc_phase = 0;
            for (int t = unit; t < total; t += n_units) {
                mbar_wait(&acc_empty[buf], acc_phase ^ 1, 2, err);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem + (uint32_t)(buf * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase, 3, err);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t ad = make_desc(sa, sbo16, 2), bd = make_desc(sa + a_bytes, sbo16, 2);
#pragma unroll
                    for (int k = 0; k < KC / 32; ++k) umma<CG>(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit<CG>(&empty[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                umma_commit<CG>(&acc_full[buf]);
                if (++buf == 2) { buf = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: TMEM -> registers -> D (row-major int32); every CTA drains its own 128 rows ----
        const int quad = warp & 3;
        int buf = 0; uint32_t acc_phase = 0;
        const uint32_t acc_empty_leader = CG == 2 ? map_to_cta(smem_u32(&acc_empty[0]), 0) : smem_u32(&acc_empty[0]);
        long long sink = 0;
        for (int t = unit; t < total; t += n_units) {
            const int mt = t / n_tiles, nt = t - mt * n_tiles;
            const int row = (mt * CG + (int)rank) * BM + quad * 32 + lane;
            mbar_wait(&acc_full[buf], acc_phase, 4, err);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * BN + c0), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (store) {
                    int4* o = reinterpret_cast<int4*>(D + (size_t)row * N + nt * BN + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = make_int4((int)v[4 * j], (int)v[4 * j + 1], (int)v[4 * j + 2], (int)v[4 * j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) sink += (int)v[j];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_leader + (uint32_t)buf * 8u);
            if (++buf == 2) { buf = 0; acc_phase ^= 1; }
        }
        if (!store && sink == 0x7fffffffffffll) D[0] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1) {
        if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeTiledFn enc, CUtensorMap* m, void* p, int rows, int K, int box_rows) {
    cuuint64_t d[2] = {(cuuint64_t)K, (cuuint64_t)rows}, s[1] = {(cuuint64_t)K};
    cuuint32_t b[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows}, e[2] = {1, 1};
    if (enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, p, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("tensor map encode failed\n");
        exit(1);
    }
}

template <int CG>
double run(EncodeTiledFn enc, int M, int N, int K, bool check, int reps, int* err_dev, int* err_host, int max_stages = 16) {
    std::vector<uint8_t> A((size_t)M * K);
    std::vector<int8_t> B((size_t)N * K);
    uint32_t s = 12345u + CG;
    for (auto& v : A) { s = s * 1664525u + 1013904223u; v = (uint8_t)(s >> 24); }
    for (auto& v : B) { s = s * 1664525u + 1013904223u; v = (int8_t)((s >> 24) % 255 - 127); }
    uint8_t *dA, *dB; int32_t* dD;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
    alignas(64) CUtensorMap ta, tb;
    make_map(enc, &ta, dA, M, K, BM);
    make_map(enc, &tb, dB, N, K, BN / CG);
    const int stage_bytes = (BM + BN / CG) * KC;
    int stages = (200 * 1024) / stage_bytes; if (stages > 16) stages = 16;
    if (stages > max_stages) stages = max_stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + 512;
    CK(cudaFuncSetAttribute(gemm_i8<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int total = (M / (BM * CG)) * (N / BN);
    int grid = sms / CG * CG; if (grid > total * CG) grid = total * CG;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    auto launch = [&](int store) { CK(cudaLaunchKernelEx(&cfg, gemm_i8<CG>, ta, tb, M, N, K, stages, dD, store, err_dev)); };
    launch(1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CG=%d kernel failed: %s (watchdog code %d)\n", CG, cudaGetErrorString(e), *err_host); exit(2); }
    if (check) {
        std::vector<int32_t> D((size_t)M * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        long long bad = 0;
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n) {
                int32_t acc = 0;
                for (int k = 0; k < K; ++k) acc += (int32_t)A[(size_t)m * K + k] * (int32_t)B[(size_t)n * K + k];
                if (acc != D[(size_t)m * N + n]) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %d\n", m, n, D[(size_t)m * N + n], acc); ++bad; }
            }
        printf("CG=%d  %dx%dx%d  %s (%lld mismatches)\n", CG, M, N, K, bad ? "MISMATCH" : "OK", bad);
        if (bad) exit(3);
    }
    double best = 0;
    if (reps > 0) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int st = 1; st >= 0; --st) {
            float bt = 1e30f;
            for (int t = 0; t < reps; ++t) {
                CK(cudaEventRecord(e0)); launch(st); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < bt) bt = ms;
            }
            const double tops = 2.0 * M * N * K / (bt * 1e-3) / 1e12;
            printf("CG=%d  %dx%dx%d  %s: %.3f ms  %.1f TOPS (best of %d, grid %d, %d stages)\n", CG, M, N, K,
                   st ? "with int32 D stores" : "accumulate only", bt, tops, reps, grid, stages);
            if (st == 1) best = tops;
        }
    }
    CK(cudaFree(dA)); CK(cudaFree(dB)); CK(cudaFree(dD));
    return best;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IONBF, 0);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
    int *eh, *ed;
    CK(cudaHostAlloc(&eh, 4, cudaHostAllocMapped)); *eh = 0;
    CK(cudaHostGetDevicePointer(&ed, eh, 0));
    const bool only2 = argc > 2 && argv[2][0] == '2';
    if (!only2) run<1>(enc, 512, 512, 384, true, 0, ed, eh);
    run<2>(enc, 512, 512, 384, true, 0, ed, eh);
    run<2>(enc, 1024, 768, 1024, true, 0, ed, eh);
    const double t1 = run<1>(enc, 8192, 8192, 8192, false, 5, ed, eh);
    const double t2 = run<2>(enc, 8192, 8192, 8192, false, 5, ed, eh);
    if (argc > 3) {   // pipeline-depth sweep: is the main loop bound by bytes in flight?
        for (int st = 2; st <= 4; ++st) run<1>(enc, 8192, 8192, 8192, false, 3, ed, eh, st);
        for (int st = 2; st <= 6; ++st) run<2>(enc, 8192, 8192, 8192, false, 3, ed, eh, st);
    }
    if (argc > 1) {
        FILE* o = fopen(argv[1], "w");
        if (o) {
            fprintf(o, "{\"what\": \"whole-chip u8 x s8 -> s32 GEMM 8192^3 on tcgen05 (tests/native/umma_pair.cu), int32 D written, best of 5\", "
                       "\"tops_cta_group1\": %.1f, \"tops_cta_group2\": %.1f, \"tops\": %.1f, \"spec_tops\": 4500}\n", t1, t2, t1 > t2 ? t1 : t2);
            fclose(o);
        }
    }
    return 0;
}
