// Shared host/device helpers for the qb200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/qb200.h"

namespace qb200 {

// ---- host side: thread-local error string + launch counter -------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define QB_REQUIRE(cond, code, ...)                 \
    do {                                            \
        if (!(cond)) {                              \
            ::qb200::set_error(__VA_ARGS__);        \
            return (code);                          \
        }                                           \
    } while (0)

#define QB_CUDA(expr)                                                                     \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ::qb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                               __FILE__, __LINE__);                                       \
            (void)cudaGetLastError(); /* do not leave a non-sticky error for the next launch check */ \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

// after a <<<>>> launch
#define QB_LAUNCH_CHECK()                                                                 \
    do {                                                                                  \
        ::qb200::count_launch();                                                          \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            ::qb200::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),\
                               __FILE__, __LINE__);                                       \
            return (int)_e;                                                               \
        }                                                                                 \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int round_up_int(int a, int b) { return (a + b - 1) / b * b; }
static inline size_t align_up_sz(size_t a, size_t b) { return (a + b - 1) / b * b; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

#ifdef __CUDACC__
// Programmatic dependent launch: the engine's kernels run back to back on one stream (quantizer -> conv -> quantizer ...)
// and each is short (20-400 us), so the launch latency and the prologue of kernel i+1 (barrier init, TMEM allocation,
// tensor-map prefetch) are overlapped with the tail of kernel i.  Every kernel launched through launch_pdl executes
// pdl_wait() before its first access to global memory that an earlier kernel may have written (and before its first
// global write), and pdl_launch_dependents() as early as possible.  QB200_PDL=0 in the environment turns the attribute
// off (A/B measurements).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same, as a thread-block cluster of `cluster_x` CTAs (the CTA-pair tensor-core kernels)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                             int cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

}  // namespace qb200
