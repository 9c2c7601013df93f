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

}  // namespace qb200
