// Kernel B: sub-byte bit pack / unpack for sm_100a.
//
// Replaces tpack_cuda_kernel / tunpack_cuda_kernel of the reference
// (engine/kernels/tpack/tpack.cu:30-84, :267-315).  Same byte stream, different machine mapping:
//
//   * a "unit" is 8 consecutive elements = exactly n_bits whole bytes (the same race-free ownership the
//     reference uses, tpack.cu:39-41), but here one LANE owns one unit and one WARP owns one "row" of
//     32 units = 256 elements = 32*n_bits contiguous output bytes;
//   * pack loads are one 256-bit LDG per lane for 4-byte inputs (1 KiB contiguous per warp instruction),
//     128-/64-bit for narrower inputs; the range check of tpack.cu:211-215 (two full reductions + two
//     D2H syncs in the reference) is fused into the same pass as an OR into a device flag;
//   * the unit is composed in registers; for n_bits in {1,2,4,8} a lane's unit is a power-of-two number
//     of bytes and is stored directly (fully coalesced), for n_bits in {3,5,6,7} the warp re-slices its
//     32*n contiguous bytes into 32-bit words with warp shuffles so that stores (pack) and loads
//     (unpack) stay 4-byte coalesced;
//   * every output byte is written exactly once, so no pre-zeroed output and no read-modify-write.
//
// The kernel is HBM-bound: algorithmic bytes/element = sizeof(in) + n/8 (pack), n/8 + 1 (unpack).
#include "common.cuh"

#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace qb200 {
namespace {

constexpr int kRowElems = 256;     // elements per warp-row (32 lanes x 8)
constexpr int kRowsPerWarp = 4;    // rows a warp handles back to back (loads issued first: MLP)
constexpr int kWarpsPerBlock = 8;  // 256 threads

// ---------------------------------------------------------------------------------------------
// element conversion: the reference does `unsigned char element = (char)x[index]` (tpack.cu:50),
// i.e. truncation toward zero for floating inputs, low 8 bits for integer inputs.
// ---------------------------------------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ int to_int(float v) { return __float2int_rz(v); }
    static __device__ __forceinline__ float to_float(float v) { return v; }
};
template <> struct Elem<double> {
    static __device__ __forceinline__ int to_int(double v) { return __double2int_rz(v); }
    static __device__ __forceinline__ float to_float(double v) { return (float)v; }
};
template <> struct Elem<__half> {
    static __device__ __forceinline__ int to_int(__half v) { return __half2int_rz(v); }
    static __device__ __forceinline__ float to_float(__half v) { return __half2float(v); }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ int to_int(__nv_bfloat16 v) { return __float2int_rz(__bfloat162float(v)); }
    static __device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
};
#define QB_INT_ELEM(T)                                                                     \
    template <> struct Elem<T> {                                                           \
        static __device__ __forceinline__ int to_int(T v) { return (int)v; }               \
        static __device__ __forceinline__ float to_float(T v) { return (float)v; }         \
    };
QB_INT_ELEM(int8_t)
QB_INT_ELEM(uint8_t)
QB_INT_ELEM(int16_t)
QB_INT_ELEM(int32_t)
QB_INT_ELEM(int64_t)
#undef QB_INT_ELEM

// ---------------------------------------------------------------------------------------------
// 8 consecutive elements per lane, streaming (read-once) loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg128(const void* p, uint32_t (&r)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "l"(p));
}
__device__ __forceinline__ void ldg64(const void* p, uint32_t (&r)[2]) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "l"(p));
}

template <typename T, int BYTES = sizeof(T)> struct Load8;
template <typename T> struct Load8<T, 1> {
    uint32_t raw[2];
    __device__ __forceinline__ void load(const T* p) { ldg64(p, raw); }
    __device__ __forceinline__ T get(int i) const { return (T)((raw[i >> 2] >> ((i & 3) * 8)) & 0xFF); }
};
template <typename T> struct Load8<T, 2> {
    uint32_t raw[4];
    __device__ __forceinline__ void load(const T* p) { ldg128(p, raw); }
    __device__ __forceinline__ T get(int i) const {
        unsigned short h = (unsigned short)((raw[i >> 1] >> ((i & 1) * 16)) & 0xFFFF);
        T v;
        memcpy(&v, &h, 2);
        return v;
    }
};
template <typename T> struct Load8<T, 4> {
    uint32_t raw[8];
    __device__ __forceinline__ void load(const T* p) { ldg256(p, raw); }
    __device__ __forceinline__ T get(int i) const {
        T v;
        memcpy(&v, &raw[i], 4);
        return v;
    }
};
template <typename T> struct Load8<T, 8> {
    uint32_t raw[16];
    __device__ __forceinline__ void load(const T* p) {
        uint32_t a[8], b[8];
        ldg256(p, a);
        ldg256(reinterpret_cast<const char*>(p) + 32, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) { raw[i] = a[i]; raw[8 + i] = b[i]; }
    }
    __device__ __forceinline__ T get(int i) const {
        T v;
        memcpy(&v, &raw[2 * i], 8);
        return v;
    }
};

// compose the n-bit fields of 8 elements into one unit (<= 64 bits); returns range violation
template <typename T, int NB>
__device__ __forceinline__ bool compose_unit(const Load8<T>& ld, int offset, float lo, float hi, uint64_t& unit) {
    constexpr uint32_t mask = (1u << NB) - 1u;
    bool bad = false;
    uint64_t u = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        T v = ld.get(i);
        float f = Elem<T>::to_float(v);
        bad |= !(f >= lo && f <= hi);  // NaN -> bad, like min()/max() propagating NaN in the reference check
        uint32_t s = ((uint32_t)(Elem<T>::to_int(v) + offset)) & mask;
        u |= (uint64_t)s << (i * NB);
    }
    unit = u;
    return bad;
}

// ---------------------------------------------------------------------------------------------
// pack, vector path: full rows only, x 32-B aligned (16/8 for narrower T), out 8-B aligned.
// ---------------------------------------------------------------------------------------------
template <typename T, int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
tpack_rows_kernel(const T* __restrict__ x, uint8_t* __restrict__ out, int64_t n_rows, int offset,
                  float lo, float hi, int32_t* __restrict__ range_flag) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t row0 = warp * kRowsPerWarp;
    if (row0 >= n_rows) return;

    Load8<T> ld[kRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        if (row0 + r < n_rows) ld[r].load(x + (row0 + r) * kRowElems + lane * 8);
    }

    bool bad = false;
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        if (row0 + r >= n_rows) break;  // warp-uniform
        uint64_t unit;
        bad |= compose_unit<T, NB>(ld[r], offset, lo, hi, unit);
        uint8_t* orow = out + (row0 + r) * (int64_t)(32 * NB);
        if constexpr (NB == 8) {
            reinterpret_cast<uint2*>(orow)[lane] = make_uint2((uint32_t)unit, (uint32_t)(unit >> 32));
        } else if constexpr (NB == 4) {
            reinterpret_cast<uint32_t*>(orow)[lane] = (uint32_t)unit;
        } else if constexpr (NB == 2) {
            reinterpret_cast<uint16_t*>(orow)[lane] = (uint16_t)unit;
        } else if constexpr (NB == 1) {
            orow[lane] = (uint8_t)unit;
        } else {
            // 32*NB contiguous bytes = 8*NB words; word w takes bytes [4w, 4w+4) which live in the
            // units of lane A = 4w/NB (from its byte a = 4w - A*NB on) and, if short, lane A+1.
            const uint32_t ulo = (uint32_t)unit, uhi = (uint32_t)(unit >> 32);
            constexpr int kWords = 8 * NB;
#pragma unroll
            for (int rd = 0; rd < (kWords + 31) / 32; ++rd) {
                const int w = rd * 32 + lane;
                const int A = (4 * w) / NB;
                const int a = 4 * w - A * NB;
                const uint32_t alo = __shfl_sync(0xffffffffu, ulo, A & 31);
                const uint32_t ahi = __shfl_sync(0xffffffffu, uhi, A & 31);
                const uint32_t blo = __shfl_sync(0xffffffffu, ulo, (A + 1) & 31);
                const uint64_t ua = ((uint64_t)ahi << 32) | alo;
                uint32_t word = (uint32_t)(ua >> (8 * a));
                const int have = NB - a;  // bytes available from lane A
                if (have < 4) word |= blo << (8 * have);
                if (w < kWords) reinterpret_cast<uint32_t*>(orow)[w] = word;
            }
        }
    }
    if (bad && range_flag != nullptr) atomicOr(range_flag, 1);
}

// pack, generic path: one thread per unit, scalar accesses, handles the ragged tail / unaligned pointers.
template <typename T>
__global__ void tpack_units_kernel(const T* __restrict__ x, uint8_t* __restrict__ out, int64_t elem0,
                                   int64_t n_elements, int n_bits, int offset, float lo, float hi,
                                   int32_t* __restrict__ range_flag) {
    const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e0 = elem0 + unit * 8;
    if (e0 >= n_elements) return;
    const uint32_t mask = (1u << n_bits) - 1u;
    uint64_t u = 0;
    bool bad = false;
    int valid = 0;
    for (int i = 0; i < 8; ++i) {
        if (e0 + i < n_elements) {
            T v = x[e0 + i];
            float f = Elem<T>::to_float(v);
            bad |= !(f >= lo && f <= hi);
            uint32_t s = ((uint32_t)(Elem<T>::to_int(v) + offset)) & mask;
            u |= (uint64_t)s << (i * n_bits);
            ++valid;
        }
    }
    const int nbytes = (valid * n_bits + 7) / 8;
    uint8_t* o = out + (e0 / 8) * n_bits;
    for (int b = 0; b < nbytes; ++b) o[b] = (uint8_t)(u >> (8 * b));
    if (bad && range_flag != nullptr) atomicOr(range_flag, 1);
}

// ---------------------------------------------------------------------------------------------
// unpack
// ---------------------------------------------------------------------------------------------
template <int NB>
__device__ __forceinline__ uint2 expand_unit(uint64_t unit, uint32_t offset) {
    uint32_t lo, hi;
    const uint32_t off4 = offset * 0x01010101u;
    if constexpr (NB == 8) {
        lo = (uint32_t)unit;
        hi = (uint32_t)(unit >> 32);
    } else if constexpr (NB == 4) {
        const uint32_t u = (uint32_t)unit;
        const uint32_t even = u & 0x0F0F0F0Fu, odd = (u >> 4) & 0x0F0F0F0Fu;
        lo = __byte_perm(even, odd, 0x5140);
        hi = __byte_perm(even, odd, 0x7362);
    } else {
        constexpr uint32_t mask = (1u << NB) - 1u;
        lo = 0;
        hi = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) lo |= ((uint32_t)(unit >> (i * NB)) & mask) << (8 * i);
#pragma unroll
        for (int i = 0; i < 4; ++i) hi |= ((uint32_t)(unit >> ((i + 4) * NB)) & mask) << (8 * i);
    }
    // per-byte (field - offset) with uint8 wrap, as tpack.cu:311 `element -= offset`
    return make_uint2(__vsub4(lo, off4), __vsub4(hi, off4));
}

template <int NB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
tunpack_rows_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out, int64_t n_rows,
                    uint32_t offset) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t row0 = warp * kRowsPerWarp;
    if (row0 >= n_rows) return;

    constexpr int kWords = 8 * NB;
    constexpr int kRounds = (kWords + 31) / 32;
    uint64_t unit[kRowsPerWarp];
    uint32_t wr[kRowsPerWarp][kRounds];

#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        if (row0 + r >= n_rows) break;
        const uint8_t* prow = packed + (row0 + r) * (int64_t)(32 * NB);
        if constexpr (NB == 8) {
            uint2 v = __ldg(reinterpret_cast<const uint2*>(prow) + lane);
            unit[r] = ((uint64_t)v.y << 32) | v.x;
        } else if constexpr (NB == 4) {
            unit[r] = __ldg(reinterpret_cast<const uint32_t*>(prow) + lane);
        } else if constexpr (NB == 2) {
            unit[r] = __ldg(reinterpret_cast<const uint16_t*>(prow) + lane);
        } else if constexpr (NB == 1) {
            unit[r] = __ldg(prow + lane);
        } else {
#pragma unroll
            for (int rd = 0; rd < kRounds; ++rd) {
                const int w = rd * 32 + lane;
                wr[r][rd] = (w < kWords) ? __ldg(reinterpret_cast<const uint32_t*>(prow) + w) : 0u;
            }
        }
    }

#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r) {
        if (row0 + r >= n_rows) break;
        if constexpr (NB != 8 && NB != 4 && NB != 2 && NB != 1) {
            // lane's unit = bytes [lane*NB, lane*NB + NB) of the row: up to 3 consecutive words
            const int b0 = lane * NB;
            const int i0 = b0 >> 2, a = b0 & 3;
            uint32_t v[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int idx = i0 + k;
                uint32_t t = __shfl_sync(0xffffffffu, wr[r][0], idx & 31);
                if constexpr (kRounds > 1) {
                    const uint32_t t1 = __shfl_sync(0xffffffffu, wr[r][1], idx & 31);
                    t = (idx >> 5) ? t1 : t;
                }
                v[k] = t;
            }
            const uint32_t lo = __funnelshift_r(v[0], v[1], 8 * a);
            const uint32_t hi = __funnelshift_r(v[1], v[2], 8 * a);
            unit[r] = ((uint64_t)hi << 32) | lo;
        }
        const uint2 o = expand_unit<NB>(unit[r], offset);
        reinterpret_cast<uint2*>(out + (row0 + r) * (int64_t)kRowElems)[lane] = o;
    }
}

__global__ void tunpack_units_kernel(const uint8_t* __restrict__ packed, uint8_t* __restrict__ out,
                                     int64_t elem0, int64_t n_elements, int n_bits, uint32_t offset,
                                     int64_t n_bytes) {
    const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e0 = elem0 + unit * 8;
    if (e0 >= n_elements) return;
    const uint8_t* p = packed + (e0 / 8) * n_bits;
    const int64_t byte0 = (e0 / 8) * n_bits;
    uint64_t u = 0;
    for (int b = 0; b < n_bits; ++b)
        if (byte0 + b < n_bytes) u |= (uint64_t)p[b] << (8 * b);
    const uint32_t mask = (1u << n_bits) - 1u;
    for (int i = 0; i < 8; ++i)
        if (e0 + i < n_elements) out[e0 + i] = (uint8_t)(((uint32_t)(u >> (i * n_bits)) & mask) - offset);
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
template <typename T, int NB>
int launch_pack_rows(const void* x, uint8_t* out, int64_t n_rows, int offset, float lo, float hi,
                     int32_t* flag, cudaStream_t st) {
    const int64_t warps = ceil_div64(n_rows, kRowsPerWarp);
    const int64_t blocks = ceil_div64(warps, kWarpsPerBlock);
    tpack_rows_kernel<T, NB><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(
        static_cast<const T*>(x), out, n_rows, offset, lo, hi, flag);
    QB_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int pack_typed(const void* x, int64_t n, int nb, int sign, uint8_t* out, int32_t* flag, cudaStream_t st) {
    const int offset = sign ? (1 << (nb - 1)) : 0;
    const float lo = sign ? -(float)(1 << (nb - 1)) : 0.f;
    const float hi = sign ? (float)((1 << (nb - 1)) - 1) : (float)((1 << nb) - 1);
    const size_t in_align = sizeof(T) * 8 >= 32 ? 32 : sizeof(T) * 8;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) % in_align == 0) &&
                         (reinterpret_cast<uintptr_t>(out) % 8 == 0);
    int64_t n_rows = aligned ? n / kRowElems : 0;
    // grid.x limit: 2^31-1 blocks of 32 rows -> far beyond any tensor that fits in HBM
    if (n_rows > 0) {
        int rc = 0;
        switch (nb) {
#define QB_CASE(NB) case NB: rc = launch_pack_rows<T, NB>(x, out, n_rows, offset, lo, hi, flag, st); break;
            QB_CASE(1) QB_CASE(2) QB_CASE(3) QB_CASE(4) QB_CASE(5) QB_CASE(6) QB_CASE(7) QB_CASE(8)
#undef QB_CASE
        }
        if (rc) return rc;
    }
    const int64_t e0 = n_rows * kRowElems;
    if (e0 < n) {
        const int64_t units = ceil_div64(n - e0, 8);
        const int threads = 128;
        tpack_units_kernel<T><<<(unsigned)ceil_div64(units, threads), threads, 0, st>>>(
            static_cast<const T*>(x), out, e0, n, nb, offset, lo, hi, flag);
        QB_LAUNCH_CHECK();
    }
    return 0;
}

template <int NB>
int launch_unpack_rows(const uint8_t* packed, uint8_t* out, int64_t n_rows, uint32_t offset, cudaStream_t st) {
    const int64_t warps = ceil_div64(n_rows, kRowsPerWarp);
    const int64_t blocks = ceil_div64(warps, kWarpsPerBlock);
    tunpack_rows_kernel<NB><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, st>>>(packed, out, n_rows, offset);
    QB_LAUNCH_CHECK();
    return 0;
}

}  // namespace
}  // namespace qb200

extern "C" {

int64_t qb200_packed_bytes(int64_t n_elements, int n_bits) {
    if (n_elements < 0 || n_bits <= 0 || n_bits > 8) return -1;
    return (n_elements * n_bits + 7) / 8;  // tpack.cu:224
}

int qb200_tpack(const void* x, int x_dtype, int64_t n_elements, int n_bits, int sign, uint8_t* out,
                int32_t* range_flag, void* stream) {
    using namespace qb200;
    QB_REQUIRE(n_bits > 0 && n_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    QB_REQUIRE(n_elements >= 0, QB200_EINVAL, "n_elements must be non-negative");
    if (n_elements == 0) return 0;
    QB_REQUIRE(x != nullptr && out != nullptr, QB200_EINVAL, "tpack: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int sg = sign ? 1 : 0;
    switch (x_dtype) {
        case QB200_U8: return pack_typed<uint8_t>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_I8: return pack_typed<int8_t>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_I16: return pack_typed<int16_t>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_I32: return pack_typed<int32_t>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_I64: return pack_typed<int64_t>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_F16: return pack_typed<__half>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_BF16: return pack_typed<__nv_bfloat16>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_F32: return pack_typed<float>(x, n_elements, n_bits, sg, out, range_flag, st);
        case QB200_F64: return pack_typed<double>(x, n_elements, n_bits, sg, out, range_flag, st);
        default: break;
    }
    set_error("tpack: unsupported dtype code %d", x_dtype);
    return QB200_EINVAL;
}

int qb200_tunpack(const uint8_t* packed, int64_t n_elements, int n_bits, int sign, void* out, void* stream) {
    using namespace qb200;
    QB_REQUIRE(n_bits > 0 && n_bits <= 8, QB200_EINVAL, "n_bits must be in the range (0, 8]");
    QB_REQUIRE(n_elements >= 0, QB200_EINVAL, "n_elements must be non-negative");
    if (n_elements == 0) return 0;
    QB_REQUIRE(packed != nullptr && out != nullptr, QB200_EINVAL, "tunpack: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t offset = sign ? (1u << (n_bits - 1)) : 0u;
    uint8_t* o = static_cast<uint8_t*>(out);
    const bool aligned = (reinterpret_cast<uintptr_t>(packed) % 8 == 0) && (reinterpret_cast<uintptr_t>(o) % 8 == 0);
    const int64_t n_rows = aligned ? n_elements / kRowElems : 0;
    if (n_rows > 0) {
        int rc = 0;
        switch (n_bits) {
#define QB_CASE(NB) case NB: rc = launch_unpack_rows<NB>(packed, o, n_rows, offset, st); break;
            QB_CASE(1) QB_CASE(2) QB_CASE(3) QB_CASE(4) QB_CASE(5) QB_CASE(6) QB_CASE(7) QB_CASE(8)
#undef QB_CASE
        }
        if (rc) return rc;
    }
    const int64_t e0 = n_rows * kRowElems;
    if (e0 < n_elements) {
        const int64_t units = ceil_div64(n_elements - e0, 8);
        const int threads = 128;
        tunpack_units_kernel<<<(unsigned)ceil_div64(units, threads), threads, 0, st>>>(
            packed, o, e0, n_elements, n_bits, offset, qb200_packed_bytes(n_elements, n_bits));
        QB_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
