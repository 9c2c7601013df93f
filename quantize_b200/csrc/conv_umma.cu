// Kernel A: int8 implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a only).
//
// Replaces the per-output-element scalar loop of the reference (quantconv2d_float_input.cu:45-121) for
// groups == 1.  GEMM view:  D[M = N*P*Q pixels][N = K out-channels] = A[M][Kg = R*S*Cp] * B[Kg][N]
//   A  = im2col of the quantized NHWC(Cp) u8 activations — never materialised: the TMA unit gathers each
//        128-pixel x KC-channel slice of one filter tap straight into swizzled shared memory
//        (cp.async.bulk.tensor.4d ... im2col; out-of-image taps are zero-filled by the hardware, which is
//        exactly the reference's bounds check at :92);
//   B  = prepared weights [K][R][S][Cp] (K-major), fetched with a tiled 2-D TMA;
//   D  = int32 accumulators in TMEM (tcgen05.mma.cta_group::1.kind::i8, u8 x s8 -> s32), double buffered so the
//        epilogue of tile i overlaps the main loop of tile i+1.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..9 = epilogue (tcgen05.ld -> dequant -> coalesced NCHW stores: a TMEM lane is an output pixel, so the 32
// lanes of a warp write 32 consecutive pixels of one output channel = one 128-byte line).  The per-channel
// constants (s_a*s_w[k], bias[k], sum(qw[k])) are staged in shared memory once per tile so that an
// interior pixel costs one convert + one or two FMAs + one store per output element (every path rounds identically); border pixels of a layer with a
// non-zero activation zero point take the 4-corner prefix-sum path.
// Persistent: grid = #SMs, static round-robin tile schedule, output-channel tiles fastest so concurrently
// running CTAs share the same activation slice in L2.
//
// Variants of the same kernel: halo (stride-1 R x S kernels with 64 padded channels: one TMA halo
// load per tile, taps through row-shifted descriptors, tap-major weights), im2col rows (few-channel stems: the A
// operand is a materialised row matrix), fused quantize (kFQ: fp32 tiles by TMA, eight quantizer warps), residual tail
// (kRes: identity tensor streamed through per-warp cp.async rings), quantized hand-off (kQ8: the epilogue writes the
// consumer layer's int8 workspace; kGroups = 2 epilogue groups when that is the only output), ragged channel counts
// (kRagged).  Each is a template parameter because the epilogue's register budget decides its speed.
#include "common.cuh"
#include "conv_common.cuh"
#include "quant_math.cuh"

#include <cuda.h>
#include <cstdint>
#include <climits>

#include <atomic>
#include <type_traits>

namespace qb200 {
namespace {

#ifndef QB200_FQ_KPRE
#define QB200_FQ_KPRE 2   // column groups whose constants the fused-quantize kernel's epilogue fetches ahead (see `fast`)
#endif
constexpr int kBM = 128;            // pixels per tile = TMEM lanes = UMMA M
constexpr int kEpiWarps = 8;        // two warps per TMEM lane quadrant, each owns half of the tile's columns
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kFqWarps = 8;         // fused-quantize variant: quantizer warps, each owns 8 channels x 128 pixels of a k-block
constexpr int kThreadsFq = kThreads + 32 + kFqWarps * 32;  // + one TMA warp for the fp32 tiles
constexpr int kMaxXStages = 6;      // fp32 staging ring of the fused-quantize variant (3..6 slots, what fits) / halo ring (<= 3)
constexpr int kFqKC = 64;           // its k-block: 64 channels (a 32 KB fp32 tile, an 8 KB u8 A tile)
constexpr int kConstFloats = 3 * 256;  // per-tile channel constants: scale, interior bias, raw bias
constexpr int kMaxBoxes = 16;        // fused stem: ring of fp32 row boxes (own barriers behind the common ones)
constexpr int kBarBytes = 576 + 2 * kMaxBoxes * 8;       // mbarriers + TMEM slot (+ the box barriers)
constexpr int kTailBytes = kBarBytes + 2 * kConstFloats * 4;  // barriers + two constant buffers
constexpr int kMaxWclsBytes = 16 * 1024;                 // one buffer of per-window-class channel sums (layers with R*S > 1)
constexpr int kMaxCls = 16;                              // distinct row (and column) windows supported by the class table
constexpr int kTmemCols = 512;
constexpr int kMaxAcc = 4;          // accumulator buffers in TMEM: 2 x 256 columns, or 4 x 128 when the tile has <= 128 channels
constexpr int kMaxStages = 24;      // small stages (weight-only tiles of the halo variant) need depth to cover TMA latency
constexpr size_t kSmemBudget = 200 * 1024;
constexpr int kResDepth = 3;                            // residual chunks in flight per epilogue warp (cp.async ring)
constexpr int kResChunkFloats = 32 * 32;                // one chunk: 32 channels x 32 pixels
constexpr size_t kResBytes = (size_t)kEpiWarps * kResDepth * kResChunkFloats * 4;  // 96 KB
constexpr int kAStatBar = 12;        // A-stationary mode: barriers of the A ring = full / empty[kAStatBar ..] (weight stages < kAStatBar)
constexpr size_t kSmemBudgetFq = 224 * 1024;  // the fused-quantize variant wants every byte for fp32 tiles in flight

// Division of a non-negative 32-bit value by a launch-time constant as one multiply-high + shift: the per-tile index
// arithmetic of the epilogue (tile -> image, row, column) otherwise costs ~40 instructions per division.
struct FastDiv {
    uint32_t mul, shift, d;
    __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shift); }
};
static FastDiv make_fastdiv(int d) {
    FastDiv f;
    f.d = (uint32_t)(d < 1 ? 1 : d);
    f.mul = 0;
    f.shift = 0;
    if (f.d > 1) {
        int lg = 0;
        while ((1u << lg) < f.d) ++lg;                        // ceil(log2 d)
        const int p = 31 + lg;                                // exact for 0 <= n < 2^31
        f.mul = (uint32_t)((((unsigned __int128)1 << p) + f.d - 1) / f.d);
        f.shift = (uint32_t)(p - 32);
    }
    return f;
}

struct UmmaParams {
    ConvGeom g;        // the convolution (epilogue: output addressing, border windows)
    ConvGeom gm;       // what the main loop iterates: g itself, or a 1x1 "conv" over materialised im2col rows
    EpilogueParams ep;
    int KC;            // channel bytes per k-block: 32, 64 or 128 (== swizzle span)
    int BN;            // out-channel tile: 64, 128 or 256
    int stages;
    int cblocks;       // Cp / KC
    int m_tiles, n_tiles;
    int64_t M;         // N*P*Q
    uint32_t idesc;
    uint32_t sbo16;    // stride-byte-offset >> 4 (8 rows * KC bytes)
    uint32_t layout;   // UMMA smem layout type
    // Zero-point term of border pixels: the in-bounds tap window of an output pixel is one of n_rcls x n_ccls
    // classes (rows x columns); the epilogue keeps sum(qw over the window) per (class, channel) in shared memory.
    int wcls_smem;     // bytes of one class-table buffer (0: 4-corner lookups in the global prefix tables instead)
    int n_rcls, n_ccls;
    uint8_t rcls[kMaxCls][2], ccls[kMaxCls][2];   // [lo, hi) tap ranges
    // closed form of the class index of output row p: min(p, cls_head_r) + max(0, p - cls_tail_r + 1) (columns alike);
    // cls_fast = the host checked it against the enumeration
    int cls_fast, cls_head_r, cls_tail_r, cls_head_c, cls_tail_c;
    int* err_flag;     // device int: set non-zero by the watchdog
    // fused-quantize variant (1x1, stride 1): A is produced from the fp32 NCHW input inside the kernel
    int tiles_per_img; // > 0: M tiles never straddle images (tile = image, 128-pixel block); 0: flat pixel tiling
    // halo variant (stride 1, R*S > 1): activations are zero-padded NHWC [N][Hp][Wp][Cp]; an M tile is 128 consecutive
    // positions of the padded-flat output grid and ONE 2-D TMA load of halo_rows rows serves every filter tap — tap
    // (r,s) is the same shared-memory region read through a descriptor shifted by (r*Wp + s) rows
    // (tests/native/desc_shift.cu shows the hardware swizzles on absolute addresses, so shifted descriptors are exact).
    int halo;          // 0/1
    int Hp, Wp, halo_rows, halo_bytes, h_stages;
    int x_stages;  // fused-quantize variant: slots of the fp32 ring
    // A-stationary fused quantize (1x1 layers with several channel tiles): a CTA owns whole pixel tiles and runs ALL their
    // channel tiles back to back; the quantized A k-blocks live in their own ring of a_slots slots (behind the fp32 ring),
    // are written once per pixel tile and released by the MMAs of the LAST channel tile; stages then hold weights only
    int a_stat, a_slots;
    // fused quantize, small planes: the quantizer threads fetch their own fp32 values with cp.async (thread-private ring
    // slots, no TMA boxes, no barriers on the input side); implies flat pixel tiling
    int x_cpasync;
    int n_acc, acc_stride;  // TMEM accumulator buffers and the columns between them
    int res_async;          // residual tail: stream the identity tensor through a per-warp cp.async ring
    FastDiv fd_ntiles, fd_tpi, fd_pq, fd_q, fd_wp;   // n_tiles, tiles_per_img, P*Q, Q, Wp
    int tap_group;     // filter taps per weight stage (1, S or R*S): one 3-D TMA box of the tap-major weight copy
    int a_tiled;       // 1x1 / stride 1 / pad 0 main loop: A is the plain matrix [N*H*W][Cp], fetched with TILED 2-D TMA boxes
    // fused stem (kStem): few-channel layer whose grouped im2col rows (conv_common.cuh: im2col8_row_bytes) are built in
    // shared memory from the fp32 input — the rows never exist in HBM.  Per tile ONE 4-D TMA box [C][st_rows_box][W] of
    // fp32 input rows (rows outside the image zero-filled) -> quantized byte planes [C][st_rows_box][st_Wq] (input column
    // iw at byte st_margin + iw; margins stay 0 = the reference's zero padding) -> 8-byte windows of the planes.
    int st_rows_box, st_Wq, st_margin, st_box_bytes, st_box_pitch, st_plane_bytes, st_ring, st_dbg;   // (pitch: boxes 128-byte aligned in the slot)
    const float* x;
    const float* q_scale;
    const float* q_zero;
    const float* q_qmin;
    const float* q_qmax;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Shared-memory loads through the shared window.  The epilogue's constant tables live in dynamic shared memory behind
// pointers the compiler cannot classify: a plain dereference compiles to a GENERIC load (LD.E), which (a) has several
// times the latency of LDS and (b) may alias the global stores of the epilogue, so it is never hoisted above them — the
// ncu source page of round 2 showed every 4-column group of every epilogue stalled on two such loads (50 % of all
// samples of the conv kernel on `long_sb`).  These wrappers are volatile without a memory clobber: ordered against the
// barrier / TMEM instructions, free to move across ordinary stores.
__device__ __forceinline__ float4 lds4(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ float lds1(const float* p) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void sts2(void* p, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(smem_u32(p)), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, uint32_t src_bytes) {   // bytes past src_bytes are zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires, so a waiting
// warp does not burn issue slots that the quantizer / epilogue warps of the same SM sub-partition need.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// Polling wait with a plain timed sleep between polls (no barrier-armed suspend): for warps that wait for a long time
// next to instruction-bound warps.  Measured: after a try_wait suspend has been cut short once, the re-armed suspends of
// the same wait return at once and the loop spins (8 M iterations on one layer), stealing issue slots.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
    long long t0 = 0;
    uint32_t spins = 0;
    while (!mbar_test_wait(bar, parity)) {
        __nanosleep(kSleepNs);
        if ((++spins & 0xFFu) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 8000000000ll) {
                if (err_flag) atomicExch(err_flag, code);
                __threadfence_system();
                __trap();
            }
        }
    }
}
// Bounded wait: a lost arrival must not hang the GPU — after ~4 s the kernel flags the error and traps.
// kSleepNs > 0: back off between polls (latency-insensitive waiters that share an SM with instruction-bound warps).
template <int kSleepNs = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err_flag, int code) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (kSleepNs > 0) __nanosleep(kSleepNs);
        if ((++spins & 0xFFu) == 0) {  // look at the clock only every 256 wake-ups
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > 8000000000ll) {
                if (err_flag) atomicExch(err_flag, code);
                __threadfence_system();
                __trap();
            }
        }
    }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"(off_w), "h"(off_h)
        : "memory");
}
// ---- CTA-pair (cta_group::2) forms: the completion is signalled on a barrier given by its shared::cluster address (the
//      leader CTA's), proven on hardware by tests/native/umma_pair.cu ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(x), "r"(y)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster, int c, int w, int h,
                                                        int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c), "r"(w), "r"(h), "r"(n),
          "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z, int w) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z), "r"(w)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], u8 x s8 -> s32
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_i8_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi)
        : "memory");
}
// CTA pair: M = 256 (128 rows per CTA), each CTA holds half of the B tile; issued by the leader for both
__device__ __forceinline__ void umma_i8_lohi_pair(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(desc_hi)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
template <int kThreadsInGroup>
__device__ __forceinline__ void epi_barrier(int group) { asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kThreadsInGroup) : "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (tcgen05): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout <<61
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo16, uint32_t layout) {
    uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major, canonical value 1
    d |= (uint64_t)(sbo16 & 0x3FFF) << 32;  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// kRes: the fused residual tail is compiled in.  It is a template parameter because its mere presence (ring state, 32 more
// live registers per chunk) cost the plain kernel 6 % on output-heavy layers.
// kQ8: the quantized hand-off (int8-out epilogue) is compiled in — same reason.
// kGroups = 2: two groups of eight epilogue warps work on alternate tiles (int8-only hand-off layers: their epilogue is
// instruction- and latency-bound — ~2 us per 128 x 64 tile with 2 warps per scheduler — so a second group doubles it).
// kRagged: K is not a multiple of the channel tile; raggedness is then resolved per 32-column chunk (chunks past K are
// skipped, only a partially valid chunk takes the per-element path).  A separate instantiation because the per-chunk
// checks cost the common kernels 10 % on output-heavy layers (register pressure in the epilogue).
// kPair: CTA-pair variant for deep reductions (tcgen05.mma.cta_group::2).  Two CTAs of a cluster compute a 256-pixel x BN
// tile: each stages its own 128 im2col rows and HALF of the weight tile (BN / 2 rows); the pair's tensor cores read both
// halves, so the operand bytes per MAC drop by a third (128 + 128 instead of 128 + 256 rows per CTA and k-block) — deep
// 3x3 layers are bound by exactly that L2 -> SM feed (ncu round 2: 9.3 TB/s of operand traffic at 37 % tensor-pipe
// activity).  Protocol as in tests/native/umma_pair.cu: both producers signal the LEADER's full barrier (cluster-scope TMA
// completion, the leader arms it with the bytes of both CTAs); the leader issues every MMA and its commits arrive, multicast,
// on the stage-empty / accumulator-full barriers of both CTAs; every epilogue warp of the pair arrives on the leader's
// accumulator-empty barrier.  Supported for the plain im2col main loop (not the halo / fused-quantize variants).
// kStem (with kFQ): the quantizer warps build the grouped im2col rows of a few-channel layer (the 7x7 RGB stem) in shared
// memory from fp32 input rows (UmmaParams::st_*); main loop and epilogue are those of the im2col-rows layer.
template <bool kFQ, bool kRes, bool kQ8, int kGroups = 1, bool kRagged = false, bool kPair = false, bool kStem = false>
__global__ void __launch_bounds__(kFQ ? kThreadsFq : 64 + kGroups * kEpiWarps * 32, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const UmmaParams prm, void* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // (stage ring base)

    const ConvGeom& g = prm.g;
    const ConvGeom& gm = prm.gm;
    const int KC = prm.KC, BN = prm.BN, stages = prm.stages;
    static_assert(!kPair || (!kFQ && kGroups == 1), "the CTA-pair variant covers the plain main loop");
    static_assert(!kStem || kFQ, "the fused stem is a producer mode of the fused-quantize kernel");
    const bool halo = !kFQ && !kPair && prm.halo != 0;
    const uint32_t pair_rank = kPair ? cluster_ctarank() : 0u;      // 0 = leader (issues the MMAs)
    const bool a_stat = kFQ && !kStem && prm.a_stat != 0;
    const uint32_t a_bytes = halo ? 0u : (uint32_t)(kBM * KC), b_bytes = (kPair ? BN / 2 : BN) * KC,
                   stage_bytes = kStem ? a_bytes * (uint32_t)prm.cblocks
                                       : (halo ? b_bytes * (uint32_t)prm.tap_group : (a_stat ? b_bytes : a_bytes + b_bytes));
    // fused stem: the whole weight matrix (cblocks tiles) stays resident in front of the ring; a stage is a whole A tile
    // (all k-blocks: one barrier round trip per tile)
    uint8_t* const bres = smem;
    if constexpr (kStem) smem += (size_t)prm.cblocks * b_bytes;
    uint8_t* xring = smem + (size_t)stages * stage_bytes;                       // fused-quantize: [x_stages][KC][128] fp32
    const uint32_t x_bytes = kStem ? (uint32_t)prm.st_box_pitch
                                   : (kFQ ? (uint32_t)KC * kBM * 4u : (halo ? (uint32_t)prm.halo_bytes : 0u));  // ring slot size
    const int x_slots = kFQ ? prm.x_stages : (halo ? prm.h_stages : 0);       // fp32 tiles (kFQ) or u8 halo tiles (halo)
    uint8_t* const aring = xring + (size_t)x_slots * x_bytes;                  // A-stationary: [a_slots][128][KC] u8
    uint64_t* bars = reinterpret_cast<uint64_t*>(aring + (a_stat ? (size_t)prm.a_slots * a_bytes : 0));
    uint64_t* full = bars;                     // [stages]
    uint64_t* empty = bars + kMaxStages;       // [stages]
    uint64_t* acc_full = bars + 2 * kMaxStages;               // [kMaxAcc]
    uint64_t* acc_empty = bars + 2 * kMaxStages + kMaxAcc;    // [kMaxAcc]
    uint64_t* xfull = bars + 2 * kMaxStages + 2 * kMaxAcc;  // [kMaxXStages]
    uint64_t* xempty = xfull + kMaxXStages;          // [kMaxXStages]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xempty + kMaxXStages);
    uint64_t* bfull = bars + 72;               // [kMaxBoxes] fused stem: one fp32 row box each (byte offset 576)
    uint64_t* bempty = bfull + kMaxBoxes;      // [kMaxBoxes]
    float* consts = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kBarBytes);  // [2][3][256]
    float* wcls_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kTailBytes);  // [2][n_cls][BN]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // work units: tiles, round-robin over the CTAs — or, for the pair variant, 256-pixel pair tiles over the CTA pairs
    // (fused stem: every CTA takes a CONTIGUOUS run of tiles, so that consecutive tiles share quantized input rows)
    // (A-stationary: work items are CTA-local — item j = channel tile j % n_tiles of the CTA's (j / n_tiles)-th pixel tile,
    // pixel tiles round-robin over the CTAs: m_tile = blockIdx.x + (j / n_tiles) * gridDim.x)
    const int all_tiles = kPair ? ((prm.m_tiles + 1) >> 1) * prm.n_tiles : prm.m_tiles * prm.n_tiles;
    const int stem_run = kStem ? (all_tiles + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int my_m = a_stat ? (prm.m_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_tiles = kStem ? min(all_tiles, ((int)blockIdx.x + 1) * stem_run) : (a_stat ? my_m * prm.n_tiles : all_tiles);
    const int unit0 = kStem ? (int)blockIdx.x * stem_run : (a_stat ? 0 : (kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x));
    const int unit_step = (kStem || a_stat) ? 1 : (kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x);
    const int kblocks = gm.R * gm.S * prm.cblocks;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int i = 0; i < stages; ++i) {
            mbar_init(&full[i], kStem ? kFqWarps : ((kFQ && !a_stat) ? 1 + kFqWarps : 1));  // TMA expect_tx arrival (+ one arrival per quantizer warp; stem: the warps only)
            mbar_init(&empty[i], 1);
        }
        if (a_stat) {
            for (int i = 0; i < prm.a_slots; ++i) {
                mbar_init(&full[kAStatBar + i], kFqWarps);    // one arrival per quantizer warp
                mbar_init(&empty[kAStatBar + i], 1);          // the commit after the last channel tile's MMAs
            }
        }
        for (int i = 0; i < kMaxAcc; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], kPair ? 2 * kEpiWarps : (kStem ? 4 : kEpiWarps));   // pair: the epilogue warps of both CTAs; stem: groups of four warps
        }
        for (int i = 0; i < kMaxXStages; ++i) {
            mbar_init(&xfull[i], 1);
            mbar_init(&xempty[i], kFQ ? kFqWarps : 1);
        }
        if constexpr (kStem) {
            for (int i = 0; i < kMaxBoxes; ++i) {
                mbar_init(&bfull[i], 1);
                mbar_init(&bempty[i], kFqWarps);
            }
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (kPair) tmem_alloc_pair(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();   // the peer's barriers must be initialised before anything can signal them
    else __syncthreads();
    // everything above touched only shared memory, TMEM and kernel parameters: it may overlap the previous kernel's tail
    pdl_launch_dependents();
    pdl_wait();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int PQ = gm.P * gm.Q;
            int hs = 0;
            uint32_t hphase = 0;
            if (halo) {
                // Work items j = (tile, channel block).  The halo of item j+1 is requested BEFORE the weight tiles of
                // item j: the weight ring is shallower than R*S taps, so the producer would otherwise only reach the
                // next halo after this item's MMAs have started, exposing a full load latency per tile.
                int h_tile = blockIdx.x, h_cb = 0;   // next halo to request
                auto request_halo = [&]() {
                    if (h_tile >= total_tiles) return;
                    const int m_tile = h_tile / prm.n_tiles;
                    const int img = m_tile / prm.tiles_per_img, t = m_tile - img * prm.tiles_per_img;
                    const int f0 = img * prm.Hp * prm.Wp + t * kBM;   // first padded-flat position of the tile
                    mbar_wait(&xempty[hs], hphase ^ 1, prm.err_flag, 6);
                    mbar_expect_tx(&xfull[hs], (uint32_t)(prm.halo_rows * KC));
                    tma_load_2d(xring + (size_t)hs * x_bytes, &tmap_a, &xfull[hs], h_cb * KC, f0);
                    if (++hs == prm.h_stages) { hs = 0; hphase ^= 1; }
                    if (++h_cb == prm.cblocks) { h_cb = 0; h_tile += gridDim.x; }
                };
                request_halo();
                for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                    const int n_tile = tile % prm.n_tiles;
                    for (int cb = 0; cb < prm.cblocks; ++cb) {
                        request_halo();
                        const int taps = gm.R * gm.S;
                        for (int tap = 0; tap < taps; tap += prm.tap_group) {   // weight tiles of tap_group taps: one box
                            mbar_wait(&empty[stage], phase ^ 1, prm.err_flag, 1);
                            mbar_expect_tx(&full[stage], stage_bytes);
                            tma_load_3d(smem + (size_t)stage * stage_bytes, &tmap_b, &full[stage], 0, n_tile * BN, cb * taps + tap);
                            if (++stage == stages) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            if constexpr (kStem) {   // the weights: once per CTA, all k-blocks, one barrier (the spare slot of the x ring's)
                uint64_t* wfull = &xfull[kMaxXStages - 1];
                mbar_expect_tx(wfull, (uint32_t)prm.cblocks * b_bytes);
                for (int cb = 0; cb < prm.cblocks; ++cb) tma_load_2d(bres + (size_t)cb * b_bytes, &tmap_b, wfull, cb * KC, 0);
            }
            for (int tile = unit0; !kStem && !halo && tile < total_tiles; tile += unit_step) {
                const int n_tile = tile % prm.n_tiles;
                // pair: this CTA loads the rows of m-tile 2 * (pair tile) + rank; an odd tail re-loads the last tile
                const int m_tile = kPair ? min(2 * (tile / prm.n_tiles) + (int)pair_rank, prm.m_tiles - 1) : tile / prm.n_tiles;
                const int64_t m0 = (int64_t)m_tile * kBM;
                const int img = (int)(m0 / PQ);
                const int rem = (int)(m0 - (int64_t)img * PQ);
                const int p0 = rem / gm.Q, q0 = rem - p0 * gm.Q;
                const int cw = q0 * gm.stride - gm.pad, ch = p0 * gm.stride - gm.pad;
                for (int r = 0; r < gm.R; ++r)
                    for (int s = 0; s < gm.S; ++s)
                        for (int cb = 0; cb < prm.cblocks; ++cb) {
                            mbar_wait<(kFQ && !kStem) ? 200 : 0>(&empty[stage], phase ^ 1, prm.err_flag, 1);
                            uint8_t* sa = smem + (size_t)stage * stage_bytes;
                            uint8_t* sb = sa + (a_stat ? 0u : a_bytes);
                            if constexpr (kPair) {
                                // both CTAs' loads complete on the LEADER's barrier, armed with the bytes of both
                                const uint32_t lead_bar = mapa_u32(smem_u32(&full[stage]), 0);
                                if (pair_rank == 0) mbar_expect_tx(&full[stage], 2 * stage_bytes);
                                if (prm.a_tiled) tma_load_2d_pair(sa, &tmap_a, lead_bar, cb * KC, (int)m0);
                                else tma_load_im2col_4d_pair(sa, &tmap_a, lead_bar, cb * KC, cw, ch, img, (uint16_t)s, (uint16_t)r);
                                tma_load_2d_pair(sb, &tmap_b, lead_bar, (r * gm.S + s) * gm.Cp + cb * KC,
                                                 n_tile * BN + (int)pair_rank * (BN / 2));
                                if (++stage == stages) { stage = 0; phase ^= 1; }
                                continue;
                            }
                            if (kFQ) {
                                mbar_expect_tx(&full[stage], b_bytes);
                            } else {
                                mbar_expect_tx(&full[stage], stage_bytes);
                                if (prm.a_tiled) tma_load_2d(sa, &tmap_a, &full[stage], cb * KC, (int)m0);
                                else tma_load_im2col_4d(sa, &tmap_a, &full[stage], cb * KC, cw, ch, img, (uint16_t)s, (uint16_t)r);
                            }
                            tma_load_2d(sb, &tmap_b, &full[stage], (r * gm.S + s) * gm.Cp + cb * KC, n_tile * BN);
                            if (++stage == stages) { stage = 0; phase ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The issuing thread is latency-bound (one dependent scalar instruction every ~5 cycles), and layers with many
        // small k-blocks (3x3, C = 64) spent ~650 cycles per k-block here — more than the MMAs themselves (ncu,
        // profiles/README.md).  So the loop is kept minimal: the whole warp runs it (uniform control flow), descriptors
        // are 32-bit low words advanced by adds (the high word is loop-invariant), and one lane issues.
        const uint32_t desc_hi = (prm.sbo16 & 0x3FFFu) | (1u << 14) | ((prm.layout & 7u) << 29);  // SBO, version 1, layout
        const uint32_t lo_flag = 1u << 16;                                                        // LBO field = 1 (unused)
        const uint32_t idesc = prm.idesc;
        const uint32_t base16 = smem_u32(smem) >> 4;            // 16-byte units
        const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4;
        const uint32_t xring16 = smem_u32(xring) >> 4, x16 = x_bytes >> 4;
        const uint32_t wp16 = (uint32_t)(prm.Wp * KC) >> 4, kc16 = (uint32_t)KC >> 4;
        const int n_mma = KC / 32;
        const int R = gm.R, S = gm.S, cblocks = prm.cblocks, h_stages = prm.h_stages;
        int stage = 0, buf = 0, hs = 0;
        uint32_t phase = 0, acc_phase = 0, hphase = 0;
        uint32_t stage_lo = base16;
        if constexpr (kStem) mbar_wait(&xfull[kMaxXStages - 1], 0, prm.err_flag, 3);   // resident weights have landed
        const uint32_t bres16 = smem_u32(bres) >> 4;
        const uint32_t aring16 = smem_u32(aring) >> 4;
        int as_base = 0, nn = 0;          // A-stationary: ring slot of the pixel tile's first k-block, channel tile of this item
        uint32_t as_phase = 0;
        for (int tile = unit0; tile < total_tiles && pair_rank == 0; tile += unit_step) {   // (pair: the leader issues for both)
            mbar_wait<(kFQ && !kStem) ? 200 : 0>(&acc_empty[buf], acc_phase ^ 1, prm.err_flag, 2);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * prm.acc_stride);
            uint32_t accumulate = 0;
            if constexpr (kStem) {
                // one stage = the tile's whole A operand; weights resident
                mbar_wait(&full[stage], phase, prm.err_flag, 3);
                tc_fence_after();
                if (lane == 0) {
                    for (int kb = 0; kb < cblocks; ++kb) {
                        const uint32_t a0 = (stage_lo + (uint32_t)kb * a16) | lo_flag, b0 = (bres16 + (uint32_t)kb * (b_bytes >> 4)) | lo_flag;
                        umma_i8_lohi(tmem_d, a0, b0, desc_hi, idesc, kb ? 1u : 0u);
                        umma_i8_lohi(tmem_d, a0 + 2, b0 + 2, desc_hi, idesc, 1u);
                    }
                    umma_commit(&empty[stage]);
                }
                stage_lo += stage16;
                if (++stage == stages) { stage = 0; stage_lo = base16; phase ^= 1; }
            } else if (halo) {
                for (int cb = 0; cb < cblocks; ++cb) {
                    mbar_wait(&xfull[hs], hphase, prm.err_flag, 7);
                    const uint32_t halo_lo = xring16 + (uint32_t)hs * x16;
                    const uint32_t b16 = b_bytes >> 4;
                    if (R == 3 && S == 3 && (n_mma == 2 || n_mma == 4) && (prm.tap_group == 9 || prm.tap_group == 3)) {
                        // 3x3 kernels: the MMAs of one filter row (3 taps x n_mma k-slices) with compile-time offsets.
                        // The generic loop below spent ~490 cycles per tap on loop bookkeeping and on moving operands to
                        // uniform registers (ncu source page: 75 instructions per tap around 2 MMAs), and the issuing
                        // warp — not the epilogue or memory — bounded these layers (64->64 3x3 @56x56: 104 -> 73 us).
                        auto mma_row = [&](auto nm_tag, uint32_t a_row, uint32_t b_row, uint32_t first_acc) {
                            constexpr int NM = decltype(nm_tag)::value;
#pragma unroll
                            for (int ss = 0; ss < 3; ++ss) {
#pragma unroll
                                for (int k = 0; k < NM; ++k)
                                    umma_i8_lohi(tmem_d, a_row + (uint32_t)ss * kc16 + 2u * k, b_row + (uint32_t)ss * b16 + 2u * k,
                                                 desc_hi, idesc, (ss | k) ? 1u : first_acc);
                            }
                        };
                        const uint32_t a_row0 = halo_lo | lo_flag;
                        const bool per_row = prm.tap_group == 3;   // one weight stage per filter row, else one per tile
                        for (int rr = 0; rr < 3; ++rr) {
                            if (rr == 0 || per_row) {
                                mbar_wait(&full[stage], phase, prm.err_flag, 3);
                                tc_fence_after();
                            }
                            if (lane == 0) {
                                const uint32_t a_row = a_row0 + (uint32_t)rr * wp16;
                                const uint32_t b_row = (stage_lo | lo_flag) + (per_row ? 0u : (uint32_t)(rr * 3) * b16);
                                if (n_mma == 2) mma_row(std::integral_constant<int, 2>{}, a_row, b_row, accumulate);
                                else mma_row(std::integral_constant<int, 4>{}, a_row, b_row, accumulate);
                                if (per_row || rr == 2) umma_commit(&empty[stage]);
                            }
                            accumulate = 1;
                            if (per_row || rr == 2) {
                                stage_lo += stage16;
                                if (++stage == stages) { stage = 0; stage_lo = base16; phase ^= 1; }
                            }
                        }
                        if (lane == 0) umma_commit(&xempty[hs]);
                        if (++hs == h_stages) { hs = 0; hphase ^= 1; }
                        continue;
                    }
                    int r = 0, s2 = 0;
                    uint32_t a_lo = halo_lo;   // tap (r, s): the halo rows shifted by r*Wp + s
                    for (int tap = 0; tap < R * S; tap += prm.tap_group) {
                        mbar_wait(&full[stage], phase, prm.err_flag, 3);
                        tc_fence_after();
                        uint32_t b_lo = stage_lo;
                        for (int t = 0; t < prm.tap_group; ++t, b_lo += b16) {
                            if (lane == 0) {
                                for (int k = 0; k < n_mma; ++k) {  // +32 bytes of K inside the swizzle atom = +2 units
                                    umma_i8_lohi(tmem_d, (a_lo + 2 * k) | lo_flag, (b_lo + 2 * k) | lo_flag, desc_hi, idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                            a_lo += kc16;
                            if (++s2 == S) { s2 = 0; ++r; a_lo = halo_lo + (uint32_t)r * wp16; }
                        }
                        if (lane == 0) umma_commit(&empty[stage]);
                        stage_lo += stage16;
                        if (++stage == stages) { stage = 0; stage_lo = base16; phase ^= 1; }
                    }
                    if (lane == 0) umma_commit(&xempty[hs]);  // halo slot reusable once every tap's MMAs have read it
                    if (++hs == h_stages) { hs = 0; hphase ^= 1; }
                }
            } else if (a_stat) {
                // A k-blocks from their own ring (waited for during the first channel tile, released after the last one);
                // the stages hold this channel tile's weights
                const bool first_n = nn == 0, last_n = nn == prm.n_tiles - 1;
                int as = as_base;
                uint32_t ap = as_phase;
                for (int kb = 0; kb < cblocks; ++kb) {
                    if (first_n) mbar_wait<200>(&full[kAStatBar + as], ap, prm.err_flag, 3);
                    mbar_wait<200>(&full[stage], phase, prm.err_flag, 3);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t a0 = (aring16 + (uint32_t)as * a16) | lo_flag, b0 = stage_lo | lo_flag;
                        umma_i8_lohi(tmem_d, a0, b0, desc_hi, idesc, accumulate);
                        umma_i8_lohi(tmem_d, a0 + 2, b0 + 2, desc_hi, idesc, 1u);      // KC = 64: two K = 32 slices
                        umma_commit(&empty[stage]);
                        if (last_n) umma_commit(&empty[kAStatBar + as]);
                    }
                    accumulate = 1;
                    stage_lo += stage16;
                    if (++stage == stages) { stage = 0; stage_lo = base16; phase ^= 1; }
                    if (++as == prm.a_slots) { as = 0; ap ^= 1; }
                }
                if (last_n) { as_base = as; as_phase = ap; nn = 0; }
                else ++nn;
            } else {
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait<(kFQ && !kStem) ? 200 : 0>(&full[stage], phase, prm.err_flag, 3);
                    tc_fence_after();
                    if (kPair) {
                        if (lane == 0) {
                            const uint32_t a0 = stage_lo | lo_flag, b0 = (stage_lo + a16) | lo_flag;
                            for (int k = 0; k < n_mma; ++k)
                                umma_i8_lohi_pair(tmem_d, a0 + 2 * k, b0 + 2 * k, desc_hi, idesc, k ? 1u : accumulate);
                            umma_commit_pair(&empty[stage]);   // the slot is reusable in BOTH CTAs
                        }
                    } else if (lane == 0) {
                        const uint32_t a0 = stage_lo | lo_flag,
                                       b0 = (kStem ? bres16 + (uint32_t)kb * (b_bytes >> 4) : stage_lo + a16) | lo_flag;
                        if (kFQ && !kStem && (prm.st_dbg & 32)) {   // ablation (bit 5): no MMAs, only the commits
                        } else
                        if (n_mma == 4) {          // KC = 128
                            umma_i8_lohi(tmem_d, a0, b0, desc_hi, idesc, accumulate);
                            umma_i8_lohi(tmem_d, a0 + 2, b0 + 2, desc_hi, idesc, 1u);
                            umma_i8_lohi(tmem_d, a0 + 4, b0 + 4, desc_hi, idesc, 1u);
                            umma_i8_lohi(tmem_d, a0 + 6, b0 + 6, desc_hi, idesc, 1u);
                        } else if (n_mma == 2) {   // KC = 64
                            umma_i8_lohi(tmem_d, a0, b0, desc_hi, idesc, accumulate);
                            umma_i8_lohi(tmem_d, a0 + 2, b0 + 2, desc_hi, idesc, 1u);
                        } else {
                            umma_i8_lohi(tmem_d, a0, b0, desc_hi, idesc, accumulate);
                        }
                        umma_commit(&empty[stage]);  // smem slot reusable once these MMAs have read it
                    }
                    accumulate = 1;
                    stage_lo += stage16;
                    if (++stage == stages) { stage = 0; stage_lo = base16; phase ^= 1; }
                }
            }
            if (lane == 0) {                                       // accumulator complete
                if constexpr (kPair) umma_commit_pair(&acc_full[buf]);
                else umma_commit(&acc_full[buf]);
            }
            __syncwarp();
            if (++buf == prm.n_acc) { buf = 0; acc_phase ^= 1; }
        }
    } else if (kFQ && warp == 2 + kEpiWarps) {
        // ===================== TMA producer of the fp32 input tiles (fused-quantize variant) =====================
        if (lane == 0 && (kStem || prm.x_cpasync == 0)) {
            // The shared-memory ring holds only x_stages tiles; HBM latency is covered by prefetching the tiles of the
            // next kPrefetch k-blocks into L2 (the ring loads then hit L2).
            constexpr int kPrefetch = 0;   // measured: prefetching 10 k-blocks ahead thrashes L2 (1.7x DRAM reads); the ring alone is better
            int pf_tile = blockIdx.x, pf_cb = 0;
            auto prefetch_next = [&]() {
                if (pf_tile >= total_tiles) return;
                const int m_tile = pf_tile / prm.n_tiles;
                const int img = m_tile / max(prm.tiles_per_img, 1), t = m_tile - img * prm.tiles_per_img;
                tma_prefetch_3d(&tmap_a, t * kBM, pf_cb * KC, img);
                if (++pf_cb == prm.cblocks) { pf_cb = 0; pf_tile += gridDim.x; }
            };
            for (int i = 0; i < kPrefetch; ++i) prefetch_next();
            int xs = 0;
            uint32_t xphase = 0;
            int st_img = -1, st_done = 0;
            for (int tile = (kStem || a_stat) ? unit0 : (int)blockIdx.x; tile < total_tiles;
                 tile += kStem ? 1 : (a_stat ? prm.n_tiles : (int)gridDim.x)) {     // (A-stationary: once per pixel tile)
                const int m_tile = a_stat ? (int)blockIdx.x + (tile / prm.n_tiles) * (int)gridDim.x : tile / prm.n_tiles;
                // flat pixel tiling (tiles_per_img == 0): the tile is 128 consecutive pixels of the batch and may end in the
                // next image — then a second box (same shape, start pixel off0 - H*W < 0: its first H*W - off0 pixels are
                // zero-filled) goes to the next ring slot and the quantizer threads read their 4 pixels from one or the other
                const bool flat = !kStem && prm.tiles_per_img == 0;
                const int img = flat ? prm.fd_pq.div(m_tile * kBM) : m_tile / prm.tiles_per_img;
                const int t = flat ? 0 : m_tile - img * prm.tiles_per_img;
                const int off0 = flat ? m_tile * kBM - img * (g.P * g.Q) : t * kBM;
                const bool two = flat && off0 + kBM > g.P * g.Q && img + 1 < g.N;
                if constexpr (kStem) {
                    // the input rows this tile adds to the ring (same bookkeeping as the quantizer warps), as boxes of
                    // `stride` rows x all channels; rows outside the image arrive as zeros
                    const int pq0 = t * kBM;
                    const int p_first = prm.fd_q.div(pq0), p_last = prm.fd_q.div(min(pq0 + kBM, g.P * g.Q) - 1);
                    const int h0 = p_first * g.stride - g.pad, h_end = p_last * g.stride - g.pad + g.R;
                    if (img != st_img) { st_img = img; st_done = h0; }
                    const int h_new = max(st_done, h0), n_new = h_end - h_new;
                    st_done = h_end;
                    if (n_new <= 0) continue;
                    const int nb = (n_new + g.stride - 1) / g.stride;
                    for (int b = 0; b < nb; ++b) {       // one ring slot and one barrier pair per box
                        mbar_wait(&bempty[xs], xphase ^ 1, prm.err_flag, 6);
                        mbar_expect_tx(&bfull[xs], (uint32_t)prm.st_box_bytes);
                        tma_load_4d(xring + (size_t)xs * x_bytes, &tmap_a, &bfull[xs], 0, h_new + b * g.stride, 0, img);
                        if (++xs == prm.x_stages) { xs = 0; xphase ^= 1; }
                    }
                    continue;
                }
                for (int cb = 0; cb < prm.cblocks; ++cb) {
                    if (kPrefetch > 0) prefetch_next();
                    mbar_wait<200>(&xempty[xs], xphase ^ 1, prm.err_flag, 6);
                    mbar_expect_tx(&xfull[xs], x_bytes);
                    // box [128 pixels][KC channels] of image img; pixels beyond H*W are zero-filled
                    tma_load_3d(xring + (size_t)xs * x_bytes, &tmap_a, &xfull[xs], off0, cb * KC, img);
                    if (++xs == prm.x_stages) { xs = 0; xphase ^= 1; }
                    if (two) {
                        mbar_wait<200>(&xempty[xs], xphase ^ 1, prm.err_flag, 6);
                        mbar_expect_tx(&xfull[xs], x_bytes);
                        tma_load_3d(xring + (size_t)xs * x_bytes, &tmap_a, &xfull[xs], off0 - g.P * g.Q, cb * KC, img + 1);
                        if (++xs == prm.x_stages) { xs = 0; xphase ^= 1; }
                    }
                }
            }
        }
    } else if (kFQ && warp > 2 + kEpiWarps) {
        // ===================== quantizer warps (fused-quantize variant: 1x1 / stride 1) =====================
        // fp32 tile [KC channels][128 pixels] in shared memory -> u8 A tile [128 pixels][KC] in the swizzled K-major
        // layout the MMA reads.  A thread owns 4 consecutive pixels x 8 channels (8 conflict-free 16-byte loads),
        // quantizes them exactly like the standalone quantizer (quant_math.cuh) and stores four 8-byte channel vectors.
        const QuantParams qp = load_params(prm.q_scale, prm.q_zero, prm.q_qmin, prm.q_qmax);
        if constexpr (kStem) {
            // ---- fused stem: fp32 input rows -> quantized byte planes -> grouped im2col rows (the A tile) ----
            // The planes are a ring of st_ring input rows per channel: consecutive tiles of the CTA's run share all but ~2 of
            // their input rows, so each row is quantized once per run, not once per tile (re-quantizing the 9-11 rows of
            // every tile made the kernel instruction-bound: 720 us instead of 450).  The warps' per-tile instruction chain
            // (~350 dependent-ish instructions at ~8 cycles each) is what bounds the kernel — not bytes, not issue slots — so
            // the loop is software-pipelined: the rows of tile t+1 are quantized BEFORE tile t is built (their box waits are
            // off the critical path), one block barrier per tile, one A stage (all k-blocks) and one arrival per tile.
            const int qw = warp - (3 + kEpiWarps);          // 0..7
            const int qt = qw * 32 + lane;                  // 0..255
            uint32_t* plane32 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + kTailBytes + 2 * prm.wcls_smem);
            const int Wq4 = prm.st_Wq >> 2, W4 = g.W >> 2, m4 = prm.st_margin >> 2;
            const int ring = prm.st_ring, rmask = ring - 1;
            const int PQs = g.P * g.Q;
            auto qbar = [] { asm volatile("bar.sync 8, %0;" ::"n"(kFqWarps * 32) : "memory"); };
            for (int i = qt; i < (prm.st_plane_bytes >> 2); i += kFqWarps * 32) plane32[i] = 0u;   // margins stay 0 for good
            qbar();
            const int row = qt & (kBM - 1), gh = qt >> 7;   // this thread's A row; its half of each k-block's 8 groups (warp-uniform)
            const uint32_t swz = (uint32_t)((row >> 1) & 3);
            const uint32_t plane_s = smem_u32(plane32);
            const uint32_t cplane = (uint32_t)(ring * Wq4) * 4u;            // bytes between channel planes
            // 16-byte chunk slots of this thread's A row (SWIZZLE_64B: chunk ^= (row >> 1) & 3), chunks 2*gh and 2*gh + 1
            const uint32_t a_row = (uint32_t)(row * kFqKC), ch0 = ((uint32_t)(2 * gh) ^ swz) << 4, ch1 = ((uint32_t)(2 * gh + 1) ^ swz) << 4;
            int xs = 0, stage = 0;
            uint32_t xphase = 0, phase = 0;
            // quantizer stream state: input rows [.., h_done) of image cur_img are in the ring; row ih lives in slot
            // (ih + ring_off) & rmask, slots are handed out consecutively across images
            int cur_img = -1, h_done = 0, ring_off = 0;
            // this thread's first two quads of a box: (line = channel * stride + row of the box, quad of 4 columns)
            const int nq = g.C * g.stride * W4;
            const int ln0 = qt < nq ? qt / W4 : -1, j0 = qt - max(ln0, 0) * W4, c0 = max(ln0, 0) / g.stride;
            const int qt1 = qt + kFqWarps * 32;
            const int ln1 = qt1 < nq ? qt1 / W4 : -1, j1 = qt1 - max(ln1, 0) * W4, c1 = max(ln1, 0) / g.stride;
            auto quantize_tile = [&](int tile) -> int {     // returns the tile's ring_off
                const int m_tile = prm.fd_ntiles.div(tile);
                const int img = prm.fd_tpi.div(m_tile), t = m_tile - img * prm.tiles_per_img;
                const int pq0 = t * kBM;
                const int p_first = prm.fd_q.div(pq0);
                const int p_last = prm.fd_q.div(min(pq0 + kBM, PQs) - 1);
                const int h0 = p_first * g.stride - g.pad;
                const int h_end = p_last * g.stride - g.pad + g.R;
                if (img != cur_img) {
                    ring_off = (cur_img < 0 ? 0 : h_done + ring_off) - h0;   // the new image's first row takes the next free slot
                    cur_img = img;
                    h_done = h0;
                }
                const int h_new = max(h_done, h0), n_new = h_end - h_new;     // rows to quantize now
                h_done = h_end;
                // box by box (`stride` rows x all channels each): a warp takes whole (channel, row) lines, a lane 4 columns
                for (int b0 = 0; b0 < n_new; b0 += g.stride) {
                    mbar_wait(&bfull[xs], xphase, prm.err_flag, 7);
                    const float* xt = reinterpret_cast<const float*>(xring + (size_t)xs * x_bytes);
                    if (!(prm.st_dbg & 1)) {
                        // the box's quads (4 columns of one (channel, row) line) spread evenly over the 256 threads
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int ln = u ? ln1 : ln0, j = u ? j1 : j0;
                            if (ln < 0) continue;
                            const int c = u ? c1 : c0, rin = ln - c * g.stride;
                            const int ih = h_new + b0 + rin;
                            if (ih >= h_end) continue;                         // (a box may reach past the tile's last row)
                            uint32_t word = 0u;
                            if ((unsigned)ih < (unsigned)g.H) {
                                const float4 v = lds4(xt + ln * g.W + 4 * j);
                                word = quant_word(v.x, v.y, v.z, v.w, qp);
                            }
                            plane32[(c * ring + ((ih + ring_off) & rmask)) * Wq4 + m4 + j] = word;
                        }
                        for (int q2 = qt + 2 * kFqWarps * 32; q2 < g.C * g.stride * W4; q2 += kFqWarps * 32) {   // (wider boxes: generic tail)
                            const int ln = q2 / W4, j = q2 - ln * W4, c = ln / g.stride, rin = ln - c * g.stride;
                            const int ih = h_new + b0 + rin;
                            if (ih >= h_end) continue;
                            uint32_t word = 0u;
                            if ((unsigned)ih < (unsigned)g.H) {
                                const float4 v = lds4(xt + ln * g.W + 4 * j);
                                word = quant_word(v.x, v.y, v.z, v.w, qp);
                            }
                            plane32[(c * ring + ((ih + ring_off) & rmask)) * Wq4 + m4 + j] = word;
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bempty[xs]);     // this warp has read its part of the box
                    if (++xs == prm.x_stages) { xs = 0; xphase ^= 1; }
                }
                return ring_off;
            };
            int off_next = unit0 < total_tiles ? quantize_tile(unit0) : 0;
            qbar();
            for (int tile = unit0; tile < total_tiles; ++tile) {
                const int off = off_next;
                if (tile + 1 < total_tiles) off_next = quantize_tile(tile + 1);   // one tile ahead of the build
                // ---- build tile `tile`: this thread's output pixel, window origin inside the planes ----
                const int m_tile = prm.fd_ntiles.div(tile);
                const int t = m_tile - prm.fd_tpi.div(m_tile) * prm.tiles_per_img;
                const int pq = min(t * kBM + row, PQs - 1);
                const int pi = prm.fd_q.div(pq), qi = pq - pi * g.Q;
                const int col = qi * g.stride - g.pad + prm.st_margin;      // first byte of the 8-byte windows
                const uint32_t wcol = plane_s + (uint32_t)(col & ~3);
                const int sh = (col & 3) * 8;
                const int slot0 = pi * g.stride - g.pad + off;              // ring slot of filter row 0 (before the mask)
                uint32_t rb[7];                                             // shared address of the window in filter row rf, channel 0
#pragma unroll
                for (int rf = 0; rf < 7; ++rf) rb[rf] = wcol + (uint32_t)(((slot0 + rf) & rmask) * Wq4) * 4u;
                mbar_wait(&empty[stage], phase ^ 1, prm.err_flag, 5);
                const uint32_t sa = smem_u32(smem) + (uint32_t)stage * stage_bytes + a_row;   // (smem = ring base, past the resident weights)
                // groups of this thread in k-block kb: gi = 8 * kb + 4 * gh + j  ->  channel gi / 7, filter row gi % 7 (R == 7)
                auto build = [&](auto gh_tag) {
                    constexpr int GH = decltype(gh_tag)::value;
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb) {
                        if (kb < prm.cblocks) {
                            // the 12 loads of the k-block's four windows in ONE asm statement: volatile asm statements keep
                            // their order, so per-group statements exposed a shared-memory latency per group
                            uint32_t a[4];
                            bool ok[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int gi = kb * 8 + GH * 4 + j;
                                const int c = gi / 7, rf = gi % 7;
                                ok[j] = c < g.C && !(prm.st_dbg & 2);        // (bytes past the last group meet zero weights)
                                a[j] = ok[j] ? rb[rf] + (uint32_t)c * cplane : rb[0];
                            }
                            uint32_t w[4][3];
                            asm volatile(
                                "ld.shared.u32 %0, [%12];\n\tld.shared.u32 %1, [%12+4];\n\tld.shared.u32 %2, [%12+8];\n\t"
                                "ld.shared.u32 %3, [%13];\n\tld.shared.u32 %4, [%13+4];\n\tld.shared.u32 %5, [%13+8];\n\t"
                                "ld.shared.u32 %6, [%14];\n\tld.shared.u32 %7, [%14+4];\n\tld.shared.u32 %8, [%14+8];\n\t"
                                "ld.shared.u32 %9, [%15];\n\tld.shared.u32 %10, [%15+4];\n\tld.shared.u32 %11, [%15+8];"
                                : "=r"(w[0][0]), "=r"(w[0][1]), "=r"(w[0][2]), "=r"(w[1][0]), "=r"(w[1][1]), "=r"(w[1][2]), "=r"(w[2][0]),
                                  "=r"(w[2][1]), "=r"(w[2][2]), "=r"(w[3][0]), "=r"(w[3][1]), "=r"(w[3][2])
                                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]));
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                if (ok[j]) {
                                    const uint32_t dst = sa + (uint32_t)kb * a_bytes + (j < 2 ? ch0 : ch1) + (uint32_t)((j & 1) << 3);
                                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(__funnelshift_r(w[j][0], w[j][1], sh)),
                                                 "r"(__funnelshift_r(w[j][1], w[j][2], sh)) : "memory");
                                }
                            }
                        }
                    }
                };
                if (gh == 0) build(std::integral_constant<int, 0>{});
                else build(std::integral_constant<int, 1>{});
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
                qbar();   // rows of tile + 1 complete before its build; this build complete before rows of tile + 2 reuse slots
            }
        } else {
        // All quantizer warps work on the same k-block (every barrier sees every phase: no parity aliasing):
        // warp pw owns channels [8*pw, 8*pw + 8) of the 64-channel k-block for all 128 pixels.
        const int pw = warp - (3 + kEpiWarps);
        const int j16 = pw >> 1;          // 16-byte channel chunk of the A row this warp contributes to
        const int half = pw & 1;          // which 8 bytes of that chunk
        const int rot = lane >> 1;
        int xs = 0, stage = 0;
        uint32_t xphase = 0, phase = 0;
        // (A-stationary: once per pixel tile, into the A ring — `stage` then walks its a_slots slots)
        const int q_stages = a_stat ? prm.a_slots : stages;
        uint64_t* const qfull = a_stat ? full + kAStatBar : full;
        uint64_t* const qempty = a_stat ? empty + kAStatBar : empty;
        uint8_t* const qbase = a_stat ? aring : smem;
        const uint32_t q_stride = a_stat ? a_bytes : stage_bytes;
        const bool flat = prm.tiles_per_img == 0;
        const int PQf = g.P * g.Q;
        // ---- cp.async input (x_cpasync): every thread streams ITS 8 channels x 4 pixels of each k-block through its own
        // 128 bytes of the ring slots, x_stages k-blocks ahead — planes of 14x14 / 28x28 pixels as 128-pixel TMA boxes cost
        // the TMA unit a full 512-byte row per channel whether 8 or 128 of its pixels are valid (two boxes per k-block for a
        // tile that ends in the next image), and that, not HBM, bounded those layers (1024 -> 256 @14x14: 1.45 us per k-block)
        const bool xcp = prm.x_cpasync != 0;
        const int tile_first = a_stat ? 0 : (int)blockIdx.x, tile_step = a_stat ? prm.n_tiles : (int)gridDim.x;
        struct { int tile, cb, slot; bool live; uint32_t bytes; const float* base; } nx = {0, 0, 0, false, 0u, nullptr};
        auto nx_set_tile = [&](int t) {
            nx.tile = t;
            nx.cb = 0;
            nx.live = t < total_tiles;
            if (nx.live) {
                const int u = prm.fd_ntiles.div(t);
                const int m_tile = a_stat ? (int)blockIdx.x + u * (int)gridDim.x : u;
                const int64_t m = (int64_t)m_tile * kBM + lane * 4;               // flat tiling; 4 pixels never straddle images
                const bool ok = m < prm.M;
                const int img = ok ? prm.fd_pq.div((int)m) : 0;
                nx.bytes = ok ? 16u : 0u;                                          // rows past the last pixel: zeros
                nx.base = prm.x + ((int64_t)img * g.C + pw * 8) * PQf + (ok ? (int)(m - (int64_t)img * PQf) : 0);
            }
        };
        auto nx_issue = [&]() {       // the thread's part of the next k-block (if any); always one commit group
            if (nx.live) {
                float* dst = reinterpret_cast<float*>(xring + (size_t)nx.slot * x_bytes) + (pw * 8) * kBM + lane * 4;
                const float* src = nx.base + (int64_t)nx.cb * kFqKC * PQf;
#pragma unroll
                for (int i = 0; i < 8; ++i) cp_async16_zfill(dst + i * kBM, src + (int64_t)i * PQf, nx.bytes);
                if (++nx.cb == prm.cblocks) nx_set_tile(nx.tile + tile_step);
            }
            cp_async_commit();
            if (++nx.slot == prm.x_stages) nx.slot = 0;
        };
        if (xcp) {
            nx_set_tile(tile_first);
            for (int i = 0; i < prm.x_stages - 1; ++i) nx_issue();
        }
        for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
            // flat pixel tiling: a tile that ends in the next image arrives as two boxes (see the producer); this thread's
            // 4 pixels (H*W % 4 == 0: never split) come from the first box when they belong to the first image
            bool two = false, second = false;
            if (flat && !xcp) {
                const int u = prm.fd_ntiles.div(tile);
                const int m_tile = a_stat ? (int)blockIdx.x + u * (int)gridDim.x : u;
                const int img = prm.fd_pq.div(m_tile * kBM), off0 = m_tile * kBM - img * PQf;
                two = off0 + kBM > PQf && img + 1 < g.N;
                second = two && off0 + lane * 4 >= PQf;
            }
            for (int cb = 0; cb < prm.cblocks; ++cb) {
                if (xcp) {
                    // refill the slot read in the previous iteration (its values were consumed by the quantizer), then wait
                    // for this iteration's group: x_stages - 1 younger groups may stay in flight
                    nx_issue();
                    switch (prm.x_stages) {
                        case 2: cp_async_wait<1>(); break;
                        case 3: cp_async_wait<2>(); break;
                        case 4: cp_async_wait<3>(); break;
                        case 5: cp_async_wait<4>(); break;
                        default: cp_async_wait<5>(); break;
                    }
                } else {
                    mbar_wait(&xfull[xs], xphase, prm.err_flag, 7);
                }
                const int xs0 = xs;
                if (two) {
                    if (++xs == prm.x_stages) { xs = 0; xphase ^= 1; }
                    mbar_wait(&xfull[xs], xphase, prm.err_flag, 7);
                }
                const float* xt = reinterpret_cast<const float*>(xring + (size_t)(second ? xs : xs0) * x_bytes) + (pw * 8) * kBM + lane * 4;
                float4 v[8];
                uint32_t w[4][2];  // [pixel][word]
                if (prm.st_dbg & 8) {   // ablation (QB200_STEM_DBG bit 3): no shared loads, no quantizer arithmetic
#pragma unroll
                    for (int t = 0; t < 4; ++t) w[t][0] = w[t][1] = 0x01010101u * (uint32_t)cb;
                } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = lds4(xt + i * kBM);      // explicit LDS.128 (a generic load has several times the latency)
                quant_tile<2>(v, w, qp);
                }
                // One arrival per WARP (after a warp sync), not per thread: every mbarrier arrival wakes the warps that
                // sleep on any barrier of the CTA, and 512 arrivals per k-block kept the idle epilogue warps spinning
                // through a third of the issue slots (ncu: 14.6 M NANOSLEEP wake-ups on one layer).
                __syncwarp();
                if (lane == 0 && !xcp) {
                    mbar_arrive(&xempty[xs]);  // this warp's part of the fp32 tile is in registers
                    if (two) mbar_arrive(&xempty[xs0]);
                }
                mbar_wait(&qempty[stage], phase ^ 1, prm.err_flag, 5);
                uint8_t* sa = qbase + (size_t)stage * q_stride;
                // lanes rotate which of their 4 pixels they store in each step so that one store instruction spreads
                // over all swizzle phases; SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3.  The rotation is applied
                // to the registers with selects (a dynamic register index compiles to a branchy loop that cost a third
                // of the quantizer's time), after which step t stores w[t] to pixel (t + rot) & 3.
                if (rot & 1) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint32_t t0 = w[0][k];
                        w[0][k] = w[1][k]; w[1][k] = w[2][k]; w[2][k] = w[3][k]; w[3][k] = t0;
                    }
                }
                if (rot & 2) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const uint32_t t0 = w[0][k], t1 = w[1][k];
                        w[0][k] = w[2][k]; w[1][k] = w[3][k]; w[2][k] = t0; w[3][k] = t1;
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int px = (t + rot) & 3;
                    const int row = lane * 4 + px;
                    const int jj = j16 ^ ((row >> 1) & 3);
                    sts2(sa + row * kFqKC + (jj << 4) + (half << 3), w[t][0], w[t][1]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) mbar_arrive(&qfull[stage]);
                if (++xs == prm.x_stages) { xs = 0; xphase ^= 1; }
                if (++stage == q_stages) { stage = 0; phase ^= 1; }
            }
        }
        }   // !kStem
    } else {
        // ===================== epilogue =====================
        // fused stem: the eight warps form TWO groups of four (one warp per lane quadrant, all columns) on alternate tiles —
        // with 64-channel tiles the per-tile latency chain of the epilogue (address set-up, waits), not its instruction
        // count, bounded the kernel: ~2200 cycles per tile with every store and FMA removed
        constexpr int kGW = kStem ? 4 : kEpiWarps;           // warps per epilogue group
        constexpr int kNG = kStem ? 2 : kGroups;             // groups
        const int grp = kStem ? (warp - 2) >> 2 : (kGroups > 1 ? (warp - 2) >> 3 : 0);   // epilogue group: tiles grp, grp + kNG, ... of this CTA
        const int e = (warp - 2) & (kGW - 1);
        const int quad = warp & 3;   // TMEM lane quadrant this warp may read
        const int half = kStem ? 0 : e >> 2;     // which half of the tile's columns
        const int et = (threadIdx.x - 64) & (kGW * 32 - 1);   // thread index inside the group
        const int row = quad * 32 + lane;
        const int PQ = g.P * g.Q;
        const EpilogueParams& ep = prm.ep;
        const EpilogueScalars es = load_epilogue_scalars(ep);
        const bool acc_out = ep.out_kind == QB200_OUT_ACC;
        QuantParams q8p = {};
        if (kQ8 && ep.q8_out != nullptr) q8p = load_params(ep.q8_scale, ep.q8_zero, ep.q8_qmin, ep.q8_qmax);
        // int8-only output: relu(v) followed by the quantizer's clamp to [xlo, xhi] is one clamp to [max(xlo, 0), xhi]
        // (a NaN ends at the lower bound either way), so the separate ReLU is dropped
        // (only the branch-free quantizer applies xlo; quant_word_exact — ranges outside a byte — keeps the explicit ReLU)
        const bool relu_folded = kQ8 && !kRes && ep.q8_out != nullptr && !ep.store_f32 && ep.relu != 0 && q8p.byte_clamp != 0;
        if (relu_folded) q8p.xlo = fmaxf(q8p.xlo, 0.f);
        const int cols = kStem ? BN : BN >> 1;    // columns per warp: 32, 64 or 128 (stem: the whole tile)
        // ---- residual stream (fused tail) ----
        // The identity tensor is the largest read of a residual layer and a warp that loads one 32-channel chunk at a
        // time keeps too few bytes in flight for HBM latency (ncu: 40 % of the epilogue's samples waited on those
        // loads).  Each warp therefore streams ITS chunks — in the order it will consume them, across tiles — through a
        // private cp.async ring, kResDepth - 1 chunks ahead.  A lane copies 16 bytes = 4 pixels of one channel;
        // chunk layout in shared memory [32 channels][32 pixels], read back conflict-free with lane = pixel.
        const bool res_async = kRes && prm.res_async != 0;
        float* res_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kTailBytes + 2 * prm.wcls_smem) +
                       e * (kResDepth * kResChunkFloats);
        struct { int tile, c, kb; bool live, pix_ok; const float* base; } cur = {0, 0, 0, false, false, nullptr};
        int issue_slot = 0, cons_slot = 0;
        // res_async == 2: planes whose size is not a multiple of 4 pixels (7x7: 49) — 16-byte pieces do not line up, so a
        // lane copies its own pixel of every channel with 4-byte cp.async (32 per chunk instead of 8)
        const bool res4 = kRes && prm.res_async == 2;
        auto cur_set_tile = [&](int t) {
            cur.tile = t;
            cur.c = 0;
            cur.live = t < total_tiles;
            if (cur.live) {
                const int mt = prm.fd_ntiles.div(t), nt = t - mt * prm.n_tiles;
                const int64_t m = (int64_t)mt * kBM + quad * 32 + (res4 ? lane : 4 * (lane & 7));
                cur.pix_ok = m < prm.M;
                const int im = cur.pix_ok ? prm.fd_pq.div((int)m) : 0;
                cur.kb = nt * BN + half * cols + (res4 ? 0 : (lane >> 3));
                cur.base = ep.residual + ((int64_t)im * g.K + cur.kb) * PQ + (int)(m - (int64_t)im * PQ);
            }
        };
        auto issue_one = [&]() {  // next chunk of the stream (if any); always one commit group per call
            if (cur.live) {
                const float* src = cur.base + (int64_t)cur.c * 32 * PQ;
                const int k = cur.kb + cur.c * 32;
                if (res4) {
                    float* dst = res_s + issue_slot * kResChunkFloats + lane;
                    if (cur.pix_ok) {
#pragma unroll 8
                        for (int i = 0; i < 32; ++i)
                            if (k + i < g.K) cp_async4(dst + i * 32, src + (int64_t)i * PQ);
                    }
                } else {
                    float* dst = res_s + issue_slot * kResChunkFloats + (lane >> 3) * 32 + 4 * (lane & 7);
                    if (cur.pix_ok && k + 28 < g.K) {
                        // whole chunk inside the tensor (the common case): pointer increments instead of a predicate and a
                        // 64-bit multiply per copy (ncu, conv3 @56x56: 9.6 instructions per copy, 9 % of the kernel's issue slots)
                        const uint32_t d32 = smem_u32(dst);
                        const char* sp = reinterpret_cast<const char*>(src);
                        const int64_t step = (int64_t)PQ * 16;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32 + (uint32_t)i * 512u), "l"(sp) : "memory");
                            sp += step;
                        }
                    } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (cur.pix_ok && k + 4 * i < g.K) cp_async16(dst + i * 128, src + (int64_t)i * 4 * PQ);
                    }
                }
                if (++cur.c == (cols >> 5)) cur_set_tile(cur.tile + (int)gridDim.x);
            }
            cp_async_commit();
            issue_slot = issue_slot == kResDepth - 1 ? 0 : issue_slot + 1;
        };
        if (res_async) {
            cur_set_tile(blockIdx.x);
#pragma unroll
            for (int i = 0; i < kResDepth - 1; ++i) issue_one();
        }
        const int acc_shift = prm.n_acc == 4 ? 2 : 1;
        int iter = 0;
        const uint32_t acc_empty_lead = kPair ? mapa_u32(smem_u32(&acc_empty[0]), 0) : 0u;   // the leader's barriers
        for (int tile = unit0 + grp * unit_step, it = grp; tile < total_tiles;
             tile += kNG * unit_step, it += kNG, ++iter) {
            const int buf = it & (prm.n_acc - 1);                       // the MMA warp fills the buffers in tile order
            const uint32_t acc_phase = (uint32_t)(it >> acc_shift) & 1u;
            const int u_tile = prm.fd_ntiles.div(tile);                 // m-tile, or pair of m-tiles
            const int n_tile = tile - u_tile * prm.n_tiles;
            const int m_tile = kPair ? 2 * u_tile + (int)pair_rank
                                     : (a_stat ? (int)blockIdx.x + u_tile * (int)gridDim.x : u_tile);   // (pair, odd tail: m_tile == m_tiles, no valid rows)
            const int k_base = n_tile * BN;
            // ---- per-tile channel constants (once per CTA when there is a single channel tile) ----
            float* sc = consts + (prm.n_tiles > 1 ? (iter & 1) : 0) * kConstFloats;   // two constant buffers, by tile parity
            float* be = sc + 256;
            float* br = sc + 512;
            const int tbl = (g.R + 1) * (g.S + 1);
            const int32_t* wtab = ep.wpre + (int64_t)k_base * tbl;  // prefix tables of this tile's channels
            const bool use_cls = prm.wcls_smem > 0 && es.z_a != 0.f;
            float* wcls = wcls_s + (prm.n_tiles > 1 ? (iter & 1) : 0) * (prm.wcls_smem / 4);
            if (iter == 0 || prm.n_tiles > 1) {
                if (use_cls) {
                    const int s1c = g.S + 1;
                    for (int i = et; i < prm.n_rcls * prm.n_ccls * BN; i += kGW * 32) {
                        const int cls = i / BN, kk = i - cls * BN;
                        float val = 0.f;
                        if (k_base + kk < g.K) {
                            const int rc = cls / prm.n_ccls, cx = cls - rc * prm.n_ccls;
                            const int r0 = prm.rcls[rc][0], r1 = prm.rcls[rc][1], c0 = prm.ccls[cx][0], c1 = prm.ccls[cx][1];
                            const int32_t* t4 = wtab + kk * tbl;
                            val = (float)(__ldg(t4 + r1 * s1c + c1) - __ldg(t4 + r0 * s1c + c1) - __ldg(t4 + r1 * s1c + c0) +
                                          __ldg(t4 + r0 * s1c + c0));
                        }
                        wcls[i] = val;
                    }
                }
                for (int ek = et; ek < BN; ek += kGW * 32) {
                    const int k = k_base + ek;
                    float scale = 0.f, bias = 0.f, beff = 0.f;  // beff: sum of all taps' weights (interior window), as float
                    if (k < g.K) {
                        scale = __fmul_rn(es.s_a, __ldg(ep.w_scale + (ep.per_tensor_w ? 0 : k)));
                        bias = ep.bias ? __ldg(ep.bias + k) : 0.f;
                        if (es.z_a != 0.f) beff = (float)__ldg(ep.wpre + (int64_t)(k + 1) * (g.R + 1) * (g.S + 1) - 1);
                    }
                    sc[ek] = scale;
                    be[ek] = beff;
                    br[ek] = bias;
                }
                epi_barrier<kGW * 32>(grp);
            }
            bool row_ok;
            int img, pq;
            if (kFQ && (kStem || prm.tiles_per_img > 0)) {  // tiles never straddle images
                img = prm.fd_tpi.div(m_tile);
                pq = (m_tile - img * prm.tiles_per_img) * kBM + row;
                row_ok = pq < PQ;
                if (!row_ok) pq = 0;
            } else if (halo) {  // rows are positions of the padded-flat output grid of one image
                img = prm.fd_tpi.div(m_tile);
                const int fl = (m_tile - img * prm.tiles_per_img) * kBM + row;
                const int pp = prm.fd_wp.div(fl), qq = fl - pp * prm.Wp;
                row_ok = pp < g.P && qq < g.Q;
                pq = row_ok ? pp * g.Q + qq : 0;
            } else {
                const int64_t m = (int64_t)m_tile * kBM + row;
                row_ok = m < prm.M;
                img = row_ok ? (prm.M < (1ll << 31) ? prm.fd_pq.div((int)m) : (int)(m / PQ)) : 0;
                pq = row_ok ? (int)(m - (int64_t)img * PQ) : 0;
            }
            const int p = prm.fd_q.div(pq), q = pq - p * g.Q;
            // int8-only hand-off instantiation (two epilogue groups): its per-tile bookkeeping is a third of the warp's
            // instructions on 64-channel tiles, so the tap-window work is skipped when the zero point is 0 (inputs that
            // follow a ReLU), where every pixel is "interior" by definition
            constexpr bool kLightSetup = kQ8 && kGroups == 2;
            const bool need_win = !kLightSetup || es.z_a != 0.f;
            PixelWindow pw = {0, g.R, 0, g.S};
            if (need_win) pw = pixel_window(g, p, q);
            const bool interior = es.z_a == 0.f || (pw.r0 == 0 && pw.r1 == g.R && pw.s0 == 0 && pw.s1 == g.S);
            // without a class table, a warp with any border pixel computes the window form for all its lanes
            const bool warp_interior = need_win ? __all_sync(0xffffffffu, interior || !row_ok) : true;
            const bool full_n = k_base + BN <= g.K;
            // this pixel's row of window sums: its class in the shared table, else the full-window sums
            const float* wrow = be;
            if (use_cls) {
                int rc = 0, cx = 0;
                if (prm.cls_fast) {   // classes in order of first occurrence: head rows, the full window, tail rows (host-verified)
                    rc = min(p, prm.cls_head_r) + max(0, p - prm.cls_tail_r + 1);
                    cx = min(q, prm.cls_head_c) + max(0, q - prm.cls_tail_c + 1);
                } else {
                    for (int i = 0; i < prm.n_rcls; ++i)
                        if (prm.rcls[i][0] == pw.r0 && prm.rcls[i][1] == pw.r1) rc = i;
                    for (int i = 0; i < prm.n_ccls; ++i)
                        if (prm.ccls[i][0] == pw.s0 && prm.ccls[i][1] == pw.s1) cx = i;
                }
                wrow = wcls + (rc * prm.n_ccls + cx) * BN;
            }
            const bool uniform_ok = use_cls || warp_interior;
            const int s1 = g.S + 1;
            const int i11 = pw.r1 * s1 + pw.s1, i01 = pw.r0 * s1 + pw.s1, i10 = pw.r1 * s1 + pw.s0, i00 = pw.r0 * s1 + pw.s0;
            const int64_t o_base = ((int64_t)img * g.K + k_base) * PQ + pq;

            if (kFQ && !kStem) mbar_wait_sleepy<500>(&acc_full[buf], acc_phase, prm.err_flag, 4);   // (stem tiles are ~1 us: no sleeps on its critical path)
            else mbar_wait(&acc_full[buf], acc_phase, prm.err_flag, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * prm.acc_stride + half * cols);
            for (int c0 = 0; c0 < cols; c0 += 32) {
                const int cc = half * cols + c0;  // first column of this chunk inside the tile
                // optional fused tail (beyond the reference op): out = relu(out + residual), rounded like the separate
                // torch ops (conv result rounded, then one add, then max).  The residual values of the whole chunk are
                // requested before the accumulator is read, so their latency overlaps the TMEM load.
                const int64_t o_off = o_base + (int64_t)cc * PQ;
                const bool has_res = kRes && ep.residual != nullptr && !acc_out;
                float rv[32];
                if (res_async) {
                    cp_async_wait<kResDepth - 2>();   // this chunk has landed (only younger groups may still be in flight)
                    __syncwarp();
                    const float* rs = res_s + cons_slot * kResChunkFloats + lane;
#pragma unroll
                    for (int j = 0; j < 32; ++j) rv[j] = lds1(rs + j * 32);
                    issue_one();   // into the slot of the PREVIOUS chunk, which every lane finished reading before the sync above
                    cons_slot = cons_slot == kResDepth - 1 ? 0 : cons_slot + 1;
                } else if (has_res && row_ok) {
                    const float* rs = ep.residual + o_off;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        rv[j] = (k_base + cc + j < g.K) ? __ldg(rs + (int64_t)j * PQ) : 0.f;
                }
                bool full_c = full_n;
                if constexpr (kRagged) {
                    full_c = k_base + cc + 32 <= g.K;
                    if (k_base + cc >= g.K) continue;
                }
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                if (!row_ok) continue;
                const bool any_tail = has_res || ep.relu;
                auto tail = [&](float val, int j) {
                    if (has_res) val = __fadd_rn(val, rv[j]);
                    if (ep.relu && !relu_folded) val = fmaxf(val, 0.f);
                    return val;
                };
                if constexpr (kQ8) {
                    if (ep.q8_out != nullptr) {
                        // Quantized hand-off: the chunk's 32 channels of this pixel, after the tail, go through the
                        // CONSUMER's activation quantizer and land as 32 contiguous bytes of its NHWC workspace.
                        // (kept compact on purpose: the first version branched per element inside the unrolled loop and
                        // the epilogue became instruction-fetch bound — ncu: stall_no_inst on 60 % of its samples)
                        float r[32];
                        if (uniform_ok && es.z_a == 0.f) {
                            // zero point 0 (inputs that follow a ReLU — every hand-off layer of a ResNet): fma(z_a, wsum, acc)
                            // is exactly acc, so the FMA and the window-sum load are skipped (ncu, conv3 @56x56: 5 % of the
                            // issue slots of an epilogue that is bound by them)
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 s4 = lds4(sc + cc + j), b4 = lds4(br + cc + j);
                                r[j + 0] = tail(__fmaf_rn(s4.x, (float)(int32_t)v[j + 0], b4.x), j + 0);
                                r[j + 1] = tail(__fmaf_rn(s4.y, (float)(int32_t)v[j + 1], b4.y), j + 1);
                                r[j + 2] = tail(__fmaf_rn(s4.z, (float)(int32_t)v[j + 2], b4.z), j + 2);
                                r[j + 3] = tail(__fmaf_rn(s4.w, (float)(int32_t)v[j + 3], b4.w), j + 3);
                            }
                        } else if (uniform_ok) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 s4 = lds4(sc + cc + j), w4 = lds4(wrow + cc + j), b4 = lds4(br + cc + j);
                                r[j + 0] = tail(__fmaf_rn(s4.x, __fmaf_rn(es.z_a, w4.x, (float)(int32_t)v[j + 0]), b4.x), j + 0);
                                r[j + 1] = tail(__fmaf_rn(s4.y, __fmaf_rn(es.z_a, w4.y, (float)(int32_t)v[j + 1]), b4.y), j + 1);
                                r[j + 2] = tail(__fmaf_rn(s4.z, __fmaf_rn(es.z_a, w4.z, (float)(int32_t)v[j + 2]), b4.z), j + 2);
                                r[j + 3] = tail(__fmaf_rn(s4.w, __fmaf_rn(es.z_a, w4.w, (float)(int32_t)v[j + 3]), b4.w), j + 3);
                            }
                        } else {
                            // border pixel of a layer without a class table (z_a != 0): window sums from the prefix tables
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float wsf = 0.f;
                                if (k_base + cc + j < g.K) {
                                    const int32_t* t4 = wtab + (cc + j) * tbl;
                                    wsf = (float)(__ldg(t4 + i11) - __ldg(t4 + i01) - __ldg(t4 + i10) + __ldg(t4 + i00));
                                }
                                r[j] = tail(__fmaf_rn(lds1(sc + cc + j), __fmaf_rn(es.z_a, wsf, (float)(int32_t)v[j]), lds1(br + cc + j)), j);
                            }
                        }
                        if (kGroups == 1 && ep.store_f32) {   // (the two-group instantiation is launched for int8-only output)
                            float* o = static_cast<float*>(out) + o_off;
                            if (full_n) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) o[(int64_t)j * PQ] = r[j];
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (k_base + cc + j < g.K) o[(int64_t)j * PQ] = r[j];
                            }
                        }
                        if (k_base + cc < ep.q8_cp) {
                            const int n_valid = g.K - (k_base + cc);  // channels of this chunk that exist (others stay 0)
                            uint32_t w[8];
                            quant_row<8>(r, w, q8p);
                            if (n_valid < 32) {
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj) {
                                    const int nb = n_valid - 4 * jj;
                                    if (nb < 4) w[jj] = nb <= 0 ? 0u : (w[jj] & ((1u << (8 * nb)) - 1u));
                                }
                            }
                            const int64_t pix = (int64_t)img * ep.q8_img_pixels + (int64_t)p * ep.q8_row_pixels + q + ep.q8_pixel_off;
                            uint4* dst = reinterpret_cast<uint4*>(ep.q8_out + pix * ep.q8_cp + (k_base + cc));
                            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                        continue;
                    }
                }
                if constexpr (kStem) { if (prm.st_dbg & 4) continue; }
                if constexpr (kFQ && !kStem) { if (prm.st_dbg & 16) continue; }   // ablation (bit 4): no dequantize / stores
                if (!acc_out && full_c && uniform_ok) {
                    float* o = static_cast<float*>(out) + o_off;
                    // the same two roundings as every other path (dequant_one): t = fma(z_a, wsum, acc); fma(sc, t, bias).
                    // kTail is a compile-time switch so that the plain op pays nothing for the optional fused tail.
                    // The constants of kPre column groups are fetched (LDS) before they are needed: all 8 groups of the
                    // chunk where registers allow it, 2 groups in the instantiations with 576 / 608 threads.
                    auto fast = [&](auto kTailTag, auto kZeroTag) {
                        constexpr bool kTail = decltype(kTailTag)::value;
                        constexpr bool kZ = decltype(kZeroTag)::value;     // activation zero point != 0
                        constexpr int kPre = kFQ ? QB200_FQ_KPRE : (kGroups > 1 ? 2 : (kZ ? 4 : 8));
#pragma unroll
                        for (int g0 = 0; g0 < 8; g0 += kPre) {
                            float4 S[kPre], B[kPre], Wz[kZ ? kPre : 1];
#pragma unroll
                            for (int g = 0; g < kPre; ++g) {
                                S[g] = lds4(sc + cc + 4 * (g0 + g));
                                B[g] = lds4(br + cc + 4 * (g0 + g));
                                if constexpr (kZ) Wz[g] = lds4(wrow + cc + 4 * (g0 + g));
                            }
#pragma unroll
                            for (int g = 0; g < kPre; ++g) {
                                const int j = 4 * (g0 + g);
                                float a0 = (float)(int32_t)v[j + 0], a1 = (float)(int32_t)v[j + 1], a2 = (float)(int32_t)v[j + 2],
                                      a3 = (float)(int32_t)v[j + 3];
                                if constexpr (kZ) {
                                    a0 = __fmaf_rn(es.z_a, Wz[g].x, a0);
                                    a1 = __fmaf_rn(es.z_a, Wz[g].y, a1);
                                    a2 = __fmaf_rn(es.z_a, Wz[g].z, a2);
                                    a3 = __fmaf_rn(es.z_a, Wz[g].w, a3);
                                }
                                float r0 = __fmaf_rn(S[g].x, a0, B[g].x);
                                float r1 = __fmaf_rn(S[g].y, a1, B[g].y);
                                float r2 = __fmaf_rn(S[g].z, a2, B[g].z);
                                float r3 = __fmaf_rn(S[g].w, a3, B[g].w);
                                if (kTail) { r0 = tail(r0, j); r1 = tail(r1, j + 1); r2 = tail(r2, j + 2); r3 = tail(r3, j + 3); }
                                const int64_t jo = (int64_t)j * PQ;
                                o[jo] = r0;
                                o[jo + PQ] = r1;
                                o[jo + 2 * (int64_t)PQ] = r2;
                                o[jo + 3 * (int64_t)PQ] = r3;
                            }
                        }
                    };
                    if (es.z_a == 0.f) {
                        if (any_tail) fast(std::true_type{}, std::false_type{});
                        else fast(std::false_type{}, std::false_type{});
                    } else {
                        if (any_tail) fast(std::true_type{}, std::true_type{});
                        else fast(std::false_type{}, std::true_type{});
                    }
                } else if (acc_out) {
                    int32_t* o = static_cast<int32_t*>(out) + o_off;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (k_base + cc + j < g.K) o[(int64_t)j * PQ] = (int32_t)v[j];
                } else {
                    // ragged channel tile and / or border pixel of a layer with a non-zero activation zero point
                    float* o = static_cast<float*>(out) + o_off;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k = k_base + cc + j;
                        if (k >= g.K) continue;
                        float r;
                        if (es.z_a == 0.f) {
                            r = __fmaf_rn(lds1(sc + cc + j), (float)(int32_t)v[j], lds1(br + cc + j));
                        } else {
                            float wsf;
                            if (use_cls) {
                                wsf = lds1(wrow + cc + j);
                            } else {
                                const int32_t* t4 = wtab + (cc + j) * tbl;
                                wsf = (float)(__ldg(t4 + i11) - __ldg(t4 + i01) - __ldg(t4 + i10) + __ldg(t4 + i00));
                            }
                            const float t = __fmaf_rn(es.z_a, wsf, (float)(int32_t)v[j]);
                            r = __fmaf_rn(lds1(sc + cc + j), t, lds1(br + cc + j));
                        }
                        o[(int64_t)j * PQ] = tail(r, j);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kPair) mbar_arrive_cluster(acc_empty_lead + (uint32_t)buf * 8u);
                else mbar_arrive(&acc_empty[buf]);
            }
        }
    }

    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();   // no CTA of the pair may exit while its partner can still signal its barriers
    else __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host: tensor maps through the driver entry points (no link-time dependency on libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct DriverApi {
    EncodeTiledFn tiled = nullptr;
    EncodeIm2colFn im2col = nullptr;
    int driver_version = 0;
    bool ok = false;
};

const DriverApi& driver_api() {
    static DriverApi api = [] {
        DriverApi a;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            a.tiled = reinterpret_cast<EncodeTiledFn>(f);
        f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            a.im2col = reinterpret_cast<EncodeIm2colFn>(f);
        cudaDriverGetVersion(&a.driver_version);
        a.ok = a.tiled && a.im2col;
        return a;
    }();
    return api;
}

// Watchdog flag in mapped pinned host memory: still readable after a trap has poisoned the context.
int* g_watchdog_host = nullptr;
int* watchdog_flag() {
    static int* dev_ptr = nullptr;
    if (!dev_ptr) {
        int* h = nullptr;
        if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
        *h = 0;
        int* d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) return nullptr;
        g_watchdog_host = h;
        dev_ptr = d;
    }
    return dev_ptr;
}

int num_sms() {
    static thread_local int cached[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return kNumSMs;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace

int watchdog_code() { return g_watchdog_host ? *(volatile int*)g_watchdog_host : 0; }

bool umma_supported(const ConvGeom& g) {
    if (g.groups != 1) return false;
    if (g.pad > 128 || g.pad - (g.R - 1) < -128 || g.pad - (g.S - 1) < -128) return false;  // im2col corner range
    if (g.stride > 8) return false;  // TMA element strides
    if ((int64_t)g.W * g.Cp >= (1ll << 32) || (int64_t)g.H * g.W * g.Cp >= (1ll << 40)) return false;
    return true;
}

// Halo variant: stride 1, spatial kernel, the whole halo of a 128-position tile fits one TMA box (<= 256 rows), and
// the padded-flat tiling does not waste more than ~25 % of the MMA rows (skips 7x7 feature maps).
bool umma_halo_supported(const ConvGeom& g) {
    if (!umma_supported(g) || g.stride != 1 || g.R * g.S == 1 || g.C <= 4) return false;
    const int Wp = g.W + 2 * g.pad;
    const int halo_rows = kBM + (g.R - 1) * Wp + (g.S - 1);
    if (halo_rows > 256) return false;
    const int span = (g.P - 1) * Wp + g.Q;                  // padded-flat positions that hold outputs of one image
    const int tiles = (span + kBM - 1) / kBM;
    return g.P * g.Q * 4 >= tiles * kBM * 3;
}

// Measured (profiles/README.md, A/B per layer): the halo variant wins where k-blocks are small (C = 64: 103 vs 122 us
// at 56x56) and loses where the weight tiles dominate the L2->SM traffic anyway (C >= 128) or the padded-flat tiling
// adds a wave (14x14).
bool umma_halo_profitable(const ConvGeom& g) { return g.Cp == 64 && g.P * g.Q >= 784; }

// CTA-pair variant (cta_group::2): deep reductions, whose main loop is bound by the L2 -> SM operand feed.  Needs whole
// channel tiles (the plain epilogue) and at least one pair of pixel tiles.  QB200_PAIR=0 turns it off (A/B measurements).
bool umma_pair_enabled() {
    static const bool on = [] {
        const char* e = getenv("QB200_PAIR");
        return !(e && e[0] == '0');
    }();
    return on;
}
// Measured per layer on ResNet-50 at batch 256 (profiles/README.md, round 2): the pair variant is a few percent faster on
// the spatial kernels with 256-wide channel tiles (3x3 @14: 37.5 -> 36.3 us, 3x3 @7: 42.0 -> 41.3, strided 3x3: 39.4 ->
// 37.3, 42.5 -> 40.9) and SLOWER with 128-wide tiles (3x3 @28, K = 128: 46.5 -> 53.7) and on deep 1x1 layers (41.6 ->
// 47.6): the main loop turned out to be bound by the latency of the stage ring (no unit above 46 % in ncu), not by the
// operand bytes the pair saves, and the cross-CTA barrier hops add to that latency.  The GEMM probe
// (tests/native/umma_pair.cu) shows the same ceiling: 2.47 -> 2.99 POPS at 8192^3.
bool umma_pair_profitable(const ConvGeom& gm, int K, int m_tiles, int min_kgemm) {
    const int64_t k_gemm = (int64_t)gm.R * gm.S * gm.Cp;
    return umma_pair_enabled() && gm.R * gm.S > 1 && k_gemm >= min_kgemm && K % 256 == 0 && m_tiles >= 2;
}
int umma_pair_min_kgemm() {
    static const int v = [] {
        const char* e = getenv("QB200_PAIR_MIN_K");
        return e ? atoi(e) : 2048;
    }();
    return v;
}

bool umma_fused_quant_supported(const ConvGeom& g, const float* x) {
    return umma_supported(g) && g.R == 1 && g.S == 1 && g.stride == 1 && g.pad == 0 && (g.H * g.W) % 4 == 0 &&
           g.C % 64 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
}

// Fused stem: a few-channel layer whose grouped im2col rows (8 bytes per (channel, filter row)) the quantizer warps can
// build in shared memory: the fp32 rows of one tile fit a ring slot, the byte planes fit their buffer, one channel tile.
namespace {
struct StemPlan { int rows_box, Wq, margin, box_bytes, plane_bytes, ring; };
bool stem_plan(const ConvGeom& g, StemPlan* out) {
    if (g.groups != 1 || g.C > 4 || g.R != 7 || !im2col_grouped(g.C, g.R, g.S)) return false;   // (the row builder is unrolled for R == 7)
    if (im2col8_row_bytes(g.C, g.R) % kFqKC != 0 || g.K > 256 || g.W % 4 != 0 || g.W > 256 || g.Q < 1) return false;
    StemPlan sp;
    const int rows_out = (kBM - 2) / g.Q + 2;              // output rows a 128-pixel tile can touch
    sp.rows_box = (rows_out - 1) * g.stride + g.R;
    sp.margin = (g.pad + 3) & ~3;
    const int need = std::max(sp.margin + g.W + g.pad, (g.Q - 1) * g.stride - g.pad + sp.margin + 12);
    sp.Wq = (need + 3) & ~3;
    sp.box_bytes = g.C * g.stride * g.W * 4;               // one TMA box: `stride` input rows of every channel
    sp.ring = 8;                                            // ring of input rows per channel plane (a power of two): the rows
    while (sp.ring < sp.rows_box + rows_out * g.stride) sp.ring <<= 1;   // of a tile plus those quantized ahead for the next
    sp.plane_bytes = (g.C * sp.ring * sp.Wq + 15) & ~15;
    if (im2col8_row_bytes(g.C, g.R) > 4 * kFqKC) return false;
    if (sp.rows_box > 256 || sp.box_bytes > 24 * 1024 || sp.plane_bytes > 32 * 1024) return false;
    if (out) *out = sp;
    return true;
}
}  // namespace
bool umma_stem_supported(const ConvGeom& g, const float* x) {
    return umma_supported(g) && stem_plan(g, nullptr) && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
           (int64_t)g.N * g.P * g.Q < (1ll << 31);
}

// Measured on B200 (profiles/README.md): the in-kernel quantizer (8 warps) sustains ~3.5 TB/s of fp32 input, the
// standalone quantizer ~5.5 TB/s.  Fusing wins when the layer is output-heavy and large: one 64-channel k-block per
// tile (C == 64) on feature maps of at least 28x28; everything else keeps the two-kernel path.
// Measured on ResNet-50's 1x1 layers at batch 256 (profiles/README.md): fusing wins when the layer is input-heavy (a
// single 64-channel k-block, or at least twice as many input as output channels) and each image has >= 7 tiles; layers
// with several channel tiles would quantize the same pixels once per channel tile, and 14x14 / 7x7 layers are ring-
// latency bound, so those keep the standalone quantizer.
// Round 2 (A/B over all 53 layers with the fused variant forced): 1024 -> 256 at 14x14 also wins (72 vs 78 us); layers with
// several channel tiles (128 -> 512: 113 vs 108 us, 256 -> 1024: 80 vs 72 us) lose because every channel tile quantizes
// the same pixels again.
// QB200_ASTAT: 0 = A-stationary mode off (several channel tiles re-quantize, as before), 1 = on (default)
int a_stat_mode() {
    static const int v = [] {
        const char* e = getenv("QB200_ASTAT");
        return e ? atoi(e) : 1;
    }();
    return v;
}
// A-stationary fused quantize: the pixel tile's quantized k-blocks (8 KB each) must fit their ring next to two fp32 slots
bool a_stat_fits(const ConvGeom& g) { return g.C / kFqKC <= kMaxStages - kAStatBar; }

bool umma_fused_quant_profitable(const ConvGeom& g) {
    // several channel tiles (K > 256), A-stationary: the pixels are quantized once and the second pass over the input is gone
    // (with flat pixel tiles: 128 -> 512 @28x28 99 vs 108 us for the two kernels, 256 -> 1024 @14x14 68 vs 72.5 us; with
    // image-aligned tiles the 14x14 layer took 79.5 us — 512 instead of 392 pixel tiles, and an output-bound layer pays for
    // every tile's epilogue)
    // (and only with at least one pixel tile per SM: a CTA runs its pixel tiles' channel tiles one after the other, so with
    // fewer pixel tiles than SMs the two-kernel path's (pixel tile, channel tile) grid keeps more of the chip busy —
    // 256 -> 1024 @14x14 at 32 images: 28.8 vs 26.7 us)
    if (g.K > 256 && a_stat_mode() != 0 && a_stat_fits(g) && g.H * g.W >= 196 && (int64_t)g.N * g.P * g.Q >= (int64_t)kBM * num_sms())
        return true;
    if (g.H * g.W >= 784) return g.C == 64 || g.C >= 2 * g.K;
    return g.H * g.W >= 196 && g.C >= 4 * g.K;
}

int launch_conv_umma(const ConvGeom& g, const uint8_t* qa, const uint8_t* wq, const EpilogueParams& ep, void* out,
                     cudaStream_t st, int gemm_rows, const float* x_fused, const qb200_act_quant* aq_fused, bool halo,
                     int pair_mode, bool fq_force) {
    QB_REQUIRE(umma_supported(g), QB200_EUNSUPPORTED, "conv_umma: shape not supported by the tensor-core kernel");
    const bool fq = x_fused != nullptr;
    QB_REQUIRE(!halo || (umma_halo_supported(g) && !fq && gemm_rows == 0), QB200_EINVAL,
               "conv_umma: layer not eligible for the halo variant");
    QB_REQUIRE(!(fq && (ep.residual || ep.q8_out)), QB200_EINVAL, "conv_umma: the fused-quantize kernel has no residual / hand-off tail");
    const bool stem = fq && gemm_rows > 0;   // fused stem: x_fused is the fp32 input of an im2col-rows layer
    StemPlan sp = {0, 0, 0, 0, 0, 0};
    if (stem) {
        QB_REQUIRE(umma_stem_supported(g, x_fused) && stem_plan(g, &sp) && gemm_rows == im2col8_row_bytes(g.C, g.R) && aq_fused &&
                       aq_fused->qmin && aq_fused->qmax,
                   QB200_EINVAL, "conv_umma: layer not eligible for the fused stem kernel");
        qa = reinterpret_cast<const uint8_t*>(x_fused);
    } else if (fq) {
        QB_REQUIRE(umma_fused_quant_supported(g, x_fused) && gemm_rows == 0 && aq_fused && aq_fused->qmin && aq_fused->qmax,
                   QB200_EINVAL, "conv_umma: layer not eligible for the fused-quantize kernel");
        qa = reinterpret_cast<const uint8_t*>(x_fused);  // only used for alignment checks below
    }
    ConvGeom gm = g;
    if (gemm_rows > 0) {  // materialised im2col rows: a 1x1 convolution over an [N, P, Q, gemm_rows] tensor
        gm.C = gm.Cg = gm.Cp = gm.Cgp = gemm_rows;
        gm.H = g.P;
        gm.W = g.Q;
        gm.R = gm.S = 1;
        gm.stride = 1;
        gm.pad = 0;
    }
    const DriverApi& api = driver_api();
    QB_REQUIRE(api.ok, QB200_EDRIVER, "cuTensorMapEncodeTiled/Im2col driver entry points unavailable");
    QB_REQUIRE(reinterpret_cast<uintptr_t>(qa) % 16 == 0 && reinterpret_cast<uintptr_t>(wq) % 16 == 0, QB200_EINVAL,
               "conv_umma: operands must be 16-B aligned");

    UmmaParams prm;
    prm.g = g;
    prm.gm = gm;
    prm.ep = ep;
    prm.M = (int64_t)g.N * g.P * g.Q;
    if (prm.M == 0) return 0;
    prm.KC = fq ? kFqKC : (gm.Cp % 128 == 0) ? 128 : (gm.Cp % 64 == 0) ? 64 : 32;
    prm.cblocks = gm.Cp / prm.KC;
    prm.halo = halo ? 1 : 0;
    {
        static const bool tiled_ok = [] {
            const char* e = getenv("QB200_A_TILED");
            return !(e && e[0] == '0');
        }();
        prm.a_tiled = (tiled_ok && !fq && !halo && gm.R == 1 && gm.S == 1 && gm.stride == 1 && gm.pad == 0 &&
                       (int64_t)gm.N * gm.H * gm.W < (1ll << 31)) ? 1 : 0;
    }
    prm.Hp = g.H + 2 * g.pad;
    prm.Wp = g.W + 2 * g.pad;
    prm.halo_rows = kBM + (g.R - 1) * prm.Wp + (g.S - 1);
    prm.halo_bytes = (int)align_up_sz((size_t)prm.halo_rows * prm.KC, 1024);
    prm.h_stages = 0;
    prm.x_stages = 0;
    prm.st_rows_box = sp.rows_box;
    prm.st_Wq = sp.Wq;
    prm.st_margin = sp.margin;
    prm.st_box_bytes = sp.box_bytes;
    prm.st_box_pitch = (sp.box_bytes + 127) & ~127;
    prm.st_plane_bytes = sp.plane_bytes;
    prm.st_ring = sp.ring;
    { const char* e = getenv("QB200_STEM_DBG"); prm.st_dbg = e ? atoi(e) : 0; }
    // fused quantize on planes that are not a multiple of 128 pixels: flat pixel tiling (a tile may end in the next image and
    // then arrives as two boxes) instead of image-aligned tiles — 14x14: 392 instead of 512 pixel tiles per 256 images,
    // 28x28: 6.125 instead of 7 per image.  Needs planes of at least one tile (at most two images per tile) and four fp32
    // slots (checked below).  QB200_FQ_FLAT=0 keeps the image-aligned tiles (A/B measurements).
    // the size-gated fused-quantize modes (flat tiles, cp.async input, A-stationary channel tiles) pay when every SM has at
    // least one pixel tile; below that (strong-scaled batches) the image-aligned TMA form with narrower channel tiles keeps
    // more SMs busy (1024 -> 256 @14x14 at 32 images: 43.4 vs 35.9 us)
    const bool fq_big = fq_force || prm.M >= (int64_t)kBM * num_sms();
    static const bool fq_flat_on = [] {
        const char* e = getenv("QB200_FQ_FLAT");
        return !(e && e[0] == '0');
    }();
    // (only where image-aligned tiles waste more than 5 % of their rows — 14x14: 31 %, 28x28: 14 %; at 56x56 (2 %) the
    // second box of every other tile costs more ring depth than the tiles save: 64 -> 256 200 vs 206 us)
    const int aligned_rows = (g.P * g.Q + kBM - 1) / kBM * kBM;
    bool fq_flat = fq && !stem && fq_flat_on && fq_big && g.P * g.Q >= kBM && 20 * (aligned_rows - g.P * g.Q) > g.P * g.Q &&
                   prm.M < (1ll << 31) - kBM;
    // QB200_FQ_CPASYNC: 0 = TMA boxes everywhere, 1 = cp.async input where flat tiles pay (default), 2 = every fused-quantize
    // layer (A/B measurements; cp.async handles any plane size, a tile may span several small images)
    static const int fq_cp_mode = [] {
        const char* e = getenv("QB200_FQ_CPASYNC");
        return e ? atoi(e) : 1;
    }();
    const bool x_cp = fq && !stem && prm.M < (1ll << 31) - kBM &&
                      (fq_cp_mode == 2 || (fq_cp_mode == 1 && fq_flat_on && fq_big && 20 * (aligned_rows - g.P * g.Q) > g.P * g.Q));
    prm.x_cpasync = x_cp ? 1 : 0;
    if (x_cp) fq_flat = true;
    prm.tiles_per_img = fq ? (fq_flat ? 0 : (g.P * g.Q + kBM - 1) / kBM) : (halo ? ((g.P - 1) * prm.Wp + g.Q + kBM - 1) / kBM : 0);
    prm.m_tiles = ((fq && !fq_flat) || halo) ? g.N * prm.tiles_per_img : (int)ceil_div64(prm.M, kBM);
    const int sms = num_sms();
    // out-channel tile: as wide as TMEM allows (fewest re-reads of A) unless that leaves SMs idle
    int BN = g.K >= 256 ? 256 : (g.K > 64 ? 128 : 64);
    const int64_t k_gemm = (int64_t)gm.R * gm.S * gm.Cp;
    // deep reductions on the plain main loop: a CTA pair per 256-pixel tile (cta_group::2), channel tile 256 (or 128)
    // pair_mode 0: never, 1: where measured to win, 2: wherever the variant is supported (tests)
    const bool pair_ok = !fq && !halo && ep.residual == nullptr && prm.m_tiles >= 2 && (g.K % 256 == 0 || g.K == 128);
    const bool pair = pair_ok && (pair_mode == 2 || (pair_mode == 1 && umma_pair_profitable(gm, g.K, prm.m_tiles, umma_pair_min_kgemm())));
    if (pair) {
        BN = g.K % 256 == 0 ? 256 : 128;
    } else if (k_gemm >= 1024 && !fq) {
        // deep reductions are bound by the operand feed (L2 -> SM, ~40 GB/s per SM): per tile k_gemm * (128 + BN) bytes,
        // and a partly filled last wave costs a whole tile time.  Pick the tile width with the least waves * bytes.
        int best = BN;
        int64_t best_cost = INT64_MAX;
        for (int cand = BN; cand >= 64; cand >>= 1) {
            const int64_t tiles = (int64_t)prm.m_tiles * ((g.K + cand - 1) / cand);
            const int64_t cost = ((tiles + sms - 1) / sms) * (kBM + cand);
            if (cost < best_cost) { best_cost = cost; best = cand; }
        }
        BN = best;
    } else if (!(fq && !stem && a_stat_mode() != 0 && a_stat_fits(g) && fq_big)) {
        // (fused quantize with the A-stationary mode: a CTA runs all channel tiles of its pixel tiles, narrower tiles add
        // no parallelism — and without it every extra channel tile quantizes the same pixels again)
        while (BN > 64 && (int64_t)prm.m_tiles * ((g.K + BN - 1) / BN) < 2 * sms) BN >>= 1;
    }
    if (stem) BN = g.K > 128 ? 256 : (g.K > 64 ? 128 : 64);   // one channel tile: the rows are built once per pixel tile
    prm.BN = BN;
    // more tiles in flight where they are small: the MMA of tile i+3 need not wait for the epilogue of tile i+1
    prm.n_acc = BN <= 128 ? 4 : 2;
    prm.acc_stride = kTmemCols / prm.n_acc;
    prm.n_tiles = (g.K + BN - 1) / BN;
    prm.fd_ntiles = make_fastdiv(prm.n_tiles);
    // halo variant: a weight stage holds the tiles of tap_group taps (all taps, one filter row, or one tap — the largest
    // that still leaves room for 3 stages next to two halo slots); fewer, larger TMA operations keep the single
    // producer thread off the critical path
    prm.fd_pq = make_fastdiv(g.P * g.Q);
    prm.fd_q = make_fastdiv(g.Q);
    prm.fd_wp = make_fastdiv(prm.Wp);
    prm.fd_tpi = make_fastdiv(prm.tiles_per_img);
    prm.tap_group = 1;
    if (halo) {
        const size_t b1 = (size_t)BN * prm.KC;
        const size_t room = kSmemBudget - 1024 - kTailBytes - 2 * (size_t)kMaxWclsBytes - 2 * (size_t)prm.halo_bytes;
        if (3 * b1 * g.R * g.S <= room) prm.tap_group = g.R * g.S;
        else if (2 * b1 * g.S <= room) prm.tap_group = g.S;
    }
    // (measured and not kept: the same decoupled rings for layers with ONE channel tile — deeper weight / A rings behind a
    // smaller fp32 ring — changed nothing, 47.4 k images/s either way; only 256 -> 128 @56x56 moves, 226 -> 198 us, and it
    // does so with three fp32 slots instead of five in either mode: fewer reads in flight leave the DRAM queues to its writes)
    const bool a_stat = fq && !stem && prm.n_tiles > 1 && a_stat_mode() != 0 && a_stat_fits(g) && fq_big;
    prm.a_stat = a_stat ? 1 : 0;
    prm.a_slots = 0;
    const size_t stage_bytes = stem ? (size_t)kBM * prm.KC * prm.cblocks
                                    : (halo ? (size_t)BN * prm.KC * prm.tap_group
                                            : (a_stat ? (size_t)BN * prm.KC : (size_t)(kBM + (pair ? BN / 2 : BN)) * prm.KC));
    const size_t b_res = stem ? (size_t)prm.cblocks * BN * prm.KC : 0;   // fused stem: the weights stay resident, stages hold A only
    // window classes of the output rows / columns (only layers with a spatial kernel have border pixels)
    prm.wcls_smem = 0;
    prm.n_rcls = prm.n_ccls = 0;
    prm.cls_fast = prm.cls_head_r = prm.cls_tail_r = prm.cls_head_c = prm.cls_tail_c = 0;
    if (g.R * g.S > 1) {
        auto classes = [&](int n_out, int in_dim, int taps, uint8_t (*dst)[2]) {
            int n = 0;
            for (int o = 0; o < n_out; ++o) {
                const int h0 = o * g.stride - g.pad;
                const int lo = h0 < 0 ? -h0 : 0, hi = taps < in_dim - h0 ? taps : in_dim - h0;
                bool found = false;
                for (int i = 0; i < n; ++i) found |= dst[i][0] == lo && dst[i][1] == hi;
                if (found) continue;
                if (n == kMaxCls) return -1;
                dst[n][0] = (uint8_t)lo;
                dst[n][1] = (uint8_t)hi;
                ++n;
            }
            return n;
        };
        const int nr = classes(g.P, g.H, g.R, prm.rcls), nc = classes(g.Q, g.W, g.S, prm.ccls);
        if (nr > 0 && nc > 0 && nr * nc * BN * 4 <= kMaxWclsBytes) {
            prm.n_rcls = nr;
            prm.n_ccls = nc;
            prm.wcls_smem = nr * nc * BN * 4;
            // closed form of "which class is output row / column o": verified here against the enumeration
            auto closed = [&](int n_out, int in_dim, int taps, const uint8_t (*cls)[2], int n_cls, int* head, int* tail) {
                *head = (g.pad + g.stride - 1) / g.stride;                        // rows whose window starts above the image
                const int t = in_dim - taps + g.pad;
                *tail = t >= 0 ? t / g.stride + 1 : 0;                            // first row whose window ends below it
                for (int o = 0; o < n_out; ++o) {
                    const int h0 = o * g.stride - g.pad;
                    const int lo = h0 < 0 ? -h0 : 0, hi = taps < in_dim - h0 ? taps : in_dim - h0;
                    const int c = std::min(o, *head) + std::max(0, o - *tail + 1);
                    if (c < 0 || c >= n_cls || cls[c][0] != lo || cls[c][1] != hi) return false;
                }
                return true;
            };
            prm.cls_fast = closed(g.P, g.H, g.R, prm.rcls, nr, &prm.cls_head_r, &prm.cls_tail_r) &&
                           closed(g.Q, g.W, g.S, prm.ccls, nc, &prm.cls_head_c, &prm.cls_tail_c);
        }
    }
    // residual tail: stream the identity through per-warp cp.async rings when 16-byte pieces line up (4 pixels of one
    // image and channel) and two operand stages still fit next to the 96 KB of rings
    // (planes that are not a multiple of 4 pixels — 7x7 — use 4-byte copies: res_async == 2; QB200_RES4=0 turns that off)
    prm.res_async = 0;
    static const bool res4_on = [] {
        const char* e = getenv("QB200_RES4");
        return !(e && e[0] == '0');
    }();
    if (!fq && !halo && !pair && ep.residual != nullptr && ep.out_kind == QB200_OUT_F32 && prm.M < (1ll << 31) &&
        kTailBytes + 2 * (size_t)prm.wcls_smem + kResBytes + 2 * stage_bytes + 1024 <= kSmemBudgetFq) {
        if ((g.P * g.Q) % 4 == 0 && reinterpret_cast<uintptr_t>(ep.residual) % 16 == 0) prm.res_async = 1;
        else if (res4_on && reinterpret_cast<uintptr_t>(ep.residual) % 4 == 0) prm.res_async = 2;
    }
    const size_t tail = kTailBytes + 2 * (size_t)prm.wcls_smem + (prm.res_async ? kResBytes : 0) + (stem ? (size_t)sp.plane_bytes : 0);
    static const size_t fq_budget = [] {      // QB200_FQ_SMEM_KB: shared-memory budget of the fused-quantize kernel (A/B measurements)
        const char* e = getenv("QB200_FQ_SMEM_KB");
        const size_t v = e ? (size_t)atoi(e) * 1024 : kSmemBudgetFq;
        return v < 96 * 1024 ? (size_t)96 * 1024 : (v > kSmemBudgetFq ? kSmemBudgetFq : v);
    }();
    size_t ring_budget = (fq && !stem ? fq_budget : ((fq || prm.res_async) ? kSmemBudgetFq : kSmemBudget)) - 1024 - tail - b_res;
    // ring slot of the fp32 input: a [64 channels][128 pixels] tile, or (fused stem) one box of `stride` rows x all channels
    const size_t xb = stem ? (size_t)((sp.box_bytes + 127) & ~127) : (size_t)kFqKC * kBM * 4;
    if (a_stat) {
        // A ring: two pixel tiles' worth of k-blocks when that leaves three fp32 slots and three weight stages, at least
        // one tile's worth plus a slot of lookahead
        const size_t a_b = (size_t)kBM * prm.KC;
        int as = std::min(2 * prm.cblocks, kMaxStages - kAStatBar);
        while (as > prm.cblocks + 1 && (size_t)as * a_b + (fq_flat ? 4 : 3) * xb + 3 * stage_bytes > ring_budget) --as;
        if (as < prm.cblocks + 1 && prm.cblocks + 1 <= kMaxStages - kAStatBar) as = prm.cblocks + 1;
        as = std::max(as, prm.cblocks);
        QB_REQUIRE((size_t)as * a_b + 2 * xb + 2 * stage_bytes <= ring_budget, QB200_EUNSUPPORTED,
                   "conv_umma: A-stationary tile does not fit shared memory");
        prm.a_slots = as;
        ring_budget -= (size_t)as * a_b;
    }
    if (fq) {
        // fp32 ring: HBM latency x bandwidth needs ~100 KB in flight per SM, so as many 32 KB slots as leave two A/B stages
        // (fused stem: up to kMaxBoxes row boxes in flight — several tiles ahead, the load latency is ~2 tile times)
        static const int fq_xmax = [] {      // QB200_FQ_XMAX: cap of the fp32 ring (A/B measurements: ring depth against stages)
            const char* e = getenv("QB200_FQ_XMAX");
            return e ? atoi(e) : kMaxXStages;
        }();
        int xs = stem ? kMaxBoxes : std::min(kMaxXStages, std::max(fq_xmax, 2));
        // cp.async input: three k-blocks (96 KB) in flight per SM are as good as five (1024 -> 256 @14x14: 66.0 / 68.1 / 69.6 us
        // with 3 / 4 / 5 slots, 512 -> 256 @28x28: 107.7 / 109.3 / 111.8) — the memory system, not the ring, bounds these
        // layers — and the freed shared memory becomes operand stages
        if (x_cp && xs > 3) xs = 3;
        while (xs > (a_stat ? 2 : 3) && (size_t)xs * xb + 2 * stage_bytes > ring_budget) --xs;
        QB_REQUIRE((size_t)xs * xb + 2 * stage_bytes <= ring_budget, QB200_EUNSUPPORTED,
                   "conv_umma: fused-quantize tile does not fit shared memory");
        if (fq_flat && !x_cp && xs < 4) {   // two boxes per k-block need ring depth: back to image-aligned tiles
            fq_flat = false;
            prm.tiles_per_img = (g.P * g.Q + kBM - 1) / kBM;
            prm.m_tiles = g.N * prm.tiles_per_img;
            prm.fd_tpi = make_fastdiv(prm.tiles_per_img);
        }
        prm.x_stages = xs;
        ring_budget -= (size_t)xs * xb;
    }
    if (halo) {
        // halo ring: as many slots (<= 3) as leave at least 4 weight stages
        int hs = 3;
        while (hs > 2 && (size_t)hs * prm.halo_bytes + (prm.tap_group > 1 ? 2 : 4) * stage_bytes > ring_budget) --hs;
        QB_REQUIRE((size_t)hs * prm.halo_bytes + 2 * stage_bytes <= ring_budget, QB200_EUNSUPPORTED,
                   "conv_umma: halo tile does not fit shared memory");
        prm.h_stages = hs;
        ring_budget -= (size_t)hs * prm.halo_bytes;
    }
    int stages = (int)(ring_budget / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (a_stat && stages > kAStatBar) stages = kAStatBar;
    QB_REQUIRE(stages >= 2, QB200_EUNSUPPORTED, "conv_umma: tile does not fit shared memory");
    QB_REQUIRE(!stem || prm.n_tiles == 1, QB200_EUNSUPPORTED, "conv_umma: the fused stem needs a single channel tile");
    prm.stages = stages;
    prm.layout = prm.KC == 128 ? 2u : (prm.KC == 64 ? 4u : 6u);  // SWIZZLE_128B / 64B / 32B
    prm.sbo16 = (uint32_t)(8 * prm.KC) >> 4;
    // instruction descriptor: D=s32, A=u8, B=s8|u8, both K-major, N>>3, M>>4
    prm.idesc = (2u << 4) | (0u << 7) | ((g.w_sign ? 1u : 0u) << 10) | ((uint32_t)(BN >> 3) << 17) |
                ((uint32_t)((pair ? 2 * kBM : kBM) >> 4) << 24);
    prm.err_flag = watchdog_flag();
    prm.x = x_fused;
    prm.q_scale = fq ? aq_fused->scale : nullptr;
    prm.q_zero = fq ? aq_fused->zero : nullptr;
    prm.q_qmin = fq ? aq_fused->qmin : nullptr;
    prm.q_qmax = fq ? aq_fused->qmax : nullptr;

    const CUtensorMapSwizzle swz = prm.KC == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : prm.KC == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                                 : CU_TENSOR_MAP_SWIZZLE_32B;
    alignas(64) CUtensorMap tmap_a, tmap_b;
    if (halo) {
        // zero-padded activations as a matrix [N*Hp*Wp rows][Cp bytes]; box = halo_rows x KC; rows past the end are zero
        cuuint64_t dims[2] = {(cuuint64_t)g.Cp, (cuuint64_t)g.N * prm.Hp * prm.Wp};
        cuuint64_t strides[1] = {(cuuint64_t)g.Cp};
        cuuint32_t box[2] = {(cuuint32_t)prm.KC, (cuuint32_t)prm.halo_rows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = api.tiled(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(qa), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeTiled(halo) failed with CUresult %d", (int)r);
    } else if (!fq && prm.a_tiled) {
        // 1x1 / stride 1 / pad 0: the im2col matrix IS the activation matrix [N*H*W rows][Cp bytes]; tiled boxes of 128 rows
        cuuint64_t dims[2] = {(cuuint64_t)gm.Cp, (cuuint64_t)gm.N * gm.H * gm.W};
        cuuint64_t strides[1] = {(cuuint64_t)gm.Cp};
        cuuint32_t box[2] = {(cuuint32_t)prm.KC, (cuuint32_t)kBM};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = api.tiled(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(qa), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeTiled(1x1 activations) failed with CUresult %d", (int)r);
    } else if (!fq) {
        // activations: (C, W, H, N) u8, im2col mode; base pixel of an output (p,q) is (q*stride - pad, p*stride - pad)
        cuuint64_t dims[4] = {(cuuint64_t)gm.Cp, (cuuint64_t)gm.W, (cuuint64_t)gm.H, (cuuint64_t)gm.N};
        cuuint64_t strides[3] = {(cuuint64_t)gm.Cp, (cuuint64_t)gm.W * gm.Cp, (cuuint64_t)gm.H * gm.W * gm.Cp};
        int lower[2] = {-gm.pad, -gm.pad};
        int upper[2] = {gm.pad - (gm.S - 1), gm.pad - (gm.R - 1)};
        cuuint32_t estr[4] = {1, (cuuint32_t)gm.stride, (cuuint32_t)gm.stride, 1};
        CUresult r = api.im2col(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t*>(qa), dims, strides, lower,
                                upper, (cuuint32_t)prm.KC, (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeIm2col failed with CUresult %d", (int)r);
        // driver <= 13.1 encodes im2col maps of tensors below 128 KiB with a bit that makes the unit fault;
        // the same adjustment NVIDIA's own CUTLASS applies (cute/atom/copy_traits_sm90_im2col.hpp)
        if (api.driver_version <= 13010 && (size_t)gm.N * gm.H * gm.W * gm.Cp < 131072)
            reinterpret_cast<uint64_t*>(&tmap_a)[1] &= ~(1ull << 21);
    }
    if (halo) {
        // tap-major weights (KC, K, cblocks*taps); box = KC x BN x tap_group, rows beyond K zero-filled
        cuuint64_t dims[3] = {(cuuint64_t)prm.KC, (cuuint64_t)g.K, (cuuint64_t)prm.cblocks * g.R * g.S};
        cuuint64_t strides[2] = {(cuuint64_t)prm.KC, (cuuint64_t)prm.KC * g.K};
        cuuint32_t box[3] = {(cuuint32_t)prm.KC, (cuuint32_t)BN, (cuuint32_t)prm.tap_group};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = api.tiled(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(wq), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeTiled(tap-major weights) failed with CUresult %d", (int)r);
    } else {
        // weights: [K rows][R*S*Cp bytes], tiled mode, rows beyond K zero-filled
        cuuint64_t dims[2] = {(cuuint64_t)gm.R * gm.S * gm.Cp, (cuuint64_t)g.K};
        cuuint64_t strides[1] = {(cuuint64_t)gm.R * gm.S * gm.Cp};
        cuuint32_t box[2] = {(cuuint32_t)prm.KC, (cuuint32_t)(pair ? BN / 2 : BN)};   // pair: each CTA stages half of the tile
        cuuint32_t estr[2] = {1, 1};
        CUresult r = api.tiled(&tmap_b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(wq), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    }

    const size_t smem = b_res + (size_t)stages * stage_bytes + (size_t)prm.h_stages * prm.halo_bytes + (fq ? prm.x_stages * xb : 0) +
                        (size_t)prm.a_slots * kBM * prm.KC + 1024 /*align*/ + tail;
    // cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a function: one flag per device (a thread
    // that moves from cuda:0 to cuda:1 must set it again; setting it twice from racing threads is harmless)
    static std::atomic<uint64_t> smem_set_mask{0};
    int cur_dev = 0;
    QB_CUDA(cudaGetDevice(&cur_dev));
    const bool smem_set = cur_dev >= 0 && cur_dev < 64 && ((smem_set_mask.load(std::memory_order_acquire) >> cur_dev) & 1ull);
    if (!smem_set) {
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, false, false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<true, false, false, 1, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, false, false, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        QB_CUDA(cudaFuncSetAttribute(conv_umma_kernel<false, false, true, 1, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudgetFq));
        if (cur_dev >= 0 && cur_dev < 64) smem_set_mask.fetch_or(1ull << cur_dev, std::memory_order_release);
    }
    const int total_tiles = a_stat ? prm.m_tiles : prm.m_tiles * prm.n_tiles;   // (A-stationary: a CTA's unit is a pixel tile)
    const int grid = total_tiles < sms ? total_tiles : sms;
    if (pair) {
        const int pair_tiles = ((prm.m_tiles + 1) / 2) * prm.n_tiles;
        const int pgrid = 2 * (pair_tiles < sms / 2 ? pair_tiles : sms / 2);
        if (ep.q8_out != nullptr)
            QB_CUDA(launch_pdl_cluster(conv_umma_kernel<false, false, true, 1, false, true>, dim3(pgrid), dim3(kThreads), smem, st, 2, tmap_a, tmap_b, prm, out));
        else
            QB_CUDA(launch_pdl_cluster(conv_umma_kernel<false, false, false, 1, false, true>, dim3(pgrid), dim3(kThreads), smem, st, 2, tmap_a, tmap_b, prm, out));
        QB_LAUNCH_CHECK();
        return 0;
    }
    if (stem) {
        // fp32 input as (W, H, C, N); box = whole rows x `stride` rows x all channels of one image, zero fill outside
        cuuint64_t dims[4] = {(cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.C, (cuuint64_t)g.N};
        cuuint64_t strides[3] = {(cuuint64_t)g.W * 4, (cuuint64_t)g.H * g.W * 4, (cuuint64_t)g.C * g.H * g.W * 4};
        cuuint32_t box[4] = {(cuuint32_t)g.W, (cuuint32_t)g.stride, (cuuint32_t)g.C, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = api.tiled(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x_fused), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeTiled(fp32 stem rows) failed with CUresult %d", (int)r);
        QB_CUDA(launch_pdl(conv_umma_kernel<true, false, false, 1, false, false, true>, dim3(grid), dim3(kThreadsFq), smem, st, tmap_a,
                           tmap_b, prm, out));
    } else if (fq) {
        // fp32 input as (pixels, channels, images); box = 128 pixels x 64 channels, no swizzle, zero fill past H*W
        cuuint64_t dims[3] = {(cuuint64_t)g.H * g.W, (cuuint64_t)g.C, (cuuint64_t)g.N};
        cuuint64_t strides[2] = {(cuuint64_t)g.H * g.W * 4, (cuuint64_t)g.C * g.H * g.W * 4};
        cuuint32_t box[3] = {(cuuint32_t)kBM, (cuuint32_t)kFqKC, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = api.tiled(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x_fused), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        QB_REQUIRE(r == CUDA_SUCCESS, QB200_EDRIVER, "cuTensorMapEncodeTiled(fp32 input) failed with CUresult %d", (int)r);
        QB_CUDA(launch_pdl(conv_umma_kernel<true, false, false>, dim3(grid), dim3(kThreadsFq), smem, st, tmap_a, tmap_b, prm, out));
    } else {
        const bool res = ep.residual != nullptr, q8 = ep.q8_out != nullptr;
        if (res && q8)
            QB_CUDA(launch_pdl(conv_umma_kernel<false, true, true>, dim3(grid), dim3(kThreads), smem, st, tmap_a, tmap_b, prm, out));
        else if (res)
            QB_CUDA(launch_pdl(conv_umma_kernel<false, true, false>, dim3(grid), dim3(kThreads), smem, st, tmap_a, tmap_b, prm, out));
        else if (q8 && !ep.store_f32 && prm.n_tiles == 1)   // int8-only output, one channel tile: two epilogue groups
            QB_CUDA(launch_pdl(conv_umma_kernel<false, false, true, 2>, dim3(grid), dim3(64 + 2 * kEpiWarps * 32), smem, st, tmap_a, tmap_b, prm, out));
        else if (q8)
            QB_CUDA(launch_pdl(conv_umma_kernel<false, false, true>, dim3(grid), dim3(kThreads), smem, st, tmap_a, tmap_b, prm, out));
        else if (g.K % BN != 0)   // ragged channel count (MobileNet-style 24 / 96 / 144 ... channels)
            QB_CUDA(launch_pdl(conv_umma_kernel<false, false, false, 1, true>, dim3(grid), dim3(kThreads), smem, st, tmap_a, tmap_b, prm, out));
        else
            QB_CUDA(launch_pdl(conv_umma_kernel<false, false, false>, dim3(grid), dim3(kThreads), smem, st, tmap_a, tmap_b, prm, out));
    }
    QB_LAUNCH_CHECK();
    return 0;
}

}  // namespace qb200
