// float32 GEMM-class steps on the 5th-generation tensor cores (round-2 kernel).
//
//   C[b, m, n] = alpha * sum_k A[b, m, k] * B[b, k, n]       (GemmDev view: groups of strided modes)
//
// 3xTF32 error-compensated split (lo*hi + hi*lo + hi*hi, FP32 accumulation in TMEM) so that float32
// results stay at sgemm-class accuracy (north_star: rtol 1e-5; one-pass TF32 gives 7e-5 per step).
//
// Warp-specialised, persistent over 128 x 128 output tiles, one CTA (13 warps) per SM:
//
//   warps 8-11  loaders + B converters.  They stage BOTH raw operand tiles in shared memory, three
//               k-blocks ahead, with a cursor that runs across tile boundaries: A by cp.async (16 B
//               along whichever of M / K is contiguous; completion counted on an mbarrier with
//               cp.async.mbarrier.arrive.noinc), B by cp.async or by ONE TMA tensor copy per k-block
//               issued by an elected thread (cp.async.bulk.tensor, mbarrier complete_tx) when the
//               operand view has a tensor map.  Then B lo tile = x - trunc(x) (LDS -> STS),
//               fence.proxy.async.
//   warps 0-3   A converters.  Thread r reads row r of the raw A tile, forms lo = x - trunc(x) and
//               writes hi and lo into TENSOR MEMORY (tcgen05.st): the MMAs read the A operand from
//               TMEM (tcgen05.mma [d], [a], b-desc), so A costs the tensor core no shared-memory
//               bandwidth and no lo tile is ever written for it.
//   warp 12     one elected lane issues the 12 tcgen05.mma (128x128x8, kind::tf32) of a k-block;
//               tcgen05.commit frees the stage / publishes the accumulator.  Owns TMEM alloc.
//   warps 4-7   epilogue.  The accumulator is DOUBLE-BUFFERED in TMEM (2 x 128 columns), so the
//               epilogue of tile i (tcgen05.ld, alpha, optional fused consumer, stores, sum|C|)
//               runs while the tensor core works on tile i+1.
//
// TMEM budget (512 columns): 2 accumulators x 128 + 4 A stages x (32 hi + 32 lo).
// Shared memory: 4 x (B_hi 16 KB + B_lo 16 KB) + 4 x A_raw 16 KB + offset tables = 197 KB.
//
// Accumulator truncation.  The tensor core truncates (round toward zero) every time it adds into
// its FP32 accumulator: a systematic relative bias of ~6e-8 per MMA in the chain (1.2e-5 at K = 512).
// Its mean would land in the log-scale, and it differs by a few percent between rows of different
// magnitude -- a difference that persists from step to step and drifts the mantissas of a 510-step
// sweep by 1e-4.  The kernel therefore makes every ROW SUM of every tile exact: the loader warps
// publish the per-k column sums of the raw B tile (folded into their lo-conversion pass), the
// A-converter thread of row m forms sum_k A[m,k] * colsum[k] from the row it holds in registers
// (= the exact sum over the tile's columns of row m; round-to-nearest FP32 FMAs per k-block, the
// k-blocks added up as a pair of floats), the epilogue sums the row as the tensor core produced it
// and rescales the row by the ratio before storing.  Data independent, no model of the hardware's
// rounding.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include <cuda.h>

#include "ctn_internal.h"
#include "device_common.cuh"

namespace ctn {
namespace t32 {

constexpr int BM = 128, BN = 128, BKF = 32;  // BKF floats = 128 bytes = one swizzle row
constexpr int kStages = 4;
constexpr int kRoleThreads = 128;
constexpr int kThreadsT = 3 * kRoleThreads + 3 * 32;  // A converters, epilogue, loaders, MMA warp, TMA issuers (A, B)
constexpr int SUB_BYTES = BN * BKF * 4;           // 16 KB: B_hi or B_lo of one stage
constexpr int STAGE_BYTES = 2 * SUB_BYTES;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t kAccCol0 = 0;       // accumulator buffers: columns [0, 128) and [128, 256)
constexpr uint32_t kAStageCol0 = 256;  // A stages: 64 columns each (hi 32 | lo 32)

// mbarrier indices
constexpr int kFullA = 0, kFullB = kFullA + kStages, kEmpty = kFullB + kStages, kTmaFull = kEmpty + kStages,
              kARaw = kTmaFull + kStages, kARawFree = kARaw + kStages, kAccFull = kARawFree + kStages, kAccEmpty = kAccFull + 2, kSmpFull = kAccEmpty + 2,
              kSmpEmpty = kSmpFull + 2, kBsum = kSmpEmpty + 2, kNumBars = kBsum + kStages;
static_assert(kNumBars <= 40, "barrier block is 40 words");

constexpr int A_OFF = kStages * STAGE_BYTES;  // raw A stages follow the B stages
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)kStages * (STAGE_BYTES + SUB_BYTES) +
                              5 * 128 * sizeof(int64_t) + 2 * 128 * sizeof(double) + 40 * sizeof(uint64_t) + 64 +
                              72 * sizeof(double) + (kStages + 1) * 4 * 32 * sizeof(float);

enum : int32_t { kKFast = 0, kMNFast = 1, kAnyLayout = 2 };  // B staging layouts
enum : int32_t { kFlagCVec = 1, kFlagTmaB = 2, kFlagTmaA = 32, // launch flags
                 kFlagDbgNoSample = 4, kFlagDbgNoLo = 8,       // timing experiments only (CTN_T32_DEBUG): wrong results
                 kFlagDbgNoLoadA = 64, kFlagDbgNoLoadB = 128,  //   "  (TMA operands only: the stage is declared full unloaded)
                 kFlagDbgNoStTmem = 1024,  //   "  (see the MMA issuer / A converters)
                 kFlagTrace = 16 };                            // CTA 0 records a time line of its pipeline events

// time line of CTA 0 (CTN_T32_DEBUG=16): SM clock of event `ev` at k-block / tile ordinal `it`, one
// fire-and-forget store per event into a fixed slot (no atomics: the probe must not stall the warp)
constexpr int kTraceEvents = 32, kTraceIts = 1024;
__device__ long long g_trace[kTraceEvents * kTraceIts];
__device__ __forceinline__ void trace(int flags, int ev, unsigned it) {
  if ((flags & kFlagTrace) && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it < (unsigned)kTraceIts)
    g_trace[ev * kTraceIts + it] = clock64();
}

// wall-clock marks (event row 29, nanoseconds of %globaltimer): [0] CTA 0 enters, [1] CTA 0 leaves, [2] the last CTA
// (whoever it is: [3]) has recorded the scale -- against the launch's duration they show what CTA 0's time line misses
__device__ __forceinline__ void trace_wall(int flags, int slot, long long extra = -1, int row = 29) {
  if (flags & kFlagTrace) {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_trace[row * kTraceIts + slot] = extra >= 0 ? extra : t;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (hi, lo) += x with Knuth's two-sum: the pair carries ~48 bits
__device__ __forceinline__ void two_sum_add(float& hi, float& lo, float x) {
  const float t = hi + x, bb = t - hi;
  lo += (hi - (t - bb)) + (x - bb);
  hi = t;
}
// one lane of the (converged) warp; the compiler knows the guarded region is single-threaded
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// (test_wait never suspends the thread; try_wait may, for a hardware-defined time)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must not hang the GPU box.  After ~2 s (or once another thread of the
// CTA has given up) the wait returns anyway, the result is flagged CTN_RESULT_INTERNAL and the host
// reports an error; `who` identifies the barrier in StateHdr::reserved.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag, StateHdr* hdr, int who) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if ((++spins & 63u) == 0u) {
      if (*abort_flag) return;
      if (clock64() - t0 > 4000000000LL) {
        *abort_flag = 1;
        atomicOr(&hdr->flags, CTN_RESULT_INTERNAL);
        hdr->reserved = (double)(who * 100000 + (int)blockIdx.x * 100 + (int)(threadIdx.x >> 5));
        return;
      }
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// (the TMEM stores carry no memory clobber: they read registers only, and independent shared-memory loads may be
// scheduled past them)
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]));
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// stores through the global window (the output pointer comes out of a table, so the compiler would
// otherwise emit generic stores)
__device__ __forceinline__ void st_global(float* p, float v) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "f"(v) : "memory");
}
__device__ __forceinline__ void st_global4(float* p, float x, float y, float z, float w) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(p)), "f"(x), "f"(y), "f"(z), "f"(w)
               : "memory");
}
__device__ __forceinline__ void bar_named(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kRoleThreads) : "memory"); }

// TMA: one tensor box, global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout).
//  K-major operands : SWIZZLE_128B (layout type 2): rows of 32 floats, 8-row groups SBO apart.
//  MN-major operands: tf32 only supports SWIZZLE_128B_BASE32B (layout type 1, Swizzle<2,5,2>):
//                     atoms of 32 (MN) x 4 (K) floats, MN-atoms LBO apart, K-atoms SBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t type) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) /*version*/ | ((uint64_t)type << 61);
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* gmem, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(gmem), "r"(sz) : "memory");
}
// the mbarrier receives one arrival once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// byte offset of B-producer thread t's i-th chunk inside a 128 x 32 operand tile
template <int LAYOUT>
__device__ __forceinline__ int chunk_offset(int i, int t, int& row, int& kk) {
  if (LAYOUT == kKFast) {
    // K-major, SWIZZLE_128B: 16-byte chunk kc of row `row`
    const int c = i * kRoleThreads + t, kc = c & 7;
    row = c >> 3;
    kk = 4 * kc;
    return (row >> 3) * 1024 + (row & 7) * 128 + ((kc ^ (row & 7)) << 4);
  } else if (LAYOUT == kMNFast) {
    // MN-major, SWIZZLE_128B_BASE32B: 32 x 4 atoms (512 B), MN-atoms 512 B apart, K-atoms
    // 2048 B apart; inside a 128-byte line the 32-byte chunk index is XORed with the line index
    const int c = i * kRoleThreads + t, mc = c & 31;
    kk = c >> 5;
    row = 4 * mc;
    return (mc >> 3) * 512 + (kk >> 2) * 2048 + (kk & 3) * 128 + (((((mc & 7) >> 1) ^ (kk & 3)) << 5) | ((mc & 1) << 4));
  } else {
    // element-wise staging into the K-major layout
    const int e = i * kRoleThreads + t;
    row = e >> 5;
    kk = e & 31;
    return (row >> 3) * 1024 + (row & 7) * 128 + (((kk >> 2) ^ (row & 7)) << 4) + (kk & 3) * 4;
  }
}

// Byte offset of element (column n, k) of the raw B tile.  TMA_MN: the MN-major tile as ONE tensor
// box leaves it ([32-column block][k][32 columns], 128B_ATOM_32B swizzle on the byte address).
template <int LAYOUT, bool TMA>
__device__ __forceinline__ int b_elem_offset(int n, int k) {
  if (LAYOUT == kMNFast) {
    const int mc = n >> 2;
    if (TMA) return (n >> 5) * 4096 + k * 128 + (((((mc & 7) >> 1) ^ (k & 3)) << 5) | ((mc & 1) << 4)) + (n & 3) * 4;
    return (mc >> 3) * 512 + (k >> 2) * 2048 + (k & 3) * 128 + (((((mc & 7) >> 1) ^ (k & 3)) << 5) | ((mc & 1) << 4)) + (n & 3) * 4;
  }
  return (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
}

template <int LAYOUT>
struct ChunkCount { static constexpr int N = (LAYOUT == kAnyLayout) ? 32 : 8; };

__device__ __forceinline__ float residual_tf32(float x) {
  // the tensor core uses trunc(x) as `hi` (it truncates fp32 operands to tf32); the residual is exact
  return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

template <int LAYOUT>
__device__ __forceinline__ void make_lo(const unsigned char* raw_tile, unsigned char* lo_tile,
                                        const int (&soff)[ChunkCount<LAYOUT>::N]) {
  if (LAYOUT == kAnyLayout) {
#pragma unroll
    for (int i = 0; i < ChunkCount<LAYOUT>::N; ++i)
      *reinterpret_cast<float*>(lo_tile + soff[i]) = residual_tf32(*reinterpret_cast<const float*>(raw_tile + soff[i]));
  } else {
    float4 x[ChunkCount<LAYOUT>::N];
#pragma unroll
    for (int i = 0; i < ChunkCount<LAYOUT>::N; ++i) x[i] = *reinterpret_cast<const float4*>(raw_tile + soff[i]);
#pragma unroll
    for (int i = 0; i < ChunkCount<LAYOUT>::N; ++i)
      *reinterpret_cast<float4*>(lo_tile + soff[i]) =
          make_float4(residual_tf32(x[i].x), residual_tf32(x[i].y), residual_tf32(x[i].z), residual_tf32(x[i].w));
  }
}
// TMA-filled tile: thread t converts bytes [t*16 + i*2048, +16) -- the swizzle is a permutation of
// 16-byte chunks, so a linear sweep over the tile covers every element exactly once
__device__ __forceinline__ void make_lo_linear(const unsigned char* raw_tile, unsigned char* lo_tile, int t) {
  float4 x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = *reinterpret_cast<const float4*>(raw_tile + t * 16 + i * 2048);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(lo_tile + t * 16 + i * 2048) =
        make_float4(residual_tf32(x[i].x), residual_tf32(x[i].y), residual_tf32(x[i].z), residual_tf32(x[i].w));
}

// TMA-filled B tile: lo conversion AND the per-k column sums of the raw tile in one pass (the values are in
// registers anyway).  Which (column, k) a thread's chunks hold follows from the tile layout:
//  K-major SWIZZLE_128B (row n = 128 bytes, 16-byte chunk c of row n at position c ^ (n & 7)): warp w owns the
//    two chunks 2w, 2w+1 (k = 8w .. 8w+7) of ALL 128 rows -- lane l takes rows l, l+32, l+64, l+96 -- so that the
//    warp's butterfly (4 + 2 + 1 halving exchanges, then lane ^ 2, lane ^ 1) ends with COMPLETE sums: part[k];
//  MN-major box layout [32-column block][k][32 columns]: chunk i = four columns of k = 16 (i&1) + (t>>3):
//    per-thread sums for its two k, then over the 8 lanes t&7 -> complete sums: part[k].
template <int LAYOUT>
__device__ __forceinline__ void make_lo_linear_sums(const unsigned char* raw_tile, unsigned char* lo_tile, int t,
                                                    float* part /* [32] of this stage */, bool do_lo) {
  const int lane = t & 31, w = t >> 5;
  float4 x[8];
  int off[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    off[i] = (LAYOUT == kMNFast) ? t * 16 + i * 2048
                                 : (lane + 32 * (i & 3)) * 128 + (((2 * w + (i >> 2)) ^ (lane & 7)) << 4);
    x[i] = *reinterpret_cast<const float4*>(raw_tile + off[i]);
  }
  if (do_lo) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      *reinterpret_cast<float4*>(lo_tile + off[i]) =
          make_float4(residual_tf32(x[i].x), residual_tf32(x[i].y), residual_tf32(x[i].z), residual_tf32(x[i].w));
  }
  if (LAYOUT == kMNFast) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      s0 += (x[i].x + x[i].y) + (x[i].z + x[i].w);
      s1 += (x[i + 1].x + x[i + 1].y) + (x[i + 1].z + x[i + 1].w);
    }
#pragma unroll
    for (int m = 1; m <= 4; m <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, m);
      s1 += __shfl_xor_sync(0xffffffffu, s1, m);
    }
    if ((lane & 7) == 0) {
      part[t >> 3] = s0;
      part[16 + (t >> 3)] = s1;
    }
  } else {
    float sv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sv[4 * h] = (x[4 * h].x + x[4 * h + 1].x) + (x[4 * h + 2].x + x[4 * h + 3].x);
      sv[4 * h + 1] = (x[4 * h].y + x[4 * h + 1].y) + (x[4 * h + 2].y + x[4 * h + 3].y);
      sv[4 * h + 2] = (x[4 * h].z + x[4 * h + 1].z) + (x[4 * h + 2].z + x[4 * h + 3].z);
      sv[4 * h + 3] = (x[4 * h].w + x[4 * h + 1].w) + (x[4 * h + 2].w + x[4 * h + 3].w);
    }
#pragma unroll
    for (int hh = 4; hh >= 1; hh >>= 1) {  // exchange with lane ^ 16, ^ 8, ^ 4: keep the half this lane's bit selects
      const bool up = (lane & (4 * hh)) != 0;
#pragma unroll
      for (int j = 0; j < hh; ++j) {
        const float send = up ? sv[j] : sv[j + hh];
        const float keep = up ? sv[j + hh] : sv[j];
        sv[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4 * hh);
      }
    }
    sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 2);
    sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
    if ((lane & 3) == 0) part[8 * w + (lane >> 2)] = sv[0];  // k = 8w + 4 bit4 + 2 bit3 + bit2
  }
}

// Output tiles in bands of 8 row tiles: the CTAs that run at the same time share few distinct
// A row blocks and B column blocks, so both stay in L2 (126 MB) at any problem size.
struct TileCoord {
  int tm, tn, b;
};
__device__ __forceinline__ TileCoord decode_tile(uint32_t w, uint32_t tilesM, uint32_t tilesN) {
  constexpr uint32_t GM = 8;
  const uint32_t per_b = tilesM * tilesN;
  const uint32_t b = w / per_b;
  const uint32_t r = w - b * per_b;
  const uint32_t band = r / (GM * tilesN);
  const uint32_t first = band * GM;
  const uint32_t gsz = min(GM, tilesM - first);
  const uint32_t in = r - band * GM * tilesN;
  TileCoord t;
  t.tn = (int)(in / gsz);
  t.tm = (int)(first + (in - (uint32_t)t.tn * gsz));
  t.b = (int)b;
  return t;
}

__device__ __forceinline__ int64_t mode_offset(const GemmMode* modes, int n, int64_t idx, int which /*0 sa 1 sb 2 sc 3 sw*/) {
  if (n == 1)  // one fused mode (the common case): no division
    return idx * (which == 0 ? modes[0].sa : which == 1 ? modes[0].sb : which == 2 ? modes[0].sc : modes[0].sw);
  int64_t o = 0;
  for (int i = n - 1; i >= 0; --i) {
    const int64_t e = modes[i].extent;
    const int64_t q = idx / e, r = idx - q * e;
    idx = q;
    o += r * (which == 0 ? modes[i].sa : which == 1 ? modes[i].sb : which == 2 ? modes[i].sc : modes[i].sw);
  }
  return o;
}


// several summed modes in A: per-element decomposition, deliberately NOT inlined (it is called from a
// fully unrolled 32-element loop in 24 kernel instantiations)
__device__ __noinline__ int64_t k_offset_a(const GemmDev& g, int64_t k) { return mode_offset(g.k, g.n_k, k, 0); }
__device__ __noinline__ int64_t k_offset_b(const GemmDev& g, int64_t k) { return mode_offset(g.k, g.n_k, k, 1); }

template <int LA, int LB, bool FUSED, bool TMAB, bool TMAA>
__global__ void __launch_bounds__(kThreadsT, 1)
    gemm_tf32x3_kernel(const __grid_constant__ StepDev s, const __grid_constant__ GemmDev g, DevState d,
                       int launch_flags, const __grid_constant__ CUtensorMap tmap_b,
                       const __grid_constant__ CUtensorMap tmap_a) {
  // (the step and its GEMM view travel in the kernel parameters: no dependent descriptor loads from
  // global memory at the head of every launch)
  extern __shared__ unsigned char smem_dyn[];
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_wall(launch_flags, 0);
  if ((threadIdx.x >> 5) == 12) trace(launch_flags, 18, 0);  // SM clock at entry
  if (threadIdx.x == 0) trace_wall(launch_flags, (int)blockIdx.x, -1, 19);  // every CTA: wall clock at entry (row 19) ...
  const uint32_t raw = smem_u32(smem_dyn);
  unsigned char* smem = smem_dyn + ((1024u - (raw & 1023u)) & 1023u);
  unsigned char* tiles = smem;           // B stages: [hi | lo] x kStages
  unsigned char* atiles = smem + A_OFF;  // raw A stages
  int64_t* coltab = reinterpret_cast<int64_t*>(smem + kStages * (STAGE_BYTES + SUB_BYTES));  // [2][128] B producers
  int64_t* rowtab = coltab + 256;                                                             // [2][128] A producers
  int64_t* colC = rowtab + 256;                                                               // [128]    epilogue
  float2* sample = reinterpret_cast<float2*>(colC + 128);                                     // [2][128] exact row sums (hi, lo)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sample + 2 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 4);
  double* red = reinterpret_cast<double*>(tmem_slot + 16);
  float* bsum = reinterpret_cast<float*>(red + 72);  // [kStages][4 warps][32 k]: column sums of the raw B tile
  float* bscr = bsum + kStages * 4 * 32;             // [4 A-converter warps][32 k]: the four partial sums added up
  __shared__ bool last_block;

  const float* A = (const float*)d.ptrs[s.in_slot[0]];
  const float* Bp = (const float*)d.ptrs[s.in_slot[1]];
  float* C = (float*)d.ptrs[s.out_slot];
  const int64_t M = g.M, N = g.N, K = g.K;
  // fused consumer: the 128 accumulator columns of a tile are pr blocks (one per value of the
  // summed index r) of npt consecutive n' each; otherwise 128 consecutive n
  const int pr = FUSED ? g.fuse_pr : 0, npt = FUSED ? BN / pr : BN;
  const uint32_t tilesM = (uint32_t)((M + BM - 1) / BM), tilesN = (uint32_t)((N + npt - 1) / npt);
  const uint32_t total = (uint32_t)g.B * tilesM * tilesN;
  const int nk = (int)((K + BKF - 1) / BKF);
  // this CTA's share of the work as one flat sequence of k-blocks: producers prefetch across tiles
  const uint32_t my_tiles = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  const uint32_t total_it = my_tiles * (uint32_t)nk;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };
  StateHdr* hdr = d.hdr;
#define MBAR_WAIT(idx, parity) mbar_wait(bar(idx), (parity), abort_flag, hdr, (idx))

  // one barrier per thread (a single thread initialising all of them cost ~1900 cycles at the head of every launch)
  if (tid < kNumBars) {
    uint32_t cnt = 4;  // FullA, FullB, ARawFree, Bsum, AccEmpty, SmpFull, SmpEmpty: one elected lane of each of four warps
    if (tid >= kEmpty && tid < kTmaFull) cnt = 5;             // tcgen05.commit + the four A-converter warps (row dots done)
    else if (tid >= kTmaFull && tid < kARaw) cnt = 1;         // TMA: one arrive.expect_tx + the transaction bytes
    else if (tid >= kARaw && tid < kARawFree) cnt = TMAA ? 1 : kRoleThreads;  // TMA transaction / cp.async completion of every loader thread
    else if (tid >= kAccFull && tid < kAccEmpty) cnt = 1;     // tcgen05.commit
    mbar_init(bar(tid), cnt);
    if (tid == 0) *abort_flag = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (tid == 0) trace(launch_flags, 18, 1);  // barriers initialised
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    trace(launch_flags, 18, 2);  // TMEM allocated
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 12) trace(launch_flags, 0, 0);

  // sum|out| and (fused) sum|X| of the epilogue threads, each an unevaluated sum of two floats (FP64 instructions
  // are kept out of the per-tile paths: their latency is what the single warp per scheduler cannot hide)
  float abs_hi = 0.f, abs_lo = 0.f, abst_hi = 0.f, abst_lo = 0.f;

  if (warp < 4) {
    // =========================== A converters ===========================
    // raw A tile (staged by TMA / the loader warps) -> registers -> TMEM (hi, lo); and, from the same registers, the
    // EXACT row sum of the tile: sum_n X[r, n] = sum_k A[r, k] * (sum_n B[k, n]).
    // The row dot of k-block `it` is taken one k-block LATE, together with the conversion of k-block it+1: it needs
    // the B tile's column sums, and waiting for them at once would put the B conversion of the same stage on the A
    // path.  With TMA operands all the waits come first and the two jobs share one basic block, so that their
    // latencies overlap (these four warps set the pace of the whole pipeline once the MMAs are issued properly).
    const int r = tid;
    uint32_t it = 0, tile_iter = 0;
    float pv[32];  // the row of the pending k-block
    bool pend = false, p_last = false;
    uint32_t p_st = 0, p_ph = 0, p_buf = 0, p_ti = 0;
    // the k-blocks are added up as an unevaluated sum of two floats (Knuth's two-sum: ~48 bits): FP64 instructions
    // in this loop cost the converters ~900 cycles per k-block (time line), the float pair costs six FP32 adds
    float row_hi = 0.f, row_lo = 0.f;
    auto add_block = [&](float x) { two_sum_add(row_hi, row_lo, x); };
    auto publish_if_last = [&]() {
      if (p_last) {  // that was the tile's last k-block: hand the exact row sums to the epilogue
        MBAR_WAIT(kSmpEmpty + p_buf, ((p_ti >> 1) & 1u) ^ 1u);
        sample[p_buf * 128 + r] = make_float2(row_hi, row_lo);
        row_hi = row_lo = 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kSmpFull + p_buf));
      }
    };
    // complete column sums (TMA path: the loader warps leave them at [k]): broadcast 16-byte loads
    auto dot_complete = [&]() {
      const float* bp = bsum + p_st * 4 * 32;
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 b4 = *reinterpret_cast<const float4*>(bp + 4 * c);
        p0 = fmaf(pv[4 * c], b4.x, p0);
        p1 = fmaf(pv[4 * c + 1], b4.y, p1);
        p2 = fmaf(pv[4 * c + 2], b4.z, p2);
        p3 = fmaf(pv[4 * c + 3], b4.w, p3);
      }
      add_block((p0 + p1) + (p2 + p3));
    };
    // cp.async path: the column sums are formed here (four partial sums per k, one per converter warp)
    auto finish_pending_cpasync = [&]() {
      // column sums of the raw B tile, per k: thread r takes column r's 32 k-values; a halving butterfly leaves
      // lane l with the warp's sum for k = l
      MBAR_WAIT(kFullB + p_st, p_ph);
      const unsigned char* bt = tiles + p_st * STAGE_BYTES;
      float bv[32];
      if (LB == kMNFast) {
#pragma unroll
        for (int j = 0; j < 32; ++j) bv[j] = *reinterpret_cast<const float*>(bt + b_elem_offset<LB, TMAB>(r, j));
      } else {
        const int base = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = *reinterpret_cast<const float4*>(bt + base + ((c ^ (r & 7)) << 4));
          bv[4 * c] = x.x; bv[4 * c + 1] = x.y; bv[4 * c + 2] = x.z; bv[4 * c + 3] = x.w;
        }
      }
#pragma unroll
      for (int h = 16; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int j = 0; j < h; ++j) {
          const float send = up ? bv[j] : bv[j + h];
          const float keep = up ? bv[j + h] : bv[j];
          bv[j] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
      }
      bsum[(p_st * 4 + warp) * 32 + lane] = bv[0];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBsum + p_st));
      MBAR_WAIT(kBsum + p_st, p_ph);
      if (!(launch_flags & kFlagDbgNoSample)) {
        // lane l adds up the four partial sums of k = l; the warp's scratch row then serves every lane
        const float* bp = bsum + p_st * 4 * 32;
        float* sc = bscr + warp * 32;
        sc[lane] = (bp[lane] + bp[32 + lane]) + (bp[64 + lane] + bp[96 + lane]);
        __syncwarp();
        float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 b4 = *reinterpret_cast<const float4*>(sc + 4 * c);
          p0 = fmaf(pv[4 * c], b4.x, p0);
          p1 = fmaf(pv[4 * c + 1], b4.y, p1);
          p2 = fmaf(pv[4 * c + 2], b4.z, p2);
          p3 = fmaf(pv[4 * c + 3], b4.w, p3);
        }
        add_block((p0 + p1) + (p2 + p3));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kEmpty + p_st));  // done with this stage's B tile and column sums
      publish_if_last();
    };
    for (uint32_t w = blockIdx.x; w < total; w += gridDim.x, ++tile_iter) {
      const uint32_t buf = tile_iter & 1u;
      for (int kb = 0; kb < nk; ++kb, ++it) {
        const uint32_t st = it & (kStages - 1), ph = (it / kStages) & 1u;
        MBAR_WAIT(kARaw + st, ph);  // the raw tile of this k-block has landed ...
        if (warp == 0) trace(launch_flags, 1, it);
        // ... and the MMAs that read this TMEM stage have completed (with TMA for A the issuer has waited for that
        // before loading the tile)
        if (!TMAA) MBAR_WAIT(kEmpty + st, ph ^ 1u);
        const bool dot_now = TMAB && pend;
        if (dot_now) MBAR_WAIT(kBsum + p_st, p_ph);
        if (warp == 0) trace(launch_flags, 2, it);
        tc_fence_after();
        float v[32];
        {
          const unsigned char* at = atiles + st * SUB_BYTES;
          if (LA == kMNFast && TMAA) {  // one TMA box (128 rows, 32 k), no swizzle: [k][row]
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = *reinterpret_cast<const float*>(at + j * 512 + r * 4);
          } else if (LA == kMNFast) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = *reinterpret_cast<const float*>(at + b_elem_offset<kMNFast, false>(r, j));
          } else {
            const int base = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              // chunk c of row r sits at position c ^ (r & 7): eight consecutive rows hit eight bank groups
              const float4 x = *reinterpret_cast<const float4*>(at + base + ((c ^ (r & 7)) << 4));
              v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
            }
          }
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + kAStageCol0 + st * 64u;
        if (!(launch_flags & kFlagDbgNoStTmem)) {
          uint32_t hi[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) hi[j] = __float_as_uint(v[j]);
          tc_st32(taddr, hi);
        }
#pragma unroll
        for (int c = 0; c < 4 && !(launch_flags & kFlagDbgNoStTmem); ++c) {  // eight columns at a time keeps the register peak low
          uint32_t lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) lo[j] = __float_as_uint(residual_tf32(v[8 * c + j]));
          tc_st8(taddr + 32u + 8u * (uint32_t)c, lo);
        }
        // (the pending row dot runs while the TMEM stores are in flight)
        if (dot_now && !(launch_flags & kFlagDbgNoSample)) dot_complete();
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar(kFullA + st));
          mbar_arrive(bar(kARawFree + st));  // this warp has read the raw stage: it may be refilled
          if (dot_now) mbar_arrive(bar(kEmpty + p_st));  // done with the pending stage's column sums
        }
        if (warp == 0) trace(launch_flags, 3, it);
        if (dot_now) publish_if_last();
        if (!TMAB && pend) finish_pending_cpasync();
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = v[j];
        pend = true; p_st = st; p_ph = ph; p_buf = buf; p_ti = tile_iter; p_last = (kb == nk - 1);
      }
    }
    if (pend) {
      if (TMAB) {
        MBAR_WAIT(kBsum + p_st, p_ph);
        if (!(launch_flags & kFlagDbgNoSample)) dot_complete();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kEmpty + p_st));
        publish_if_last();
      } else {
        finish_pending_cpasync();
      }
    }
  } else if (warp < 8) {
    // =========================== epilogue ===========================
    const int q = warp - 4, r = q * 32 + lane, te = tid - kRoleThreads;
    const float alpha0 = (float)(d.inv_kappa[s.in_slot[0]] * d.inv_kappa[s.in_slot[1]]);
    const float* Wp = FUSED ? (const float*)d.ptrs[s.in_slot[2]] : nullptr;
    const float alpha_w = FUSED ? (float)d.inv_kappa[s.in_slot[2]] : 0.f;
    const bool c_vec = (launch_flags & kFlagCVec) != 0;
    uint32_t tile_iter = 0;
    for (uint32_t w = blockIdx.x; w < total; w += gridDim.x, ++tile_iter) {
      const TileCoord tc = decode_tile(w, tilesM, tilesN);
      const int64_t m = (int64_t)tc.tm * BM + r;
      const bool row_ok = m < M;
      int64_t oc = 0, ow = 0;
      if (row_ok) {
        oc = mode_offset(g.m, g.n_m, m, 2);
        if (FUSED) ow = mode_offset(g.m, g.n_m, m, 3);
      }
      if (g.n_batch > 0) oc += mode_offset(g.batch, g.n_batch, tc.b, 2);
      // column offsets: one column mode (the common case) is plain arithmetic; several modes go through a table
      const bool lin = g.n_n == 1;
      const int64_t scn = g.n[0].sc;
      const int ncols = (int)min((int64_t)npt, N - (int64_t)tc.tn * npt);  // valid columns of this tile
      if (!lin) {
        bar_named(2);  // every epilogue thread is done with the previous tile's column table
        int64_t off = 0;
        if (te < npt) {
          const int64_t n = (int64_t)tc.tn * npt + te;
          if (n < N) off = mode_offset(g.n, g.n_n, n, 2);
        }
        colC[te] = off;
        bar_named(2);
      }
      if (q == 0) trace(launch_flags, 17, tile_iter);
      float wv4[4] = {0.f, 0.f, 0.f, 0.f};  // fused consumer: this row's weights (pr <= 4), off the critical path
      if (FUSED && row_ok) {
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
          if (rr < pr) wv4[rr] = __ldg(Wp + ow + (int64_t)rr * g.fuse_sw) * alpha_w;
      }
      const uint32_t buf = tile_iter & 1u;
      MBAR_WAIT(kAccFull + buf, (tile_iter >> 1) & 1u);
      if (q == 0) trace(launch_flags, 10, tile_iter);
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + kAccCol0 + buf * 128u;
      // ---- pass 1: this row's sum over the tile as the tensor core produced it; with the samplers' exact row
      // sum it gives the row's accumulator-truncation factor, applied to every entry of the row below ----
      float rs_hi = 0.f, rs_lo = 0.f, rs_abs = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t a[32];
        tc_ld32(tbase + (uint32_t)cc * 32u, a);
        tc_ld_wait();
        float pa[4] = {0.f, 0.f, 0.f, 0.f}, part_abs = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          pa[j & 3] += __uint_as_float(a[j]);
          part_abs += fabsf(__uint_as_float(a[j]));
        }
        two_sum_add(rs_hi, rs_lo, (pa[0] + pa[1]) + (pa[2] + pa[3]));
        rs_abs += part_abs;
      }
      if (q == 0 && FUSED) trace(launch_flags, 13, tile_iter);
      MBAR_WAIT(kSmpFull + buf, (tile_iter >> 1) & 1u);
      const float2 ex = sample[buf * 128 + r];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kSmpEmpty + buf));
      // rho - 1 = (exact - tc) / tc: the leading parts are within a factor of 1 + 1e-3 of each other, so their
      // difference is exact (Sterbenz) and the quotient needs no more than FP32.  The ratio is well conditioned only
      // when the row does not cancel (same-sign data: |sum| = sum|.|); a row of mixed signs is left alone (its
      // truncation errors are not a common factor anyway)
      float rel = 0.f;
      if (fabsf(rs_hi) > 0.5f * rs_abs && ex.x != 0.f) rel = ((ex.x - rs_hi) + (ex.y - rs_lo)) / rs_hi;
      if (!(fabsf(rel) < 1e-3f)) rel = 0.f;
      const float alpha = fmaf(alpha0, rel, alpha0);
      if (q == 0 && FUSED) trace(launch_flags, 14, tile_iter);
      float lf = 0.f, lt = 0.f;
      float* Crow = C + oc + (lin ? (int64_t)tc.tn * npt * scn : 0);
      // (rows / columns outside the problem were zero-filled on the way in, so their accumulator entries are
      // exact zeros: the sums below need no predicate, only the stores do)
      if (!FUSED) {
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t a[32];
          tc_ld32(tbase + (uint32_t)cc * 32u, a);
          tc_ld_wait();
          if (q == 0) trace(launch_flags, 13 + cc, tile_iter);
          if (cc == 3) {
            // accumulator fully read: the MMA warp may start the tile after next in this buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(kAccEmpty + buf));
          }
          if (row_ok) {
            if (c_vec) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const int c0 = cc * 32 + 4 * j4;
                float* p = lin ? Crow + c0 : Crow + colC[c0];  // (c_vec: unit column stride)
                const float x = __uint_as_float(a[4 * j4]) * alpha, y = __uint_as_float(a[4 * j4 + 1]) * alpha;
                const float z = __uint_as_float(a[4 * j4 + 2]) * alpha, w4 = __uint_as_float(a[4 * j4 + 3]) * alpha;
                if (c0 < ncols) st_global4(p, x, y, z, w4);
                lf += (fabsf(x) + fabsf(y)) + (fabsf(z) + fabsf(w4));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c0 = cc * 32 + j;
                float* p = lin ? Crow + (int64_t)c0 * scn : Crow + colC[c0];
                const float v = __uint_as_float(a[j]) * alpha;
                if (c0 < ncols) st_global(p, v);
                lf += fabsf(v);
              }
            }
          }
        }
      } else {
        // ---- fused consumer: out[m, n'] = sum_r W[m, r] * X[m, (r, n')], X never leaves the SM ----
        const int cpr = npt >> 5;  // 32-column chunks per value of r
#pragma unroll 1
        for (int nc = 0; nc < cpr; ++nc) {
          float o[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = 0.f;
#pragma unroll 1
          for (int rr = 0; rr < pr; ++rr) {
            const int cc = rr * cpr + nc;
            uint32_t a[32];
            tc_ld32(tbase + (uint32_t)cc * 32u, a);
            tc_ld_wait();
            if (nc == cpr - 1 && rr == pr - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(kAccEmpty + buf));
            }
            // rows / columns outside the problem were zero-filled, so their entries are exact zeros
            const float wv = rr == 0 ? wv4[0] : rr == 1 ? wv4[1] : rr == 2 ? wv4[2] : wv4[3];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = __uint_as_float(a[j]) * alpha;  // the un-materialised X entry
              lt += fabsf(t);
              o[j] = fmaf(t, wv, o[j]);
            }
          }
          if (q == 0 && nc == 0) trace(launch_flags, 15, tile_iter);
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c0 = nc * 32 + j;
              float* p = lin ? Crow + (int64_t)c0 * scn : Crow + colC[c0];
              if (c0 < ncols) st_global(p, o[j]);
              lf += fabsf(o[j]);
            }
          }
          if (q == 0 && nc == 0) trace(launch_flags, 16, tile_iter);
        }
      }
      two_sum_add(abs_hi, abs_lo, lf);
      two_sum_add(abst_hi, abst_lo, lt);
      if (q == 0) trace(launch_flags, 12, tile_iter);
    }
    // this warp's share of sum|out| and sum|X| (the only non-zero ones of the CTA)
    const double wa = warp_sum((double)abs_hi + (double)abs_lo), wt = warp_sum((double)abst_hi + (double)abst_lo);
    if (lane == 0) {
      red[2 * q] = wa;
      red[2 * q + 1] = wt;
    }
  } else if (warp < 12) {
    // =========================== loaders (A and B) + B converters ===========================
    const int tb = tid - 2 * kRoleThreads;
    const int n_k = g.n_k;
    const int64_t ska = g.k[0].sa, skb = g.k[0].sb;
    constexpr int NA = ChunkCount<LA>::N, NB = ChunkCount<LB>::N;
    constexpr bool kPtrA = LA != kAnyLayout && !TMAA, kPtrB = LB != kAnyLayout && !TMAB;  // chunk pointers kept in registers
    // issue cursor (runs kStages - 1 k-blocks ahead of the conversion loop, across tiles)
    uint32_t it_i = 0, wi = blockIdx.x, ji = 0;
    int kbi = 0;
    TileCoord tci = decode_tile(total > 0 ? min(wi, total - 1) : 0, tilesM, tilesN);
    const float* abase_i = A;
    const float* bbase_i = Bp;
    bool rows_full_i = false, cols_full_i = false;
    const float* srcA[kPtrA ? NA : 1];
    const float* srcB[kPtrB ? NB : 1];
    auto setup_issue_tile = [&]() {
      tci = decode_tile(wi, tilesM, tilesN);
      int64_t* rt = rowtab + (ji & 1u) * 128;
      int64_t* ct = coltab + (ji & 1u) * 128;
      if (!TMAA) {
        const int64_t m = (int64_t)tci.tm * BM + tb;
        rt[tb] = m < M ? mode_offset(g.m, g.n_m, m, 0) : -1;
      }
      if (!TMAB) {
        const int rr = FUSED ? tb / npt : 0;
        const int64_t n = FUSED ? (int64_t)tci.tn * npt + (tb - rr * npt) : (int64_t)tci.tn * BN + tb;
        int64_t ob = -1;
        if (n < N) {
          ob = mode_offset(g.n, g.n_n, n, 1);
          if (FUSED) ob += (int64_t)rr * g.fuse_sb;
        }
        ct[tb] = ob;
      }
      abase_i = A + (g.n_batch > 0 ? mode_offset(g.batch, g.n_batch, tci.b, 0) : 0);
      bbase_i = Bp + (g.n_batch > 0 ? mode_offset(g.batch, g.n_batch, tci.b, 1) : 0);
      rows_full_i = ((int64_t)tci.tm + 1) * BM <= M;
      cols_full_i = ((int64_t)tci.tn + 1) * npt <= N;
      if (!(TMAA && TMAB)) bar_named(1);  // tables complete (their other halves may still be in use for the previous tile)
      if (kPtrA) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          int row, kk;
          chunk_offset<LA>(i, tb, row, kk);
          const int64_t ro = rt[row];
          srcA[kPtrA ? i : 0] = ro >= 0 ? abase_i + ro + (n_k == 1 ? (int64_t)kk * ska : 0) : nullptr;
        }
      }
      if (kPtrB) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          int row, kk;
          chunk_offset<LB>(i, tb, row, kk);
          const int64_t ro = ct[row];
          srcB[kPtrB ? i : 0] = ro >= 0 ? bbase_i + ro + (n_k == 1 ? (int64_t)kk * skb : 0) : nullptr;
        }
      }
    };
    auto issue_next = [&]() {
      if (it_i >= total_it) return;
      const uint32_t st = it_i & (kStages - 1), ph = (it_i / kStages) & 1u;
      const int64_t k0 = (int64_t)kbi * BKF;
      const bool k_in = k0 + BKF <= K && n_k == 1;
      // ---- A: raw tile, 16-byte cp.async; completion is counted on the stage's mbarrier ----
      if (TMAA) {
        // (warp 13 issues the TMA boxes of A)
      } else {
      MBAR_WAIT(kARawFree + st, ph ^ 1u);  // the four converter warps have read the previous contents
      if (warp == 8) trace(launch_flags, 21, it_i);
      {
        const uint32_t stage = smem_u32(atiles + st * SUB_BYTES);
        const int64_t* rt = rowtab + (ji & 1u) * 128;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          int row, kk;
          const int soff = chunk_offset<LA>(i, tb, row, kk);
          if (kPtrA && n_k == 1) {
            const float* base = srcA[kPtrA ? i : 0];
            const bool valid = base != nullptr && (k_in || k0 + kk < K);
            cp_async16(stage + soff, valid ? base + k0 * ska : A, valid);
          } else {
            const int64_t ro = rt[row];
            const int64_t k = k0 + kk;
            const bool valid = ro >= 0 && k < K;
            const float* src = valid ? abase_i + ro + (n_k == 1 ? k * ska : k_offset_a(g, k)) : A;
            if (LA == kAnyLayout) cp_async4(stage + soff, src, valid);
            else cp_async16(stage + soff, src, valid);
          }
        }
        if (warp == 8) trace(launch_flags, 22, it_i);
        cp_async_mbar_arrive(bar(kARaw + st));
      }
      }
      if (warp == 8) trace(launch_flags, 23, it_i);
      // ---- B: one TMA tensor box issued by an elected thread, or cp.async chunks ----
      if (TMAB) {
        // (warp 14 issues the TMA boxes of B)
      } else {
        MBAR_WAIT(kEmpty + st, ph ^ 1u);
        const uint32_t stage = smem_u32(tiles + st * STAGE_BYTES);
        const int64_t* ct = coltab + (ji & 1u) * 128;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          int row, kk;
          const int soff = chunk_offset<LB>(i, tb, row, kk);
          if (kPtrB && n_k == 1) {
            const float* base = srcB[kPtrB ? i : 0];
            const bool valid = base != nullptr && (k_in || k0 + kk < K);
            cp_async16(stage + soff, valid ? base + k0 * skb : Bp, valid);
          } else {
            const int64_t ro = ct[row];
            const int64_t k = k0 + kk;
            const bool valid = ro >= 0 && k < K;
            const float* src = valid ? bbase_i + ro + (n_k == 1 ? k * skb : k_offset_b(g, k)) : Bp;
            if (LB == kAnyLayout) cp_async4(stage + soff, src, valid);
            else cp_async16(stage + soff, src, valid);
          }
        }
      }
      if (warp == 8 && !TMAB) trace(launch_flags, 4, it_i);
      ++it_i;
      if (++kbi == nk) {
        kbi = 0;
        wi += gridDim.x;
        ++ji;
        if (it_i < total_it) setup_issue_tile();
      }
    };
    if (total_it > 0) setup_issue_tile();
    if (warp == 8) trace(launch_flags, 20, 0);
#pragma unroll 1
    for (int n = 0; n < kStages - 1; ++n) {
      issue_next();
      cp_async_commit();
    }
    int soffB[NB];
    if (!TMAB) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        int row, kk;
        soffB[i] = chunk_offset<LB>(i, tb, row, kk);
      }
    }
    for (uint32_t it = 0; it < total_it; ++it) {
      const uint32_t st = it & (kStages - 1), ph = (it / kStages) & 1u;
      unsigned char* stage = tiles + st * STAGE_BYTES;
      if (TMAB) {
        MBAR_WAIT(kTmaFull + st, ph);
        if (warp == 8) trace(launch_flags, 5, it);
        make_lo_linear_sums<LB>(stage, stage + SUB_BYTES, tb, bsum + st * 4 * 32, !(launch_flags & kFlagDbgNoLo));
      } else {
        cp_async_wait<kStages - 2>();  // this thread's chunks of k-block `it` have landed
        if (warp == 8) trace(launch_flags, 5, it);
        if (!(launch_flags & kFlagDbgNoLo)) make_lo<LB>(stage, stage + SUB_BYTES, soffB);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(kFullB + st));
        if (TMAB) mbar_arrive(bar(kBsum + st));  // (the column sums were written before the __syncwarp)
      }
      if (warp == 8) trace(launch_flags, 6, it);
      issue_next();
      cp_async_commit();
    }
    cp_async_wait<0>();
  } else if (warp == 12) {
    // =========================== MMA issuer (warp 12) ===========================
    // instruction descriptor (cute::UMMA::InstrDescriptor): f32 accumulate, tf32 x tf32, A (TMEM) K-major
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((LB == kMNFast ? 1u : 0u) << 16) |
                           ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    // B descriptors: K-major SWIZZLE_128B; MN-major SWIZZLE_128B_BASE32B with the atom strides of whoever
    // filled the tile (cp.async: MN-atoms 512 B / K-atoms 2048 B; TMA box: 4096 B / 512 B)
    const uint32_t b_lbo = (LB == kMNFast) ? (TMAB ? 4096u : 512u) : 16u;
    const uint32_t b_sbo = (LB == kMNFast) ? (TMAB ? 512u : 2048u) : 1024u;
    const uint32_t b_ty = (LB == kMNFast) ? 1u : 2u;
    // The WHOLE warp runs the loop (warp-uniform control flow, uniform registers for the descriptors) and one
    // lane chosen by elect.sync issues: issuing from a divergent `if (lane == 0)` region makes the compiler wrap
    // every tcgen05 instruction in a serialising loop over the possibly-active lanes, which cost ~120 cycles per
    // MMA -- more than the MMA itself (measured with the time line: 360 cycles per k-step whether it held one
    // MMA or three).
    const uint64_t desc0 = make_desc(0u, b_lbo, b_sbo, b_ty);  // + (shared address >> 4): stays inside the 14-bit field
    {
      uint32_t it = 0, tile_iter = 0;
      for (uint32_t w = blockIdx.x; w < total; w += gridDim.x, ++tile_iter) {
        const uint32_t buf = tile_iter & 1u;
        MBAR_WAIT(kAccEmpty + buf, ((tile_iter >> 1) & 1u) ^ 1u);  // the epilogue has drained this buffer
        const uint32_t d_tmem = tmem_base + kAccCol0 + buf * 128u;
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t st = it & (kStages - 1), ph = (it / kStages) & 1u;
          MBAR_WAIT(kFullA + st, ph);
          trace(launch_flags, 7, it);
          MBAR_WAIT(kFullB + st, ph);
          trace(launch_flags, 8, it);
          tc_fence_after();
          const uint32_t a_col = tmem_base + kAStageCol0 + st * 64u;
          const uint64_t d_hi = desc0 + (uint64_t)(smem_u32(tiles + st * STAGE_BYTES) >> 4), d_lo = d_hi + (uint64_t)(SUB_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BKF / 8; ++ks) {
              // one UMMA_K = 8 floats: 32 bytes along K (K-major) or two k-atoms (MN-major)
              const uint32_t bo = (LB == kMNFast) ? (uint32_t)ks * 2u * b_sbo : (uint32_t)ks * 32u;
              const uint64_t dbh = d_hi + (uint64_t)(bo >> 4), dbl = d_lo + (uint64_t)(bo >> 4);
              const uint32_t a_hi = a_col + (uint32_t)ks * 8u, a_lo = a_col + 32u + (uint32_t)ks * 8u;
              tc_mma_ts(d_tmem, a_lo, dbh, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
              tc_mma_ts(d_tmem, a_hi, dbl, idesc, 1u);
              tc_mma_ts(d_tmem, a_hi, dbh, idesc, 1u);
            }
            tc_commit(bar(kEmpty + st));  // stage free once these MMAs have read it
            if (kb == nk - 1) tc_commit(bar(kAccFull + buf));
          }
          __syncwarp();
          trace(launch_flags, 9, it);
        }
      }
    }
  } else if (warp == 13) {
    // =========================== TMA issuer, A (warp 13) ===========================
    // ONE tensor box per k-block (K-fast: (32 k, 128 rows) SWIZZLE_128B; M-fast: (128 rows, 32 k)).  A thread of
    // its own: it blocks on "stage free" without holding up anybody's conversions, so all four stages stay in flight.
    // (warp-uniform loop, one elected lane issues: see the MMA issuer)
    if (TMAA) {
      uint32_t it = 0;
      for (uint32_t w = blockIdx.x; w < total; w += gridDim.x) {
        const TileCoord tc = decode_tile(w, tilesM, tilesN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t st = it & (kStages - 1), ph = (it / kStages) & 1u;
          MBAR_WAIT(kARawFree + st, ph ^ 1u);  // the four converter warps have read the previous contents ...
          MBAR_WAIT(kEmpty + st, ph ^ 1u);     // ... and the MMAs that read the TMEM stage this tile is converted
                                               // into have completed (the converters rely on that)
          trace(launch_flags, 21, it);
          if (elect_one()) {
            if (launch_flags & kFlagDbgNoLoadA) {
              mbar_arrive(bar(kARaw + st));
            } else {
              mbar_expect_tx(bar(kARaw + st), (uint32_t)SUB_BYTES);
              const uint32_t dst = smem_u32(atiles + st * SUB_BYTES);
              if (LA == kMNFast) tma_load_4d(dst, &tmap_a, bar(kARaw + st), tc.tm * BM, kb * BKF, tc.b, 0);
              else tma_load_4d(dst, &tmap_a, bar(kARaw + st), kb * BKF, tc.tm * BM, tc.b, 0);
            }
          }
          __syncwarp();
          trace(launch_flags, 23, it);
        }
      }
    }
  } else if (warp == 14) {
    // =========================== TMA issuer, B (warp 14) ===========================
    if (TMAB) {
      uint32_t it = 0;
      for (uint32_t w = blockIdx.x; w < total; w += gridDim.x) {
        const TileCoord tc = decode_tile(w, tilesM, tilesN);
        for (int kb = 0; kb < nk; ++kb, ++it) {
          const uint32_t st = it & (kStages - 1), ph = (it / kStages) & 1u;
          MBAR_WAIT(kEmpty + st, ph ^ 1u);  // tensor core and row dots have finished with this stage
          trace(launch_flags, 24, it);
          if (elect_one()) {
            if (launch_flags & kFlagDbgNoLoadB) {
              mbar_arrive(bar(kTmaFull + st));
            } else {
              mbar_expect_tx(bar(kTmaFull + st), (uint32_t)SUB_BYTES);
              const uint32_t dst = smem_u32(tiles + st * STAGE_BYTES);
              if (LB == kMNFast)  // dims: (32 n-inner, k, n-outer [32-column blocks], r or batch)
                tma_load_4d(dst, &tmap_b, bar(kTmaFull + st), 0, kb * BKF, tc.tn * (npt >> 5), FUSED ? 0 : tc.b);
              else                // dims: (k, n, batch, 1)
                tma_load_4d(dst, &tmap_b, bar(kTmaFull + st), kb * BKF, tc.tn * BN, tc.b, 0);
            }
          }
          __syncwarp();
          trace(launch_flags, 4, it);
        }
      }
    }
  }
#undef MBAR_WAIT

  // ---- teardown + fused stabilize epilogue (sum|C| -> kappa, log-scale, measured truncation bias) ----
  tc_fence_before();
  __syncthreads();
  if (warp == 12) trace(launch_flags, 30, 0);
  if (warp == 12) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
  if (warp == 12) trace(launch_flags, 27, 0);
  if (tid == 0) trace_wall(launch_flags, (int)blockIdx.x, -1, 25);  // ... and when its roles are done (row 25)
  if (tid == 0) {  // (the four epilogue warps left their sums in `red` before the barrier above)
    d.partials[blockIdx.x] = (red[0] + red[2]) + (red[4] + red[6]);
    d.partials[kMaxPartials + blockIdx.x] = (red[1] + red[3]) + (red[5] + red[7]);
    __threadfence();
    const uint32_t tk = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (warp == 12) trace(launch_flags, 28, 0);
  if (last_block) {
    __threadfence();
    double a0 = 0.0, a1 = 0.0;
    for (int i = tid; i < (int)gridDim.x; i += blockDim.x) {
      a0 += __ldcg(&d.partials[i]);
      a1 += __ldcg(&d.partials[kMaxPartials + i]);
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      red[8 + 2 * warp] = a0;
      red[9 + 2 * warp] = a1;
    }
    __syncthreads();
    double tot[2] = {0.0, 0.0};
    if (tid == 0) {
      const int nw = min(kThreadsT / 32, ((int)gridDim.x + 31) / 32);  // (the other warps hold zeros)
      for (int i = 0; i < nw; ++i) {
        tot[0] += red[8 + 2 * i];
        tot[1] += red[9 + 2 * i];
      }
    }
    const double norm = tot[0], norm_x = tot[1];
    if (tid == 0) {
      if (FUSED) {
        // the reference stabilises X, then out = X_hat * W: both of its decisions must be "rescale"
        // for the fused result to be the same up to the gauge (einsum.py:102)
        const double kappa_x = norm_x / ((double)s.out_numel * g.fuse_pr);
        if (!(norm_x > kMinNorm) || !(norm / kappa_x > kMinNorm)) d.hdr->flags |= CTN_RESULT_DEGENERATE;
      }
      record_scale(d, s.out_slot, norm, s.out_numel);
      d.tickets[s.ticket] = 0;
      trace_wall(launch_flags, 2);
      trace_wall(launch_flags, 3, (long long)blockIdx.x);
    }
  }
  if (warp == 12) trace(launch_flags, 31, 0);
  if (blockIdx.x == 0 && threadIdx.x == 0) trace_wall(launch_flags, 1);
}

static int pick_layout4(const GemmMode* rows, int n_rows, const GemmMode* ks, int n_k, bool use_b) {
  auto stride = [&](const GemmMode& m) { return use_b ? m.sb : m.sa; };
  auto others_ok = [&](const GemmMode* skip) {
    for (int i = 0; i < n_rows; ++i)
      if (&rows[i] != skip && (stride(rows[i]) & 3)) return false;
    for (int i = 0; i < n_k; ++i)
      if (&ks[i] != skip && (stride(ks[i]) & 3)) return false;
    return true;
  };
  if (n_k > 0) {
    const GemmMode& in = ks[n_k - 1];
    if (stride(in) == 1 && (in.extent % 4 == 0) && others_ok(&in)) return kKFast;
  }
  if (n_rows > 0) {
    const GemmMode& in = rows[n_rows - 1];
    if (stride(in) == 1 && (in.extent % 4 == 0) && others_ok(&in)) return kMNFast;
  }
  return kAnyLayout;
}

// ---------------------------------------------------------------------------
// TMA tensor maps for the B operand (driver entry point fetched at run time: the library must
// load on machines without libcuda, e.g. the CPU-only build check)
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      p = nullptr;
    }
    return (EncodeTiledFn)p;
  }();
  return fn;
}

// Shape of the B view as a <= 4-D tensor map, or false.  dims/strides innermost first (elements).
struct TmaShape {
  uint64_t dim[4];
  uint64_t stride[4];  // stride[0] == 1 implied
  uint32_t box[4];
  int swizzle;         // 0 none, 1 SWIZZLE_128B, 2 SWIZZLE_128B_ATOM_32B
};

static bool tma_shape_ok(TmaShape& t) {
  for (int i = 1; i < 4; ++i) {
    if (t.stride[i] == 0) {
      if (t.dim[i] != 1) return false;  // broadcast dimension
      t.stride[i] = 4;
    }
    if (t.stride[i] % 4 != 0) return false;  // byte strides must be multiples of 16
    if ((t.stride[i] * 4) >> 40) return false;
  }
  for (int i = 0; i < 4; ++i)
    if (t.dim[i] == 0 || t.dim[i] > (1ull << 32) || t.box[i] == 0 || t.box[i] > 256) return false;
  return true;
}

static bool tma_shape_b(const GemmDev& g, int lb, TmaShape& t) {
  if (g.n_k != 1 || g.K % 4 != 0) return false;
  const bool fused = g.fuse_pr > 0;
  const int npt = fused ? BN / g.fuse_pr : BN;
  if (lb == kMNFast) {
    // innermost N mode contiguous; a tile's npt columns must sit inside it: extent % npt == 0
    if (g.n_n != 1 || g.n[0].sb != 1 || g.n[0].extent % npt != 0 || npt % 32 != 0) return false;
    if (g.n_batch > 1 || (fused && g.n_batch != 0)) return false;
    t.swizzle = 2;
    t.dim[0] = 32; t.stride[0] = 1; t.box[0] = 32;
    t.dim[1] = (uint64_t)g.K; t.stride[1] = (uint64_t)g.k[0].sb; t.box[1] = BKF;
    t.dim[2] = (uint64_t)(g.n[0].extent / 32); t.stride[2] = 32; t.box[2] = (uint32_t)(npt / 32);
    if (fused) {
      t.dim[3] = (uint64_t)g.fuse_pr; t.stride[3] = (uint64_t)g.fuse_sb; t.box[3] = (uint32_t)g.fuse_pr;
    } else {
      t.dim[3] = g.n_batch ? (uint64_t)g.batch[0].extent : 1;
      t.stride[3] = g.n_batch ? (uint64_t)g.batch[0].sb : (uint64_t)(g.K * g.k[0].sb > 0 ? 4 : 4);
      t.box[3] = 1;
    }
  } else if (lb == kKFast) {
    if (fused || g.n_n != 1 || g.k[0].sb != 1 || g.n_batch > 1) return false;
    t.swizzle = 1;
    t.dim[0] = (uint64_t)g.K; t.stride[0] = 1; t.box[0] = BKF;
    t.dim[1] = (uint64_t)g.n[0].extent; t.stride[1] = (uint64_t)g.n[0].sb; t.box[1] = BN;
    t.dim[2] = g.n_batch ? (uint64_t)g.batch[0].extent : 1;
    t.stride[2] = g.n_batch ? (uint64_t)g.batch[0].sb : 4;
    t.box[2] = 1;
    t.dim[3] = 1; t.stride[3] = 4; t.box[3] = 1;
  } else {
    return false;
  }
  return tma_shape_ok(t);
}

// The A view (rows = the M group) as a tensor map: raw tile only, read back row-wise by the
// converter warps, so the layout just has to be conflict-free for them.
static bool tma_shape_a(const GemmDev& g, int la, TmaShape& t) {
  if (g.n_k != 1 || g.n_m != 1 || g.n_batch > 1) return false;
  if (la == kKFast) {
    if (g.k[0].sa != 1) return false;
    t.swizzle = 1;  // (32 k, 128 rows): the layout the cp.async path produces
    t.dim[0] = (uint64_t)g.K; t.stride[0] = 1; t.box[0] = BKF;
    t.dim[1] = (uint64_t)g.M; t.stride[1] = (uint64_t)g.m[0].sa; t.box[1] = BM;
  } else if (la == kMNFast) {
    if (g.m[0].sa != 1) return false;
    t.swizzle = 0;  // (128 rows, 32 k) -> [k][row], 512-byte lines
    t.dim[0] = (uint64_t)g.M; t.stride[0] = 1; t.box[0] = BM;
    t.dim[1] = (uint64_t)g.K; t.stride[1] = (uint64_t)g.k[0].sa; t.box[1] = BKF;
  } else {
    return false;
  }
  t.dim[2] = g.n_batch ? (uint64_t)g.batch[0].extent : 1;
  t.stride[2] = g.n_batch ? (uint64_t)g.batch[0].sa : 4;
  t.box[2] = 1;
  t.dim[3] = 1; t.stride[3] = 4; t.box[3] = 1;
  return tma_shape_ok(t);
}

struct TmapKey {
  const void* base;
  uint64_t dim[4], stride[4];
  uint32_t box[4];
  int swizzle;
  bool operator<(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) < 0; }
};

static bool get_tensor_map(const TmaShape& t, const void* base, CUtensorMap& out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  static std::mutex mu;
  static std::map<TmapKey, CUtensorMap> cache;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base;
  for (int i = 0; i < 4; ++i) { key.dim[i] = t.dim[i]; key.stride[i] = t.stride[i]; key.box[i] = t.box[i]; }
  key.swizzle = t.swizzle;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { out = it->second; return true; }
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) { dims[i] = t.dim[i]; box[i] = t.box[i]; }
  for (int i = 1; i < 4; ++i) strides[i - 1] = t.stride[i] * 4;
  CUtensorMap m;
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  t.swizzle == 2   ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                  : t.swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                   : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 8192) cache.clear();
  cache[key] = m;
  out = m;
  return true;
}

template <int LA, int LB, bool FUSED, bool TMAB, bool TMAA>
static cudaError_t launch_one(const ExecCtx& c, const Launch& l, const GemmDev& g, int flags, const CUtensorMap& tmap_b,
                              const CUtensorMap& tmap_a) {
  const Plan& p = *c.plan;
  auto kern = gemm_tf32x3_kernel<LA, LB, FUSED, TMAB, TMAA>;
  static AttrOnce attr;  // per instantiation
  {
    cudaError_t e = attr.ensure(kern, p.device, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const int64_t npt = g.fuse_pr ? BN / g.fuse_pr : BN;  // output columns covered by one tile
  const int64_t total = g.B * ((g.M + BM - 1) / BM) * ((g.N + npt - 1) / npt);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(total, p.sm_count));
  kern<<<grid, kThreadsT, SMEM_BYTES, c.stream>>>(p.steps[l.first_step], g, dev_state(c), flags, tmap_b, tmap_a);
  return cudaGetLastError();
}

// TMA variants exist for the 16-byte layouts only: B alone, or A and B (kAnyLayout never carries the flags)
template <int LA, int LB, bool FUSED>
static cudaError_t launch_tma(const ExecCtx& c, const Launch& l, const GemmDev& g, int flags, const CUtensorMap& tmap_b,
                              const CUtensorMap& tmap_a) {
  constexpr int LBt = LB == kAnyLayout ? kKFast : LB, LAt = LA == kAnyLayout ? kKFast : LA;
  if (LB != kAnyLayout && (flags & kFlagTmaB)) {
    if (LA != kAnyLayout && (flags & kFlagTmaA)) return launch_one<LAt, LBt, FUSED, true, true>(c, l, g, flags, tmap_b, tmap_a);
    return launch_one<LA, LBt, FUSED, true, false>(c, l, g, flags, tmap_b, tmap_a);
  }
  return launch_one<LA, LB, FUSED, false, false>(c, l, g, flags, tmap_b, tmap_a);
}

template <int LB>
static cudaError_t launch_modes(const ExecCtx& c, const Launch& l, const GemmDev& g, int la, int flags,
                                const CUtensorMap& tmap, const CUtensorMap& tmap_a) {
  const bool fused = g.fuse_pr > 0;
#define CTN_T32_LA(LA_)                                                                                  \
  if (la == LA_)                                                                                         \
    return fused ? launch_tma<LA_, LB, true>(c, l, g, flags, tmap, tmap_a) : launch_tma<LA_, LB, false>(c, l, g, flags, tmap, tmap_a);
  CTN_T32_LA(kKFast)
  CTN_T32_LA(kMNFast)
  CTN_T32_LA(kAnyLayout)
#undef CTN_T32_LA
  return cudaErrorInvalidValue;
}

}  // namespace t32

// Debug aid (not part of the documented ABI): copy out and clear CTA 0's event time line
// ([event][ordinal] SM clocks, 0 = not recorded).
namespace t32 { static int g_trace_launch = 0; }  // gemm launches since the time line was last read
extern "C" int ctn_debug_t32_trace(long long* host_out, int cap) {
  t32::g_trace_launch = 0;
  const int n = std::min(cap, t32::kTraceEvents * t32::kTraceIts);
  if (cudaMemcpyFromSymbol(host_out, t32::g_trace, (size_t)n * sizeof(long long)) != cudaSuccess) return -1;
  static long long zeros[t32::kTraceEvents * t32::kTraceIts];
  cudaMemcpyToSymbol(t32::g_trace, zeros, sizeof(zeros));
  return n;
}

bool gemm_tf32_supported(const Plan& p, const StepDev& s, const GemmDev& g) {
  if (p.dtype != CTN_F32) return false;
  if (s.n_in != 2 || g.K < 8 || g.n_k > 4) return false;
  // tensor-core tiles are 128 x 128: leave skinny / tiny products to the generic kernel
  if (g.M < 16 || g.N < 16) return false;
  const int64_t npt = g.fuse_pr ? t32::BN / g.fuse_pr : t32::BN;
  const double tiles = (double)g.B * (double)((g.M + t32::BM - 1) / t32::BM) * (double)((g.N + npt - 1) / npt);
  if (tiles >= 2.0e9) return false;  // 32-bit tile counters
  return true;
}

// cfg 8 = this kernel.  la / lb are the staging layouts of A and B; fuse_pad bit 0 says the output
// can be written with 16-byte stores (subject to the pointer's alignment at launch).
void gemm_tf32_configure(const Plan&, GemmDev& g) {
  using namespace t32;
  g.cfg = 8;
  g.splits = 1;
  bool batch_ok_a = true, batch_ok_b = true, c_vec = g.fuse_pr == 0 && g.n_n > 0;
  for (int i = 0; i < g.n_batch; ++i) {
    batch_ok_a = batch_ok_a && !(g.batch[i].sa & 3);
    batch_ok_b = batch_ok_b && !(g.batch[i].sb & 3);
    c_vec = c_vec && !(g.batch[i].sc & 3);
  }
  for (int i = 0; i < g.n_m; ++i) c_vec = c_vec && !(g.m[i].sc & 3);
  for (int i = 0; i < g.n_n; ++i) {
    if (i == g.n_n - 1) c_vec = c_vec && g.n[i].sc == 1 && (g.n[i].extent % 4 == 0);
    else c_vec = c_vec && !(g.n[i].sc & 3);
  }
  g.la = batch_ok_a ? pick_layout4(g.m, g.n_m, g.k, g.n_k, false) : kAnyLayout;
  g.lb = batch_ok_b ? pick_layout4(g.n, g.n_n, g.k, g.n_k, true) : kAnyLayout;
  g.fuse_pad = c_vec ? 1 : 0;
  g.scratch_off = 0;
}

cudaError_t launch_gemm_tf32(const ExecCtx& c, const Launch& l, bool a_aligned, bool b_aligned, bool c_aligned,
                             const void* b_ptr, const void* a_ptr) {
  using namespace t32;
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  const GemmDev& g = p.gemms[s.gemm_off];
  const int la = a_aligned ? g.la : kAnyLayout, lb = b_aligned ? g.lb : kAnyLayout;
  int flags = (g.fuse_pad & 1) && c_aligned ? kFlagCVec : 0;
  CUtensorMap tmap, tmap_a;
  memset(&tmap, 0, sizeof(tmap));
  memset(&tmap_a, 0, sizeof(tmap_a));
  static const bool no_tma = getenv("CTN_TF32_NO_TMA") != nullptr;
  TmaShape shape;
  if (!no_tma && b_aligned && tma_shape_b(g, lb, shape) && get_tensor_map(shape, b_ptr, tmap)) {
    flags |= kFlagTmaB;
    static const bool no_tma_a = getenv("CTN_TF32_NO_TMA_A") != nullptr;
    if (!no_tma_a && a_aligned && tma_shape_a(g, la, shape) && get_tensor_map(shape, a_ptr, tmap_a)) flags |= kFlagTmaA;
  }
  static const int dbg = getenv("CTN_T32_DEBUG") ? atoi(getenv("CTN_T32_DEBUG")) &
                                                       (kFlagDbgNoSample | kFlagDbgNoLo | kFlagTrace | kFlagDbgNoLoadA | kFlagDbgNoLoadB | kFlagDbgNoStTmem)
                                                 : 0;
  flags |= dbg;
  // CTN_T32_TRACE_LAUNCH=n: only the n-th launch after the time line was last read records (default: every launch
  // does, the last one's records survive)
  static const int trace_launch = getenv("CTN_T32_TRACE_LAUNCH") ? atoi(getenv("CTN_T32_TRACE_LAUNCH")) : -1;
  if (trace_launch >= 0 && g_trace_launch++ != trace_launch) flags &= ~kFlagTrace;
  switch (lb) {
    case kKFast: return launch_modes<kKFast>(c, l, g, la, flags, tmap, tmap_a);
    case kMNFast: return launch_modes<kMNFast>(c, l, g, la, flags, tmap, tmap_a);
    default: return launch_modes<kAnyLayout>(c, l, g, la, flags, tmap, tmap_a);
  }
}

}  // namespace ctn
