// "Skinny" pairwise steps: one huge operand against a small one,
//
//     C[r, n] = alpha * sum_k A[r, k] * B[k, n],   rows r ~ 10^5..10^8,  K*N <= 4096
//
// the shape of every boundary-absorbs-a-site step of a PEPS sweep (cfg 5: K, N in
// {8, 64}) and of the short-K steps of MPS sweeps.  These steps sit at or below
// the FP64 ridge (a few FLOP/byte), so the job is to stream the big operand
// through the SMs exactly once: no shared-memory tiles and no per-tile barriers;
// the small operand lives in shared memory (broadcast reads), one thread owns
// one output row with all its N accumulators in registers, row offsets of the
// strided views are decomposed once per row.  Same fused stabilize epilogue as
// the other kernels (alpha = pending 1/kappa, per-CTA sum|C|, last CTA -> kappa).
#include <algorithm>
#include <cstdlib>

#include "ctn_internal.h"
#include "device_common.cuh"

namespace ctn {

constexpr int kSkinnyThreads = 256;
constexpr int kSkinnyMaxK = 256;
constexpr int kSkinnyDefaultRtWide = 2;  // 8-row tiles per warp for 40..64 columns: 128 registers, two CTAs per SM (cfg 5: 32.6 -> 29.2 ms against 4 tiles, 222 registers, one CTA per SM)
constexpr int kSkinnyMaxKN = 4096;  // doubles of shared memory for the small operand

__device__ __forceinline__ void sk_offsets(const GemmMode* modes, int n, int64_t idx, bool use_b, int64_t& ox,
                                           int64_t& oc) {
  ox = 0;
  oc = 0;
  for (int i = n - 1; i >= 0; --i) {
    const int64_t e = modes[i].extent;
    const int64_t q = idx / e, r = idx - q * e;
    idx = q;
    ox += r * (use_b ? modes[i].sb : modes[i].sa);
    oc += r * modes[i].sc;
  }
}

// SWAP = false: rows come from operand 0 (groups m), the small operand is operand 1 (groups n)
// SWAP = true : rows come from operand 1 (groups n), the small operand is operand 0 (groups m)
// Work item = RPT consecutive rows x one block of NT columns; a CTA's threads are split into
// `col_blocks` groups that take the same rows and different column blocks (the big operand's
// row is then re-read by the sibling group from L1, not from HBM).  Every value of the small
// operand read from shared memory feeds RPT FMAs, which keeps the loop on the FP64 pipe.
template <typename T, int NT, int RPT, bool SWAP>
__global__ void __launch_bounds__(kSkinnyThreads) skinny_kernel(const StepDev* __restrict__ steps,
                                                                const GemmDev* __restrict__ gemms, int step_idx,
                                                                DevState d) {
  __shared__ double small_s[kSkinnyMaxKN];     // [K][n_pad], n_pad = col_blocks * NT
  __shared__ int64_t koff_s[kSkinnyMaxK];      // offset of summed index k in the big operand
  __shared__ int64_t ccol_s[64];               // offset of small-side index n in C
  __shared__ double red[33];
  __shared__ bool last_block;
  __shared__ int64_t crow_w[kSkinnyThreads];   // per warp: C row offsets of the rows being stored (-1: dead)
  extern __shared__ double stage_s[];          // per warp: 32 x (NT + 1) doubles, for coalesced stores
  const StepDev& s = steps[step_idx];
  const GemmDev& g = gemms[s.gemm_off];
  const T* big = (const T*)d.ptrs[s.in_slot[SWAP ? 1 : 0]];
  const T* small_g = (const T*)d.ptrs[s.in_slot[SWAP ? 0 : 1]];
  T* C = (T*)d.ptrs[s.out_slot];
  const double alpha = d.inv_kappa[s.in_slot[0]] * d.inv_kappa[s.in_slot[1]];
  const GemmMode* row_modes = SWAP ? g.n : g.m;
  const GemmMode* col_modes = SWAP ? g.m : g.n;
  const int n_row_modes = SWAP ? g.n_n : g.n_m, n_col_modes = SWAP ? g.n_m : g.n_n;
  const int64_t R = SWAP ? g.N : g.M;
  const int N = (int)(SWAP ? g.M : g.N);
  const int K = (int)g.K;
  const int col_blocks = (N + NT - 1) / NT;
  const int n_pad = col_blocks * NT;

  // stage the small operand [K][n_pad] (zero padded), the K offsets and the C column offsets
  for (int n = threadIdx.x; n < n_pad; n += blockDim.x) {
    int64_t ox = 0, oc = 0;
    if (n < N) sk_offsets(col_modes, n_col_modes, n, !SWAP, ox, oc);
    ccol_s[n] = oc;
  }
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    int64_t ob, dummy;
    sk_offsets(g.k, g.n_k, k, SWAP, ob, dummy);  // big operand's K offset
    koff_s[k] = ob;
  }
  for (int e = threadIdx.x; e < K * n_pad; e += blockDim.x) {
    const int k = e / n_pad, n = e - k * n_pad;
    double v = 0.0;
    if (n < N) {
      int64_t on, ok, dummy;
      sk_offsets(col_modes, n_col_modes, n, !SWAP, on, dummy);
      sk_offsets(g.k, g.n_k, k, !SWAP, ok, dummy);
      v = (double)__ldg(small_g + on + ok);
    }
    small_s[e] = v;
  }
  __syncthreads();

  const int group_threads = kSkinnyThreads / col_blocks;       // col_blocks in {1, 2}
  const int cb = threadIdx.x / group_threads, tg = threadIdx.x - cb * group_threads;
  const int64_t rows_per_cta = (int64_t)group_threads * RPT;
  const double* bcol = small_s + cb * NT;
  double local_abs = 0.0;
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_cta; r0 < R; r0 += (int64_t)gridDim.x * rows_per_cta) {
    // rows r0 + tg + j * group_threads: consecutive threads take consecutive rows (coalescing)
    const T* arow[RPT];
    int64_t crow[RPT];
    bool live[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int64_t r = r0 + tg + (int64_t)j * group_threads;
      live[j] = r < R;
      int64_t oa = 0, oc = 0;
      if (live[j]) sk_offsets(row_modes, n_row_modes, r, SWAP, oa, oc);
      arow[j] = big + oa;
      crow[j] = oc;
    }
    double acc[RPT][NT];
#pragma unroll
    for (int j = 0; j < RPT; ++j)
#pragma unroll
      for (int n = 0; n < NT; ++n) acc[j][n] = 0.0;
    int k = 0;
    for (; k + 2 <= K; k += 2) {
      double a0[RPT], a1[RPT];
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        a0[j] = live[j] ? (double)__ldg(arow[j] + koff_s[k]) : 0.0;
        a1[j] = live[j] ? (double)__ldg(arow[j] + koff_s[k + 1]) : 0.0;
      }
      const double* b0 = bcol + k * n_pad;
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const double x0 = b0[n], x1 = b0[n_pad + n];
#pragma unroll
        for (int j = 0; j < RPT; ++j) acc[j][n] = fma(a1[j], x1, fma(a0[j], x0, acc[j][n]));
      }
    }
    for (; k < K; ++k) {
      const double* b0 = bcol + k * n_pad;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const double a0 = live[j] ? (double)__ldg(arow[j] + koff_s[k]) : 0.0;
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[j][n] = fma(a0, b0[n], acc[j][n]);
      }
    }
    // Stores: a thread owns whole rows, so writing them directly would scatter 8-byte
    // pieces over 32 rows per instruction (one L2 transaction per element).  Transpose
    // each warp's 32 x NT block through shared memory and write 32 consecutive
    // elements per instruction instead.
    {
      const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
      double* tile = stage_s + (threadIdx.x >> 5) * (32 * (NT + 1));
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        __syncwarp();
        crow_w[threadIdx.x] = live[j] ? crow[j] : -1;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const T v = (T)(acc[j][n] * alpha);
          tile[lane * (NT + 1) + n] = (double)v;
          if (live[j] && cb * NT + n < N) local_abs += fabs((double)v);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const int idx = i * 32 + lane, row = idx / NT, col = idx - row * NT;
          const int64_t ro = crow_w[wbase + row];
          if (ro >= 0 && cb * NT + col < N) C[ro + ccol_s[cb * NT + col]] = (T)tile[row * (NT + 1) + col];
        }
      }
      __syncwarp();
    }
  }

  const double bsum = block_sum(local_abs, red);
  if (threadIdx.x == 0) {
    d.partials[blockIdx.x] = bsum;
    __threadfence();
    const uint32_t tk = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (last_block) {
    __threadfence();
    double a2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a2 += __ldcg(&d.partials[i]);
    const double norm = block_sum(a2, red);
    if (threadIdx.x == 0) {
      record_scale(d, s.out_slot, norm, s.out_numel);
      d.tickets[s.ticket] = 0;
    }
  }
}

// ---------------------------------------------------------------------------
// float64 variant on the FP64 tensor pipe: the big operand's 8x4 DMMA fragments are loaded
// straight from global memory (a lane reads A[row = lane/4][k = lane%4]: whole 32-byte sectors
// when the summed index is contiguous, coalesced rows when a free index is), the small operand's
// fragments come from shared memory, a warp owns 32 rows x NT8*8 columns of accumulators and
// writes them as 16-byte pairs -- no shared-memory tiles, no barriers in the main loop.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void sk_dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// RT = 8-row tiles per warp: 4 (32 rows, 222 registers at 64 columns: one CTA per SM) or 2 (16 rows: two CTAs per
// SM, twice the warps to hide the latency of the direct global loads of the big operand)
template <int NT8, bool SWAP, int RT>
__global__ void __launch_bounds__(kSkinnyThreads) skinny_dmma_kernel(const StepDev* __restrict__ steps,
                                                                     const GemmDev* __restrict__ gemms, int step_idx,
                                                                     DevState d) {
  constexpr int NP = NT8 * 8, LDB = NP + 4;   // LDB = 4 or 12 (mod 16): conflict-free fragment reads
  extern __shared__ double sk_smem[];          // [K_pad][LDB] small operand
  __shared__ int64_t koff_s[kSkinnyMaxK];      // offset of summed index k in the big operand (-1: padding)
  __shared__ int64_t skof_s[kSkinnyMaxK];      // ... and in the small operand
  __shared__ int64_t ccol_s[64];
  __shared__ int64_t scol_s[64];               // offset of small-side index n in the small operand
  __shared__ double red[33];
  __shared__ bool last_block;
  const StepDev& s = steps[step_idx];
  const GemmDev& g = gemms[s.gemm_off];
  const double* big = (const double*)d.ptrs[s.in_slot[SWAP ? 1 : 0]];
  const double* small_g = (const double*)d.ptrs[s.in_slot[SWAP ? 0 : 1]];
  double* C = (double*)d.ptrs[s.out_slot];
  const double alpha = d.inv_kappa[s.in_slot[0]] * d.inv_kappa[s.in_slot[1]];
  const GemmMode* row_modes = SWAP ? g.n : g.m;
  const GemmMode* col_modes = SWAP ? g.m : g.n;
  const int n_row_modes = SWAP ? g.n_n : g.n_m, n_col_modes = SWAP ? g.n_m : g.n_n;
  const int64_t R = SWAP ? g.N : g.M;
  const int N = (int)(SWAP ? g.M : g.N);
  const int K = (int)g.K, K_pad = (K + 3) & ~3;

  // the index decompositions (64-bit divisions) are done once per column and once per k, not once per
  // element of the small operand: the prologue was ~13 us per CTA, a third of a launch on 2.6e5 rows
  for (int n = threadIdx.x; n < NP; n += blockDim.x) {
    int64_t ox = 0, oc = 0;
    if (n < N) sk_offsets(col_modes, n_col_modes, n, !SWAP, ox, oc);
    ccol_s[n] = n < N ? oc : -1;
    scol_s[n] = ox;
  }
  for (int k = threadIdx.x; k < K_pad; k += blockDim.x) {
    int64_t ob = -1, os = 0, dummy;
    if (k < K) {
      sk_offsets(g.k, g.n_k, k, SWAP, ob, dummy);
      sk_offsets(g.k, g.n_k, k, !SWAP, os, dummy);
    }
    koff_s[k] = ob;
    skof_s[k] = os;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < K_pad * NP; e += blockDim.x) {
    const int k = e / NP, n = e - k * NP;
    sk_smem[k * LDB + n] = (n < N && k < K) ? __ldg(small_g + scol_s[n] + skof_s[k]) : 0.0;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int frow = lane >> 2, fk = lane & 3;
  constexpr int WR = RT * 8;  // rows per warp item
  const int64_t rows_per_cta = (kSkinnyThreads / 32) * WR;
  double local_abs = 0.0;
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_cta + warp * WR; r0 < R; r0 += (int64_t)gridDim.x * rows_per_cta) {
    const double* arow[RT];
    int64_t crow[RT];
    bool live[RT];
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      const int64_t r = r0 + t * 8 + frow;
      live[t] = r < R;
      int64_t oa = 0, oc = 0;
      if (live[t]) sk_offsets(row_modes, n_row_modes, r, SWAP, oa, oc);
      arow[t] = big + oa;
      crow[t] = oc;
    }
    double acc[RT][NT8][2];
#pragma unroll
    for (int t = 0; t < RT; ++t)
#pragma unroll
      for (int j = 0; j < NT8; ++j) acc[t][j][0] = acc[t][j][1] = 0.0;
#pragma unroll 4
    for (int ks = 0; ks < K_pad; ks += 4) {
      const int64_t ko = koff_s[ks + fk];
      double a[RT];
#pragma unroll
      for (int t = 0; t < RT; ++t) a[t] = (ko >= 0 && live[t]) ? __ldg(arow[t] + ko) : 0.0;
      const double* brow = sk_smem + (ks + fk) * LDB + frow;
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const double b = brow[j * 8];
#pragma unroll
        for (int t = 0; t < RT; ++t) sk_dmma(acc[t][j], a[t], b);
      }
    }
    // C fragment: row = lane/4, columns 2*(lane%4) + {0, 1} of each 8x8 tile
#pragma unroll
    for (int t = 0; t < RT; ++t) {
      if (!live[t]) continue;
      double* out_row = C + crow[t];
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const int c0 = j * 8 + fk * 2;
        const int64_t o0 = ccol_s[c0], o1 = ccol_s[c0 + 1];
        const double v0 = acc[t][j][0] * alpha, v1 = acc[t][j][1] * alpha;
        if (o0 >= 0 && o1 == o0 + 1 && (((uintptr_t)(out_row + o0)) & 15) == 0) {
          *reinterpret_cast<double2*>(out_row + o0) = make_double2(v0, v1);
          local_abs += fabs(v0) + fabs(v1);
        } else {
          if (o0 >= 0) { out_row[o0] = v0; local_abs += fabs(v0); }
          if (o1 >= 0) { out_row[o1] = v1; local_abs += fabs(v1); }
        }
      }
    }
  }

  const double bsum = block_sum(local_abs, red);
  if (threadIdx.x == 0) {
    d.partials[blockIdx.x] = bsum;
    __threadfence();
    const uint32_t tk = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (last_block) {
    __threadfence();
    double a2 = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) a2 += __ldcg(&d.partials[i]);
    const double norm = block_sum(a2, red);
    if (threadIdx.x == 0) {
      record_scale(d, s.out_slot, norm, s.out_numel);
      d.tickets[s.ticket] = 0;
    }
  }
}

template <bool SWAP>
static cudaError_t launch_dmma(const ExecCtx& c, const Launch& l, int64_t rows, int cols, int K) {
  const Plan& p = *c.plan;
  DevState d = dev_state(c);
  const int nt8 = (cols + 7) / 8;
  const int np = nt8 * 8, ldb = np + 4, k_pad = (K + 3) & ~3;
  const size_t smem = (size_t)k_pad * ldb * sizeof(double);
  // CTN_SKINNY_RT (experiment knob): 8-row tiles per warp, 2 or 4
  static const int rt_env = getenv("CTN_SKINNY_RT") ? atoi(getenv("CTN_SKINNY_RT")) : 0;
  const int rt = rt_env == 2 || rt_env == 4 ? rt_env : (nt8 >= 5 ? kSkinnyDefaultRtWide : 4);
  const int64_t per_cta = (kSkinnyThreads / 32) * rt * 8;
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024));
    if (e != cudaSuccess) return e;
    // one resident wave: a CTA that waits for an SM would run the prologue (small operand -> shared memory)
    // a second time on it.  The 64-column kernel with 32 rows per warp holds 222 registers: ONE CTA per SM
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSkinnyThreads, smem);
    if (e != cudaSuccess) return e;
    const int64_t wave = (int64_t)p.sm_count * std::max(1, std::min(occ, 4));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((rows + per_cta - 1) / per_cta,
                                                                 std::min<int64_t>(wave, kMaxPartials)));
    kern<<<grid, kSkinnyThreads, smem, c.stream>>>(p.d_steps, p.d_gemms, l.first_step, d);
    return cudaGetLastError();
  };
  if (rt == 2) {
    switch (nt8) {
      case 1: return launch(skinny_dmma_kernel<1, SWAP, 2>);
      case 2: return launch(skinny_dmma_kernel<2, SWAP, 2>);
      case 3: return launch(skinny_dmma_kernel<3, SWAP, 2>);
      case 4: return launch(skinny_dmma_kernel<4, SWAP, 2>);
      case 5: return launch(skinny_dmma_kernel<5, SWAP, 2>);
      case 6: return launch(skinny_dmma_kernel<6, SWAP, 2>);
      case 7: return launch(skinny_dmma_kernel<7, SWAP, 2>);
      default: return launch(skinny_dmma_kernel<8, SWAP, 2>);
    }
  }
  switch (nt8) {
    case 1: return launch(skinny_dmma_kernel<1, SWAP, 4>);
    case 2: return launch(skinny_dmma_kernel<2, SWAP, 4>);
    case 3: return launch(skinny_dmma_kernel<3, SWAP, 4>);
    case 4: return launch(skinny_dmma_kernel<4, SWAP, 4>);
    case 5: return launch(skinny_dmma_kernel<5, SWAP, 4>);
    case 6: return launch(skinny_dmma_kernel<6, SWAP, 4>);
    case 7: return launch(skinny_dmma_kernel<7, SWAP, 4>);
    default: return launch(skinny_dmma_kernel<8, SWAP, 4>);
  }
}

// column block width and rows per thread for a small side of `n` columns
static int pick_nt(int64_t n) { return n <= 8 ? 8 : n <= 16 ? 16 : 32; }   // 33..64 columns: two blocks of 32
static int padded_cols(int64_t n) { const int nt = pick_nt(n); return (int)((n + nt - 1) / nt) * nt; }

// Returns 0 (not eligible), 3 (rows from operand 0) or 4 (rows from operand 1).
int skinny_config(const Plan& p, const StepDev& s, const GemmDev& g) {
  if (s.n_in != 2 || g.n_batch != 0 || g.K < 1 || g.K > kSkinnyMaxK) return 0;
  // EXPERIMENT knob: products with K * cols above this go to the tiled tensor-core kernels
  static const int64_t max_kn = getenv("CTN_SKINNY_MAX_KN") ? atoll(getenv("CTN_SKINNY_MAX_KN")) : kSkinnyMaxKN;
  if (p.dtype == CTN_F64 && g.K * std::min(g.M, g.N) > max_kn) return 0;
  auto ok = [&](int64_t rows, int64_t cols) {
    return cols <= 64 && rows >= 2048 && rows >= 16 * cols && g.K * padded_cols(cols) <= kSkinnyMaxKN;
  };
  if (ok(g.M, g.N)) return 3;
  if (ok(g.N, g.M)) return 4;
  return 0;
}

template <typename T, bool SWAP>
static cudaError_t launch_nt(const ExecCtx& c, const Launch& l, int nt, int64_t rows, int col_blocks) {
  const Plan& p = *c.plan;
  DevState d = dev_state(c);
  const int64_t cap = std::min<int64_t>((int64_t)p.sm_count * 4, kMaxPartials);
  auto grid_for = [&](int rpt) {
    const int64_t per_cta = (int64_t)(kSkinnyThreads / col_blocks) * rpt;
    return (int)std::max<int64_t>(1, std::min<int64_t>((rows + per_cta - 1) / per_cta, cap));
  };
  auto launch = [&](auto kern, int rpt, int nt_) -> cudaError_t {
    const size_t smem = (size_t)(kSkinnyThreads / 32) * 32 * (nt_ + 1) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_for(rpt), kSkinnyThreads, smem, c.stream>>>(p.d_steps, p.d_gemms, l.first_step, d);
    return cudaGetLastError();
  };
  switch (nt) {
    case 8: return launch(skinny_kernel<T, 8, 4, SWAP>, 4, 8);
    case 16: return launch(skinny_kernel<T, 16, 2, SWAP>, 2, 16);
    default: return launch(skinny_kernel<T, 32, 2, SWAP>, 2, 32);
  }
}

cudaError_t launch_skinny(const ExecCtx& c, const Launch& l) {
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  const GemmDev& g = p.gemms[s.gemm_off];
  const bool swap = g.cfg == 4;
  const int64_t rows = swap ? g.N : g.M, cols = swap ? g.M : g.N;
  const int nt = pick_nt(cols);
  const int col_blocks = (int)((cols + nt - 1) / nt);
  if (p.dtype == CTN_F64 && !getenv("CTN_SKINNY_NO_DMMA"))
    return swap ? launch_dmma<true>(c, l, rows, (int)cols, (int)g.K) : launch_dmma<false>(c, l, rows, (int)cols, (int)g.K);
  if (p.dtype == CTN_F32)
    return swap ? launch_nt<float, true>(c, l, nt, rows, col_blocks) : launch_nt<float, false>(c, l, nt, rows, col_blocks);
  return swap ? launch_nt<double, true>(c, l, nt, rows, col_blocks) : launch_nt<double, false>(c, l, nt, rows, col_blocks);
}

}  // namespace ctn
