// Permute-free strided batched GEMM for pairwise steps with a summed index
// (opt_einsum classes GEMM / TDOT / DOT / GEMV): float64 on the FP64 tensor
// pipe (DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no f64 kind).
//
//   C[b, m, n] = alpha * sum_k A[b, m, k] * B[b, k, n]
//
// b, m, n, k are *groups* of tensor modes with arbitrary element strides: the
// operands are read where they lie (no transposes are materialised) and C is
// written directly in the order the step's output subscript asks for.  Row /
// column offsets of a tile are decomposed once per tile into shared memory, K
// offsets on the fly.  cp.async multi-stage pipeline (16-byte copies along
// whichever of M/N or K is contiguous in memory, 8-byte otherwise), fragments
// from padded, bank-conflict-free shared tiles, fused epilogue (alpha = pending
// 1/kappa of the inputs, per-CTA sum|C|, last CTA records kappa and ln kappa:
// stabilize, contractn/einsum.py:89-107).  Persistent grid, deterministic
// two-pass split-K for small outputs with long K.
#include <algorithm>

#include "ctn_internal.h"
#include "device_common.cuh"

namespace ctn {

constexpr int kGemmThreads = 256;
constexpr int BK = 16;
constexpr int kStages = 3;

enum : int32_t { kKFast = 0, kMNFast = 1, kAnyLayout = 2 };

template <int WARPS_M, int WARPS_N, int TM, int TN>
struct TileCfg {
  static constexpr int kWarpsM = WARPS_M, kWarpsN = WARPS_N, kTM = TM, kTN = TN;
  static constexpr int BM = WARPS_M * TM * 8;
  static constexpr int BN = WARPS_N * TN * 8;
  static constexpr int A_ELEMS = (BM * (BK + 4) > BK * (BM + 4)) ? BM * (BK + 4) : BK * (BM + 4);
  static constexpr int B_ELEMS = (BN * (BK + 4) > BK * (BN + 4)) ? BN * (BK + 4) : BK * (BN + 4);
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static constexpr size_t SMEM = (size_t)kStages * STAGE_ELEMS * sizeof(double) + (size_t)2 * (BM + BN) * sizeof(int64_t);
};
using Cfg128x128 = TileCfg<2, 4, 8, 4>;
using Cfg128x64 = TileCfg<4, 2, 4, 4>;   // 90 KB of shared memory, <= 128 registers: two CTAs per SM
using Cfg64x128 = TileCfg<2, 4, 4, 4>;
using Cfg128x32 = TileCfg<8, 1, 2, 4>;
using Cfg32x128 = TileCfg<1, 8, 4, 2>;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// offsets of a flattened group index into two tensors (x: operand, c: output)
__device__ __forceinline__ void group_offsets(const GemmMode* modes, int n, int64_t idx, bool use_b, int64_t& ox,
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

__device__ __forceinline__ int64_t k_offset(const GemmMode* modes, int n, int64_t k, bool use_b) {
  if (n == 1) return k * (use_b ? modes[0].sb : modes[0].sa);
  int64_t o = 0;
  for (int i = n - 1; i >= 0; --i) {
    const int64_t e = modes[i].extent;
    const int64_t q = k / e, r = k - q * e;
    k = q;
    o += r * (use_b ? modes[i].sb : modes[i].sa);
  }
  return o;
}

// Copy one BK-deep slab of an operand tile into shared memory.
//  LAYOUT kKFast : smem [row][k], ld = BK+4, 16-byte copies along k
//  LAYOUT kMNFast: smem [k][row], ld = ROWS+4, 16-byte copies along rows
//  LAYOUT kAnyLayout: smem [row][k], 8-byte copies
template <int ROWS, int LAYOUT>
__device__ __forceinline__ void load_slab(double* tile, const double* base, const int64_t* row_off,
                                          const GemmMode* kmodes, int n_k, bool use_b, int64_t k0, int64_t k_end) {
  const int tid = threadIdx.x;
  if (LAYOUT == kKFast) {
    constexpr int LD = BK + 4, CPR = BK / 2, CHUNKS = ROWS * CPR;
#pragma unroll
    for (int c0 = 0; c0 < CHUNKS; c0 += kGemmThreads) {
      const int c = c0 + tid;
      if (CHUNKS % kGemmThreads != 0 && c >= CHUNKS) break;
      const int row = c / CPR, kc = c % CPR;
      const int64_t k = k0 + 2 * kc;
      const int64_t ro = row_off[row];
      const bool valid = ro >= 0 && k < k_end;
      const double* src = valid ? base + ro + k_offset(kmodes, n_k, k, use_b) : base;
      cp_async16(tile + row * LD + 2 * kc, src, valid);
    }
  } else if (LAYOUT == kMNFast) {
    constexpr int LD = ROWS + 4, CPR = ROWS / 2, CHUNKS = BK * CPR;
#pragma unroll
    for (int c0 = 0; c0 < CHUNKS; c0 += kGemmThreads) {
      const int c = c0 + tid;
      if (CHUNKS % kGemmThreads != 0 && c >= CHUNKS) break;
      const int kk = c / CPR, rc = c % CPR;
      const int64_t k = k0 + kk;
      const int64_t ro = row_off[2 * rc];
      const bool valid = ro >= 0 && k < k_end;
      const double* src = valid ? base + ro + k_offset(kmodes, n_k, k, use_b) : base;
      cp_async16(tile + kk * LD + 2 * rc, src, valid);
    }
  } else {
    constexpr int LD = BK + 4, ELEMS = ROWS * BK;
#pragma unroll
    for (int e0 = 0; e0 < ELEMS; e0 += kGemmThreads) {
      const int e = e0 + tid;
      if (ELEMS % kGemmThreads != 0 && e >= ELEMS) break;
      const int row = e / BK, kk = e % BK;
      const int64_t k = k0 + kk;
      const int64_t ro = row_off[row];
      const bool valid = ro >= 0 && k < k_end;
      const double* src = valid ? base + ro + k_offset(kmodes, n_k, k, use_b) : base;
      cp_async8(tile + row * LD + kk, src, valid);
    }
  }
}

template <int ROWS, int LAYOUT>
__device__ __forceinline__ double frag(const double* tile, int row, int k) {
  if (LAYOUT == kMNFast) return tile[k * (ROWS + 4) + row];
  return tile[row * (BK + 4) + k];
}

template <class Cfg, int LA, int LB>
__global__ void __launch_bounds__(kGemmThreads, (Cfg::BM * Cfg::BN <= 128 * 64) ? 2 : 1)
    gemm_f64_kernel(const StepDev* __restrict__ steps, const GemmDev* __restrict__ gemms, int step_idx, DevState d) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, TM = Cfg::kTM, TN = Cfg::kTN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tiles = reinterpret_cast<double*>(smem_raw);
  int64_t* rowA = reinterpret_cast<int64_t*>(tiles + kStages * Cfg::STAGE_ELEMS);
  int64_t* colB = rowA + BM;
  int64_t* rowC = colB + BN;
  int64_t* colC = rowC + BM;
  __shared__ double red[33];
  __shared__ bool last_block;

  const StepDev& s = steps[step_idx];
  const GemmDev& g = gemms[s.gemm_off];
  const double* A = (const double*)d.ptrs[s.in_slot[0]];
  const double* Bp = (const double*)d.ptrs[s.in_slot[1]];
  double* C = (double*)d.ptrs[s.out_slot];
  double* scratch = (double*)(d.arena + g.scratch_off);
  const double alpha = d.inv_kappa[s.in_slot[0]] * d.inv_kappa[s.in_slot[1]];
  const int splits = g.splits;
  const int64_t M = g.M, N = g.N, K = g.K;
  const int64_t tilesM = (M + BM - 1) / BM, tilesN = (N + BN - 1) / BN;
  const int64_t k_slabs = (K + BK - 1) / BK;
  const int64_t slabs_per_split = (k_slabs + splits - 1) / splits;
  const int64_t total = g.B * tilesM * tilesN * splits;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp % Cfg::kWarpsM, wn = warp / Cfg::kWarpsM;
  const int frow = lane >> 2, fk = lane & 3;
  double local_abs = 0.0;

  for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
    int64_t t = w;
    const int64_t tm = t % tilesM; t /= tilesM;
    const int64_t tn = t % tilesN; t /= tilesN;
    const int64_t sp = t % splits;
    const int64_t b = t / splits;
    int64_t offAb = 0, offBb = 0, offCb = 0;
    if (g.n_batch > 0) {
      int64_t dummy;
      group_offsets(g.batch, g.n_batch, b, false, offAb, offCb);
      group_offsets(g.batch, g.n_batch, b, true, offBb, dummy);
    }
    __syncthreads();  // previous tile fully consumed (smem tiles and offset tables)
    for (int r = threadIdx.x; r < BM; r += kGemmThreads) {
      const int64_t m = tm * BM + r;
      int64_t oa = -1, oc = 0;
      if (m < M) group_offsets(g.m, g.n_m, m, false, oa, oc);
      rowA[r] = oa;
      rowC[r] = oc;
    }
    for (int r = threadIdx.x; r < BN; r += kGemmThreads) {
      const int64_t n = tn * BN + r;
      int64_t ob = -1, oc = 0;
      if (n < N) group_offsets(g.n, g.n_n, n, true, ob, oc);
      colB[r] = ob;
      colC[r] = oc;
    }
    __syncthreads();

    const int64_t k_begin = sp * slabs_per_split * BK;
    const int64_t k_end = min(K, k_begin + slabs_per_split * BK);
    const int64_t nk = k_end > k_begin ? (k_end - k_begin + BK - 1) / BK : 0;
    const double* Ab = A + offAb;
    const double* Bb = Bp + offBb;

    double acc[TM][TN][2];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int st = 0; st < kStages - 1; ++st) {
      if (st < nk) {
        double* sa = tiles + st * Cfg::STAGE_ELEMS;
        load_slab<BM, LA>(sa, Ab, rowA, g.k, g.n_k, false, k_begin + (int64_t)st * BK, k_end);
        load_slab<BN, LB>(sa + Cfg::A_ELEMS, Bb, colB, g.k, g.n_k, true, k_begin + (int64_t)st * BK, k_end);
      }
      cp_async_commit();
    }
    for (int64_t it = 0; it < nk; ++it) {
      cp_async_wait<kStages - 2>();
      __syncthreads();
      const int64_t nxt = it + kStages - 1;
      if (nxt < nk) {
        double* sa = tiles + (nxt % kStages) * Cfg::STAGE_ELEMS;
        load_slab<BM, LA>(sa, Ab, rowA, g.k, g.n_k, false, k_begin + nxt * BK, k_end);
        load_slab<BN, LB>(sa + Cfg::A_ELEMS, Bb, colB, g.k, g.n_k, true, k_begin + nxt * BK, k_end);
      }
      cp_async_commit();
      const double* As = tiles + (it % kStages) * Cfg::STAGE_ELEMS;
      const double* Bs = As + Cfg::A_ELEMS;
#pragma unroll
      for (int ks = 0; ks < BK / 4; ++ks) {
        double a[TM], bf[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = frag<BM, LA>(As, (wm * TM + i) * 8 + frow, ks * 4 + fk);
#pragma unroll
        for (int j = 0; j < TN; ++j) bf[j] = frag<BN, LB>(Bs, (wn * TN + j) * 8 + frow, ks * 4 + fk);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma884(acc[i][j], a[i], bf[j]);
      }
    }
    cp_async_wait<0>();

    // epilogue
    if (splits == 1) {
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int r = (wm * TM + i) * 8 + frow;
        if (rowA[r] < 0) continue;
        const int64_t ro = offCb + rowC[r];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = (wn * TN + j) * 8 + fk * 2 + e;
            if (colB[c] < 0) continue;
            const double v = acc[i][j][e] * alpha;
            C[ro + colC[c]] = v;
            local_abs += fabs(v);
          }
        }
      }
    } else {
      // raw partial sums, [split][b][m][n] contiguous; reduced by gemm_splitk_reduce
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const int64_t m = tm * BM + (wm * TM + i) * 8 + frow;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int64_t n = tn * BN + (wn * TN + j) * 8 + fk * 2 + e;
            if (n >= N) continue;
            scratch[((sp * g.B + b) * M + m) * N + n] = acc[i][j][e];
          }
        }
      }
    }
  }
  if (splits != 1) return;

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

// Second pass of split-K: sum the partial tiles in split order (deterministic),
// apply alpha, scatter to C's strides, fused stabilize epilogue.
__global__ void __launch_bounds__(kThreads) gemm_splitk_reduce_kernel(const StepDev* __restrict__ steps,
                                                                      const GemmDev* __restrict__ gemms, int step_idx,
                                                                      DevState d) {
  __shared__ double red[33];
  __shared__ bool last_block;
  const StepDev& s = steps[step_idx];
  const GemmDev& g = gemms[s.gemm_off];
  double* C = (double*)d.ptrs[s.out_slot];
  const double* scratch = (const double*)(d.arena + g.scratch_off);
  const double alpha = d.inv_kappa[s.in_slot[0]] * d.inv_kappa[s.in_slot[1]];
  const int64_t per = g.B * g.M * g.N;
  double local_abs = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += (int64_t)gridDim.x * blockDim.x) {
    double sum = 0.0;
    for (int sp = 0; sp < g.splits; ++sp) sum += scratch[(int64_t)sp * per + e];
    const int64_t n = e % g.N, m = (e / g.N) % g.M, b = e / (g.N * g.M);
    int64_t ox, ob, om, on;
    group_offsets(g.batch, g.n_batch, b, false, ox, ob);
    group_offsets(g.m, g.n_m, m, false, ox, om);
    group_offsets(g.n, g.n_n, n, true, ox, on);
    const double v = sum * alpha;
    C[ob + om + on] = v;
    local_abs += fabs(v);
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
// host side: eligibility, tile / split / layout choice, launch
// ---------------------------------------------------------------------------
static int pick_layout(const GemmMode* rows, int n_rows, const GemmMode* ks, int n_k, bool use_b) {
  auto stride = [&](const GemmMode& m) { return use_b ? m.sb : m.sa; };
  // 16-byte copies need: a unit-stride innermost mode of even extent, and every
  // other stride of this operand even (so that all chunk addresses stay aligned)
  auto others_even = [&](const GemmMode* skip) {
    for (int i = 0; i < n_rows; ++i)
      if (&rows[i] != skip && (stride(rows[i]) & 1)) return false;
    for (int i = 0; i < n_k; ++i)
      if (&ks[i] != skip && (stride(ks[i]) & 1)) return false;
    return true;
  };
  if (n_k > 0) {
    const GemmMode& in = ks[n_k - 1];
    if (stride(in) == 1 && (in.extent % 2 == 0) && others_even(&in)) return kKFast;
  }
  if (n_rows > 0) {
    const GemmMode& in = rows[n_rows - 1];
    if (stride(in) == 1 && (in.extent % 2 == 0) && others_even(&in)) return kMNFast;
  }
  return kAnyLayout;
}

bool gemm_supported(const Plan& p, const StepDev& s, const GemmDev& g) {
  if (p.dtype != CTN_F64) return false;
  if (s.n_in != 2 || g.K < 8) return false;
  if (std::max(g.M, g.N) < 8 && g.K < 1024) return false;
  return true;
}

void gemm_configure(const Plan& p, GemmDev& g) {
  // tile shapes: 0 128x128, 1 128x32, 2 32x128, 5 128x64, 6 64x128 (3 and 4 are the skinny kernel)
  g.cfg = (g.N <= 32) ? 1 : (g.M <= 32) ? 2 : (g.N <= 64) ? 5 : (g.M <= 64) ? 6 : 0;
  const int64_t bm = g.cfg == 2 ? 32 : g.cfg == 6 ? 64 : 128, bn = g.cfg == 1 ? 32 : g.cfg == 5 ? 64 : 128;
  const int64_t tiles = g.B * ((g.M + bm - 1) / bm) * ((g.N + bn - 1) / bn);
  g.splits = 1;
  if (tiles < p.sm_count && g.K >= 1024) {
    const int64_t by_k = (g.K + 255) / 256;
    const int64_t by_sm = (2 * (int64_t)p.sm_count + tiles - 1) / tiles;
    g.splits = (int32_t)std::max<int64_t>(1, std::min<int64_t>(std::min(by_k, by_sm), 1024));
  }
  // batch offsets are added to both the 16-byte-aligned base and the row
  // offsets, so odd batch strides also rule out vector copies
  bool batch_even_a = true, batch_even_b = true;
  for (int i = 0; i < g.n_batch; ++i) {
    batch_even_a = batch_even_a && !(g.batch[i].sa & 1);
    batch_even_b = batch_even_b && !(g.batch[i].sb & 1);
  }
  g.la = batch_even_a ? pick_layout(g.m, g.n_m, g.k, g.n_k, false) : kAnyLayout;
  g.lb = batch_even_b ? pick_layout(g.n, g.n_n, g.k, g.n_k, true) : kAnyLayout;
  g.scratch_off = 0;
}

size_t gemm_scratch_bytes(const GemmDev& g) {
  return g.splits > 1 ? (size_t)g.splits * g.B * g.M * g.N * sizeof(double) : 0;
}

template <class Cfg, int LA, int LB>
static cudaError_t launch_cfg(const ExecCtx& c, const Launch& l, const GemmDev& g) {
  const Plan& p = *c.plan;
  auto kern = gemm_f64_kernel<Cfg, LA, LB>;
  static AttrOnce attr;  // per instantiation
  {
    cudaError_t e = attr.ensure(kern, p.device, (int)Cfg::SMEM);
    if (e != cudaSuccess) return e;
  }
  const int64_t tilesM = (g.M + Cfg::BM - 1) / Cfg::BM, tilesN = (g.N + Cfg::BN - 1) / Cfg::BN;
  const int64_t total = g.B * tilesM * tilesN * g.splits;
  const int ctas_per_sm = (Cfg::BM * Cfg::BN <= 128 * 64) ? 2 : 1;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(total, (int64_t)p.sm_count * ctas_per_sm));
  kern<<<grid, kGemmThreads, Cfg::SMEM, c.stream>>>(p.d_steps, p.d_gemms, l.first_step, dev_state(c));
  return cudaGetLastError();
}

template <class Cfg>
static cudaError_t launch_layouts(const ExecCtx& c, const Launch& l, const GemmDev& g, int la, int lb) {
#define CTN_GEMM_CASE(A_, B_) \
  if (la == A_ && lb == B_) return launch_cfg<Cfg, A_, B_>(c, l, g);
  CTN_GEMM_CASE(kKFast, kKFast)
  CTN_GEMM_CASE(kKFast, kMNFast)
  CTN_GEMM_CASE(kKFast, kAnyLayout)
  CTN_GEMM_CASE(kMNFast, kKFast)
  CTN_GEMM_CASE(kMNFast, kMNFast)
  CTN_GEMM_CASE(kMNFast, kAnyLayout)
  CTN_GEMM_CASE(kAnyLayout, kKFast)
  CTN_GEMM_CASE(kAnyLayout, kMNFast)
  CTN_GEMM_CASE(kAnyLayout, kAnyLayout)
#undef CTN_GEMM_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_gemm(const ExecCtx& c, const Launch& l, bool a_aligned, bool b_aligned) {
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  const GemmDev& g = p.gemms[s.gemm_off];
  const int la = a_aligned ? g.la : kAnyLayout, lb = b_aligned ? g.lb : kAnyLayout;
  switch (g.cfg) {
    case 0: return launch_layouts<Cfg128x128>(c, l, g, la, lb);
    case 1: return launch_layouts<Cfg128x32>(c, l, g, la, lb);
    case 2: return launch_layouts<Cfg32x128>(c, l, g, la, lb);
    case 5: return launch_layouts<Cfg128x64>(c, l, g, la, lb);
    default: return launch_layouts<Cfg64x128>(c, l, g, la, lb);
  }
}

cudaError_t launch_gemm_reduce(const ExecCtx& c, const Launch& l) {
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  const GemmDev& g = p.gemms[s.gemm_off];
  const int64_t per = g.B * g.M * g.N;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((per + kThreads - 1) / kThreads, (int64_t)p.sm_count * 4));
  gemm_splitk_reduce_kernel<<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_gemms, l.first_step, dev_state(c));
  return cudaGetLastError();
}

}  // namespace ctn
