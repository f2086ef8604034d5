// Device kernels of the B200 contraction engine (sm_100a).
//
//  generic_step   one (n-ary) einsum step with arbitrary element strides:
//                 C[o] = alpha * sum_k prod_i A_i[off_i(o,k)], one output element
//                 per thread, fused epilogue: alpha = prod of the inputs' pending
//                 1/kappa, per-CTA sum|C|, last-arriving CTA turns the partials
//                 into kappa = mean|C| and adds ln(kappa) to the log-scale
//                 register (stabilize, contractn/einsum.py:89-107).
//  chain          n-ary same-layout Hadamard product (copy-node hubs,
//                 "ab,ab,...,ab->ab"): every operand is streamed from HBM exactly
//                 once with 128-bit loads, products carry a per-element exponent
//                 so no intermediate rescale pass is needed.
//  interp         a run of tiny steps executed by ONE CTA in one launch (the
//                 1000-step 3x3 chain): no launch or grid-wide sync per step.
//  finalize       T / kappa (and * exp(c) for the dense output mode).
#include <cstdlib>
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>

#include "ctn_internal.h"
#include "device_common.cuh"

namespace ctn {

// (Re)initialise the per-execution state block: log-scale 0, flags 0, every pending scale 1,
// arrival counters 0, and the slot -> device pointer table, which travels in the kernel
// parameters (no host->device copy sits between two executions on a stream).
__global__ void init_state_kernel(DevState d, int n_slots, int n_tickets, int chunk_start, int chunk_n, PtrChunk chunk) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
  if (chunk_start == 0) {
    if (t == 0) {
      d.hdr->log_scale = 0.0;
      d.hdr->flags = 0;
      d.hdr->chain_exp = 0;
      d.hdr->chain_mean = 1.0;
      d.hdr->reserved = 0.0;
      d.hdr->final_abs = 0.0;
    }
    for (int i = t; i < n_tickets; i += stride) d.tickets[i] = 0u;
    for (int i = t; i < n_slots; i += stride) {
      d.inv_kappa[i] = 1.0;
      d.kappa[i] = 1.0;
    }
  }
  for (int i = t; i < chunk_n; i += stride) d.ptrs[chunk_start + i] = chunk.p[i];
}

cudaError_t launch_init_state(const Plan& p, char* ws, cudaStream_t stream, const void* const* slot_ptrs) {
  ExecCtx c{&p, ws, stream, 0, nullptr, nullptr};
  DevState d = dev_state(c);
  const int n_slots = (int)p.slots.size();
  for (int start = 0; start < n_slots; start += kInitChunk) {
    PtrChunk chunk;
    const int n = std::min(kInitChunk, n_slots - start);
    for (int i = 0; i < n; ++i) chunk.p[i] = const_cast<void*>(slot_ptrs[start + i]);
    const int grid = start == 0 ? std::max(1, std::min(64, (std::max(n_slots, p.n_tickets) + 255) / 256)) : 2;
    init_state_kernel<<<grid, 256, 0, stream>>>(d, n_slots, p.n_tickets, start, n, chunk);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

template <typename T, bool COHERENT>
__device__ __forceinline__ double load_elem(const T* p) {
  // COHERENT: the value may have been produced earlier in this same launch, in global OR in
  // shared memory (interpreter): a volatile generic load is valid for both and is not cached in L1
  if (COHERENT) return (double)*reinterpret_cast<const volatile T*>(p);
  return (double)__ldg(p);
}

// One output element of a generic step.
template <typename T, bool COHERENT, typename IDX>
__device__ __forceinline__ double step_element(const StepDev& s, const ModeDev* __restrict__ modes,
                                               const T* const* in, IDX o) {
  const int n_in = s.n_in;
  int64_t off[kMaxIn];
#pragma unroll
  for (int i = 0; i < kMaxIn; ++i) off[i] = 0;
  IDX rem = o;
  for (int m = s.n_out_modes - 1; m >= 0; --m) {
    const ModeDev& md = modes[m];
    const IDX e = (IDX)md.extent;
    const IDX idx = rem % e;
    rem /= e;
#pragma unroll
    for (int i = 0; i < kMaxIn; ++i)
      if (i < n_in) off[i] += (int64_t)idx * md.stride[i];
  }
  if (s.n_k_modes == 0) {
    double pr = 1.0;
#pragma unroll
    for (int i = 0; i < kMaxIn; ++i)
      if (i < n_in) pr *= load_elem<T, COHERENT>(in[i] + off[i]);
    return pr;
  }
  if (s.k_numel == 0) return 0.0;  // a summed index of extent 0: empty sum
  const ModeDev* km = modes + s.n_out_modes;
  const ModeDev& inner = km[s.n_k_modes - 1];
  const IDX inner_n = (IDX)inner.extent;
  const IDX outer_n = (IDX)(s.k_numel / inner.extent);
  double acc = 0.0;
  for (IDX ko = 0; ko < outer_n; ++ko) {
    int64_t o2[kMaxIn];
#pragma unroll
    for (int i = 0; i < kMaxIn; ++i) o2[i] = off[i];
    IDX r = ko;
    for (int m = s.n_k_modes - 2; m >= 0; --m) {
      const ModeDev& md = km[m];
      const IDX e = (IDX)md.extent;
      const IDX idx = r % e;
      r /= e;
#pragma unroll
      for (int i = 0; i < kMaxIn; ++i)
        if (i < n_in) o2[i] += (int64_t)idx * md.stride[i];
    }
    for (IDX kk = 0; kk < inner_n; ++kk) {
      double pr = 1.0;
#pragma unroll
      for (int i = 0; i < kMaxIn; ++i)
        if (i < n_in) pr *= load_elem<T, COHERENT>(in[i] + o2[i] + (int64_t)kk * inner.stride[i]);
      acc += pr;
    }
  }
  return acc;
}

// One output element of a PAIRWISE step with 32-bit offsets (interpreter short path: the generic
// routine above spends ~3000 cycles per 3x3 product on predicated 8-operand 64-bit index algebra).
template <typename T>
__device__ __forceinline__ double pair_element(const StepDev& s, const ModeDev* __restrict__ modes, const T* a,
                                               const T* b, uint32_t o) {
  uint32_t oa = 0, ob = 0, rem = o;
  for (int m = s.n_out_modes - 1; m > 0; --m) {
    const ModeDev& md = modes[m];
    const uint32_t e = (uint32_t)md.extent;
    const uint32_t q = rem / e, idx = rem - q * e;
    rem = q;
    oa += idx * (uint32_t)md.stride[0];
    ob += idx * (uint32_t)md.stride[1];
  }
  if (s.n_out_modes > 0) {
    oa += rem * (uint32_t)modes[0].stride[0];
    ob += rem * (uint32_t)modes[0].stride[1];
  }
  if (s.n_k_modes == 0) return load_elem<T, true>(a + oa) * load_elem<T, true>(b + ob);
  if (s.k_numel == 0) return 0.0;
  const ModeDev* km = modes + s.n_out_modes;
  const int nk = s.n_k_modes;
  const uint32_t inner_n = (uint32_t)km[nk - 1].extent;
  const uint32_t isa = (uint32_t)km[nk - 1].stride[0], isb = (uint32_t)km[nk - 1].stride[1];
  const uint32_t outer_n = (uint32_t)s.k_numel / inner_n;
  double acc = 0.0;
  for (uint32_t ko = 0; ko < outer_n; ++ko) {
    uint32_t pa = oa, pb = ob, r = ko;
    for (int m = nk - 2; m >= 0; --m) {
      const uint32_t e = (uint32_t)km[m].extent;
      const uint32_t q = r / e, idx = r - q * e;
      r = q;
      pa += idx * (uint32_t)km[m].stride[0];
      pb += idx * (uint32_t)km[m].stride[1];
    }
    for (uint32_t kk = 0; kk < inner_n; ++kk) {
      double pr = load_elem<T, true>(a + pa);
      pr *= load_elem<T, true>(b + pb);
      acc += pr;
      pa += isa;
      pb += isb;
    }
  }
  return acc;
}

// ---------------------------------------------------------------------------
// generic step kernel
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) generic_step_kernel(const StepDev* __restrict__ steps,
                                                                const ModeDev* __restrict__ all_modes, int step_idx,
                                                                DevState d) {
  __shared__ double red[33];
  __shared__ bool last_block;
  const StepDev& s = steps[step_idx];
  const ModeDev* modes = all_modes + s.mode_off;
  const T* in[kMaxIn];
  double alpha = 1.0;
#pragma unroll
  for (int i = 0; i < kMaxIn; ++i) {
    in[i] = nullptr;
    if (i < s.n_in) {
      in[i] = (const T*)d.ptrs[s.in_slot[i]];
      alpha *= d.inv_kappa[s.in_slot[i]];
    }
  }
  T* out = (T*)d.ptrs[s.out_slot];
  const int64_t numel = s.out_numel;
  double local = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool small = numel < (int64_t)0x7fffffff && s.k_numel < (int64_t)0x7fffffff;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < numel; o += stride) {
    double v = small ? step_element<T, false, uint32_t>(s, modes, in, (uint32_t)o)
                     : step_element<T, false, uint64_t>(s, modes, in, (uint64_t)o);
    const T r = (T)(v * alpha);
    out[o] = r;
    local += fabs((double)r);
  }
  const double bsum = block_sum(local, red);
  if (threadIdx.x == 0) {
    d.partials[blockIdx.x] = bsum;
    __threadfence();
    const uint32_t t = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last_block) {
    __threadfence();
    double acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(&d.partials[i]);
    const double norm = block_sum(acc, red);
    if (threadIdx.x == 0) {
      record_scale(d, s.out_slot, norm, numel);
      d.tickets[s.ticket] = 0;
    }
  }
}

// Row-wise variant for steps whose innermost output mode is long: a CTA takes
// one value of all outer output modes (decomposed once per row, not per element)
// and its threads run along the innermost mode, so loads and stores are
// coalesced whenever that mode is contiguous and no div/mod is left in the loop.
constexpr int kRowsKTable = 512;  // summed-mode offsets kept in shared memory up to this many
constexpr int kRowsPerItem = 4;      // default
constexpr int kRowsPerItemWide = 8;  // plenty of rows: fewer barriers and row decodes per byte moved

template <typename T, int NIN, typename ACC, int kRowsPerItem>
__global__ void __launch_bounds__(kThreads) generic_rows_kernel(const StepDev* __restrict__ steps,
                                                                const ModeDev* __restrict__ all_modes, int step_idx,
                                                                DevState d) {
  __shared__ double red[33];
  __shared__ bool last_block;
  __shared__ int64_t row_base[kRowsPerItem][NIN];
  __shared__ int64_t k_off[NIN][kRowsKTable];
  const StepDev& s = steps[step_idx];
  const ModeDev* modes = all_modes + s.mode_off;
  const T* in[NIN];
  double alpha = 1.0;
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    in[i] = (const T*)d.ptrs[s.in_slot[i]];
    alpha *= d.inv_kappa[s.in_slot[i]];
  }
  T* out = (T*)d.ptrs[s.out_slot];
  const ModeDev& inner = modes[s.n_out_modes - 1];
  const int64_t inner_n = inner.extent;
  const int64_t n_rows = s.out_numel / inner_n;
  const int64_t chunks = (inner_n + kThreads - 1) / kThreads;
  const int kn = (int)s.k_numel;  // <= kRowsKTable (checked by the launcher)
  const int n_k_modes = s.n_k_modes, n_outer = s.n_out_modes - 1;
  const ModeDev* km = modes + s.n_out_modes;
  // offsets of every summed index, once per CTA
  for (int k = threadIdx.x; k < kn; k += blockDim.x) {
    int64_t o[NIN];
#pragma unroll
    for (int i = 0; i < NIN; ++i) o[i] = 0;
    int r = k;
    for (int m = n_k_modes - 1; m >= 0; --m) {
      const int e = (int)km[m].extent;
      const int idx = r % e;
      r /= e;
#pragma unroll
      for (int i = 0; i < NIN; ++i) o[i] += (int64_t)idx * km[m].stride[i];
    }
#pragma unroll
    for (int i = 0; i < NIN; ++i) k_off[i][k] = o[i];
  }
  int64_t inner_stride[NIN];
#pragma unroll
  for (int i = 0; i < NIN; ++i) inner_stride[i] = inner.stride[i];
  const ACC alpha_t = (ACC)alpha;

  // a work item is kRowsPerItem consecutive rows x one 256-wide chunk of the inner mode: several
  // rows per barrier pair, and their loads are all in flight before the first result is needed
  double local = 0.0;
  const int64_t row_groups = (n_rows + kRowsPerItem - 1) / kRowsPerItem;
  const int64_t work_items = row_groups * chunks;
  for (int64_t w = blockIdx.x; w < work_items; w += gridDim.x) {
    const int64_t rg = w / chunks, chunk = w - rg * chunks;
    const int64_t row0 = rg * kRowsPerItem;
    __syncthreads();
    if (threadIdx.x < NIN * kRowsPerItem) {
      const int i = threadIdx.x % NIN, rr = threadIdx.x / NIN;
      int64_t o = 0;
      if (n_rows <= 0x7fffffff) {  // 32-bit divisions: the 64-bit ones are a ~400-cycle subroutine each
        uint32_t rem = (uint32_t)(row0 + rr);
        for (int m = n_outer - 1; m >= 0; --m) {
          const uint32_t e = (uint32_t)modes[m].extent;
          const uint32_t q = rem / e, idx = rem - q * e;
          rem = q;
          o += (int64_t)idx * modes[m].stride[i];
        }
      } else {
        int64_t rem = row0 + rr;
        for (int m = n_outer - 1; m >= 0; --m) {
          const int64_t e = modes[m].extent;
          const int64_t idx = rem % e;
          rem /= e;
          o += idx * modes[m].stride[i];
        }
      }
      row_base[rr][i] = o;
    }
    __syncthreads();
    const int64_t j = chunk * kThreads + threadIdx.x;
    if (j < inner_n) {
      ACC acc[kRowsPerItem];
      if (n_k_modes == 0) {
#pragma unroll
        for (int rr = 0; rr < kRowsPerItem; ++rr) {
          acc[rr] = (ACC)1;
          if (row0 + rr < n_rows) {
#pragma unroll
            for (int i = 0; i < NIN; ++i) acc[rr] *= (ACC)__ldg(in[i] + row_base[rr][i] + j * inner_stride[i]);
          }
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < kRowsPerItem; ++rr) acc[rr] = (ACC)0;
        for (int k = 0; k < kn; ++k) {
#pragma unroll
          for (int rr = 0; rr < kRowsPerItem; ++rr) {
            if (row0 + rr < n_rows) {
              ACC pr = (ACC)__ldg(in[0] + row_base[rr][0] + j * inner_stride[0] + k_off[0][k]);
#pragma unroll
              for (int i = 1; i < NIN; ++i) pr *= (ACC)__ldg(in[i] + row_base[rr][i] + j * inner_stride[i] + k_off[i][k]);
              acc[rr] += pr;
            }
          }
        }
      }
#pragma unroll
      for (int rr = 0; rr < kRowsPerItem; ++rr) {
        if (row0 + rr < n_rows) {
          const T r = (T)(acc[rr] * alpha_t);
          out[(row0 + rr) * inner_n + j] = r;
          local += (double)fabs(r);
        }
      }
    }
  }
  const double bsum = block_sum(local, red);
  if (threadIdx.x == 0) {
    d.partials[blockIdx.x] = bsum;
    __threadfence();
    const uint32_t t = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last_block) {
    __threadfence();
    double acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(&d.partials[i]);
    const double norm = block_sum(acc, red);
    if (threadIdx.x == 0) {
      record_scale(d, s.out_slot, norm, s.out_numel);
      d.tickets[s.ticket] = 0;
    }
  }
}

// ---------------------------------------------------------------------------
// dot kernel: <= kDotMaxOut outputs, each a sum over >= kDotMinK terms (inner products, norms, the
// closing step of a closed network).  Every CTA streams a grid-strided share of the summed index,
// per-(output, CTA) partials go to scratch, the last CTA adds them in a fixed order.
// ---------------------------------------------------------------------------
template <typename T, int NIN>
__global__ void __launch_bounds__(kThreads) dot_kernel(const StepDev* __restrict__ steps,
                                                       const ModeDev* __restrict__ all_modes, int step_idx,
                                                       DevState d) {
  __shared__ double red[33];
  __shared__ bool last_block;
  __shared__ int64_t obase[kDotMaxOut][NIN];
  const StepDev& s = steps[step_idx];
  const ModeDev* modes = all_modes + s.mode_off;
  const ModeDev* km = modes + s.n_out_modes;
  const int n_out = (int)s.out_numel, n_k_modes = s.n_k_modes;
  const T* in[NIN];
  double alpha = 1.0;
#pragma unroll
  for (int i = 0; i < NIN; ++i) {
    in[i] = (const T*)d.ptrs[s.in_slot[i]];
    alpha *= d.inv_kappa[s.in_slot[i]];
  }
  if (threadIdx.x < n_out) {
    int64_t o[NIN];
#pragma unroll
    for (int i = 0; i < NIN; ++i) o[i] = 0;
    int rem = threadIdx.x;
    for (int m = s.n_out_modes - 1; m >= 0; --m) {
      const int e = (int)modes[m].extent;
      const int idx = rem % e;
      rem /= e;
#pragma unroll
      for (int i = 0; i < NIN; ++i) o[i] += (int64_t)idx * modes[m].stride[i];
    }
#pragma unroll
    for (int i = 0; i < NIN; ++i) obase[threadIdx.x][i] = o[i];
  }
  __syncthreads();
  const ModeDev& inner = km[n_k_modes - 1];
  const int64_t inner_n = inner.extent;
  int64_t istr[NIN];
#pragma unroll
  for (int i = 0; i < NIN; ++i) istr[i] = inner.stride[i];
  double acc[kDotMaxOut];
#pragma unroll
  for (int o = 0; o < kDotMaxOut; ++o) acc[o] = 0.0;
  const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
  int64_t k_first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n_k_modes == 1 && n_out == 1) {
    // plain inner product: four independent strided loads per operand in flight per thread
    const T* p[NIN];
#pragma unroll
    for (int i = 0; i < NIN; ++i) p[i] = in[i] + obase[0][i];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (; k_first + 3 * gstride < s.k_numel; k_first += 4 * gstride) {
      double v0 = (double)__ldg(p[0] + k_first * istr[0]);
      double v1 = (double)__ldg(p[0] + (k_first + gstride) * istr[0]);
      double v2 = (double)__ldg(p[0] + (k_first + 2 * gstride) * istr[0]);
      double v3 = (double)__ldg(p[0] + (k_first + 3 * gstride) * istr[0]);
#pragma unroll
      for (int i = 1; i < NIN; ++i) {
        v0 *= (double)__ldg(p[i] + k_first * istr[i]);
        v1 *= (double)__ldg(p[i] + (k_first + gstride) * istr[i]);
        v2 *= (double)__ldg(p[i] + (k_first + 2 * gstride) * istr[i]);
        v3 *= (double)__ldg(p[i] + (k_first + 3 * gstride) * istr[i]);
      }
      a0 += v0; a1 += v1; a2 += v2; a3 += v3;
    }
    acc[0] = (a0 + a1) + (a2 + a3);
  }
#pragma unroll 2
  for (int64_t k = k_first; k < s.k_numel; k += gstride) {
    int64_t koff[NIN];
    if (n_k_modes == 1) {
#pragma unroll
      for (int i = 0; i < NIN; ++i) koff[i] = k * istr[i];
    } else {
      int64_t rem = k / inner_n;
      const int64_t j = k - rem * inner_n;
#pragma unroll
      for (int i = 0; i < NIN; ++i) koff[i] = j * istr[i];
      for (int m = n_k_modes - 2; m >= 0; --m) {
        const int64_t e = km[m].extent;
        const int64_t q = rem / e, idx = rem - q * e;
        rem = q;
#pragma unroll
        for (int i = 0; i < NIN; ++i) koff[i] += idx * km[m].stride[i];
      }
    }
    if (n_out == 1) {
      double pr = (double)__ldg(in[0] + obase[0][0] + koff[0]);
#pragma unroll
      for (int i = 1; i < NIN; ++i) pr *= (double)__ldg(in[i] + obase[0][i] + koff[i]);
      acc[0] += pr;
    } else {
#pragma unroll
      for (int o = 0; o < kDotMaxOut; ++o) {
        if (o < n_out) {
          double pr = (double)__ldg(in[0] + obase[o][0] + koff[0]);
#pragma unroll
          for (int i = 1; i < NIN; ++i) pr *= (double)__ldg(in[i] + obase[o][i] + koff[i]);
          acc[o] += pr;
        }
      }
    }
  }
  double* scratch = (double*)(d.arena + s.scratch_off);
#pragma unroll
  for (int o = 0; o < kDotMaxOut; ++o) {
    if (o < n_out) {
      const double b = block_sum(acc[o], red);
      if (threadIdx.x == 0) scratch[(size_t)o * gridDim.x + blockIdx.x] = b;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const uint32_t t = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  T* out = (T*)d.ptrs[s.out_slot];
  double norm = 0.0;
  for (int o = 0; o < n_out; ++o) {
    double part = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) part += __ldcg(&scratch[(size_t)o * gridDim.x + b]);
    const double total = block_sum(part, red);  // fixed tree: the same bits every run
    const T r = (T)(total * alpha);
    if (threadIdx.x == 0) out[o] = r;
    norm += fabs((double)r);
  }
  if (threadIdx.x == 0) {
    record_scale(d, s.out_slot, norm, s.out_numel);
    d.tickets[s.ticket] = 0;
  }
}

template <typename T>
static cudaError_t launch_dot(const ExecCtx& c, const Launch& l, int grid, int n_in) {
  const Plan& p = *c.plan;
  DevState d = dev_state(c);
  switch (n_in) {
    case 1: dot_kernel<T, 1><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d); break;
    case 2: dot_kernel<T, 2><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d); break;
    case 3: dot_kernel<T, 3><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d); break;
    case 4: dot_kernel<T, 4><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

template <typename T, typename ACC, int RPI>
static cudaError_t launch_rows(const ExecCtx& c, const Launch& l, int grid, int n_in) {
  const Plan& p = *c.plan;
  DevState d = dev_state(c);
#define CTN_ROWS_CASE(N_)                                                                                              \
  case N_:                                                                                                             \
    generic_rows_kernel<T, N_, ACC, RPI><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d); \
    break;
  switch (n_in) {
    CTN_ROWS_CASE(1)
    CTN_ROWS_CASE(2)
    CTN_ROWS_CASE(3)
    CTN_ROWS_CASE(4)
    default:
      return cudaErrorInvalidValue;
  }
#undef CTN_ROWS_CASE
  return cudaGetLastError();
}

cudaError_t launch_generic(const ExecCtx& c, const Launch& l) {
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  DevState d = dev_state(c);
  const int64_t cap = std::min<int64_t>((int64_t)p.sm_count * 8, kMaxPartials);
  if (s.tracked) {  // few outputs x long sum (set by the plan together with the scratch for the partials)
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((s.k_numel + 4 * kThreads - 1) / (4 * kThreads), cap));
    return p.dtype == CTN_F32 ? launch_dot<float>(c, l, grid, s.n_in) : launch_dot<double>(c, l, grid, s.n_in);
  }
  // long innermost output mode and few summed indices: row-wise kernel
  if (s.n_out_modes >= 1 && s.k_numel <= kRowsKTable && s.n_in <= 4) {
    const int64_t inner_n = p.modes[s.mode_off + s.n_out_modes - 1].extent;
    if (inner_n >= 64) {
      const int64_t n_rows = s.out_numel / inner_n;
      const int64_t chunks = (inner_n + kThreads - 1) / kThreads;
      // wide items only when they still give every CTA of a full grid several items
      const int64_t work_wide = ((n_rows + kRowsPerItemWide - 1) / kRowsPerItemWide) * chunks;
      const bool wide = work_wide >= 4 * cap;
      const int64_t work = wide ? work_wide : ((n_rows + kRowsPerItem - 1) / kRowsPerItem) * chunks;
      const int grid = (int)std::max<int64_t>(1, std::min(work, cap));
      // float data: short sums stay in fp32 (what the reference's einsum does too);
      // long ones are accumulated in fp64
      if (wide) {
        if (p.dtype == CTN_F64) return launch_rows<double, double, kRowsPerItemWide>(c, l, grid, s.n_in);
        if (s.k_numel <= 64) return launch_rows<float, float, kRowsPerItemWide>(c, l, grid, s.n_in);
        return launch_rows<float, double, kRowsPerItemWide>(c, l, grid, s.n_in);
      }
      if (p.dtype == CTN_F64) return launch_rows<double, double, kRowsPerItem>(c, l, grid, s.n_in);
      if (s.k_numel <= 64) return launch_rows<float, float, kRowsPerItem>(c, l, grid, s.n_in);
      return launch_rows<float, double, kRowsPerItem>(c, l, grid, s.n_in);
    }
  }
  int64_t blocks = (s.out_numel + kThreads - 1) / kThreads;
  int grid = (int)std::max<int64_t>(1, std::min(blocks, cap));
  if (p.dtype == CTN_F32)
    generic_step_kernel<float><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d);
  else
    generic_step_kernel<double><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_modes, l.first_step, d);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// finalize: canonical split output (and destabilize for the dense mode)
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T finish_value(T raw, double kappa, bool dense, double expc) {
  T v = raw / (T)kappa;                      // einsum.py:103  tensor / rescale
  if (dense) v = (T)((double)v * expc);      // einsum.py:114  tensor * exp(log_scale)
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) finalize_kernel(DevState d, int slot, int64_t numel, int dense,
                                                            double* log_scale_out, int32_t* flags_out) {
  T* out = (T*)d.ptrs[slot];
  const double kappa = d.kappa[slot];
  const double c = d.hdr->log_scale;
  const double expc = dense ? exp(c) : 1.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    out[i] = finish_value<T>(out[i], kappa, dense != 0, expc);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *log_scale_out = c;
    if (flags_out) *flags_out = d.hdr->flags;
  }
}

cudaError_t launch_finalize(const ExecCtx& c, const Launch& l) {
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  int64_t blocks = (s.out_numel + kThreads - 1) / kThreads;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)p.sm_count * 8));
  DevState d = dev_state(c);
  if (p.dtype == CTN_F32)
    finalize_kernel<float><<<grid, kThreads, 0, c.stream>>>(d, s.out_slot, s.out_numel, c.out_mode == CTN_OUT_DENSE,
                                                           c.log_scale_out, c.flags_out);
  else
    finalize_kernel<double><<<grid, kThreads, 0, c.stream>>>(d, s.out_slot, s.out_numel, c.out_mode == CTN_OUT_DENSE,
                                                            c.log_scale_out, c.flags_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// interpreter: a run of tiny steps in ONE launch.  One CTA; every step is
// executed by one warp (steps are dealt round-robin to the warps), warps only
// wait for the steps that produce their inputs (ready flags in shared memory),
// so independent branches of the contraction tree overlap and a dependent chain
// costs one L2 round trip per step instead of a kernel launch.
// ---------------------------------------------------------------------------
constexpr int kInterpWarps = 8;
#ifndef CTN_INTERP_BACKOFF
#define CTN_INTERP_BACKOFF do {} while (0)
#endif
#ifdef CTN_INTERP_PROFILE
__device__ long long g_interp_prof[4096 * 8];
extern "C" int ctn_debug_interp_profile(long long* host_out, int n) {
  return (int)cudaMemcpyFromSymbol(host_out, g_interp_prof, (size_t)n * 8 * sizeof(long long));
}
#define PROF(k) do { if (lane == 0 && li < 4096) g_interp_prof[li * 8 + (k)] = clock64(); } while (0)
#else
#define PROF(k) do {} while (0)
#endif

__device__ __forceinline__ void interp_fetch(unsigned char* dst, const char* src, int lane) {
  // one 1 KB record = 64 chunks of 16 bytes: two per lane, asynchronous
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(dst + lane * 16);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0), "l"(src + lane * 16) : "memory");
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 512), "l"(src + 512 + lane * 16) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <typename T>
__global__ void __launch_bounds__(kInterpWarps * 32) interp_kernel(const StepDev* __restrict__ steps,
                                                                  const char* __restrict__ recs,
                                                                  const InterpPreload* __restrict__ pre, int n_pre,
                                                                  int first, int last, DevState d, int dense,
                                                                  double* log_scale_out, int32_t* flags_out) {
  extern __shared__ __align__(16) unsigned char interp_smem[];
  // [per-warp record buffers 2 x 1 KB][ready flags: int x n_steps][pending 1/kappa: double x n_steps][pool]
  const int n_steps = last - first + 1;
  unsigned char* rec_buf = interp_smem;
  unsigned char* after = interp_smem + kInterpWarps * 2 * kInterpRecBytes;
  volatile int* ready = reinterpret_cast<volatile int*>(after);
  volatile double* inv_s = reinterpret_cast<volatile double*>(after + ((n_steps * 4 + 15) & ~15));
  unsigned char* pool = after + ((n_steps * 4 + 15) & ~15) + (size_t)n_steps * 8;
  __shared__ double warp_log[kInterpWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < n_steps; i += blockDim.x) ready[i] = 0;
  unsigned char* my_rec = rec_buf + warp * 2 * kInterpRecBytes;
  if (warp < n_steps) interp_fetch(my_rec, recs + (size_t)warp * kInterpRecBytes, lane);
  // stage the launch's external inputs (one thread per tensor: independent loads, all in flight together)
  for (int e = threadIdx.x; e < n_pre; e += blockDim.x) {
    const InterpPreload q = pre[e];
    const T* src = (const T*)d.ptrs[q.slot];
    T* dst = reinterpret_cast<T*>(pool + q.pool_off);
#pragma unroll 4
    for (int i = 0; i < q.n_elems; ++i) dst[i] = __ldcg(src + i);
  }
  __syncthreads();
  double log_acc = 0.0;  // lane 0 of each warp: sum of ln(kappa) of the steps it ran
  int buf = 0;
  for (int li = warp; li < n_steps; li += kInterpWarps, buf ^= 1) {
    // descriptor + modes of this step were fetched (cp.async) while the previous one ran;
    // start fetching the next one now
    if (li + kInterpWarps < n_steps) {
      interp_fetch(my_rec + (buf ^ 1) * kInterpRecBytes, recs + (size_t)(li + kInterpWarps) * kInterpRecBytes, lane);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    const StepDev& s = *reinterpret_cast<const StepDev*>(my_rec + buf * kInterpRecBytes);
    const ModeDev* modes = reinterpret_cast<const ModeDev*>(my_rec + buf * kInterpRecBytes + 256);
    const T* in[kMaxIn];
#pragma unroll
    for (int i = 0; i < kMaxIn; ++i) {
      in[i] = nullptr;
      if (i < s.n_in)
        in[i] = s.smem_in[i] >= 0 ? reinterpret_cast<const T*>(pool + s.smem_in[i]) : (const T*)d.ptrs[s.in_slot[i]];
    }
    T* out = (T*)d.ptrs[s.out_slot];
    T* out_s = s.smem_out >= 0 ? reinterpret_cast<T*>(pool + s.smem_out) : nullptr;
    PROF(0);
    const int numel = (int)s.out_numel;
    // Short path (pairwise, <= 32 outputs, at most one summed mode): one output per lane, and ALL of
    // its index algebra happens here, before the wait -- after it only loads and an FMA chain remain.
    const bool quick = out_s != nullptr && s.tracked && numel <= 32 && s.n_k_modes <= 1;
    uint32_t qa = 0, qb = 0, qsa = 0, qsb = 0, qn = 0;
    if (quick && lane < numel) {
      uint32_t rem = (uint32_t)lane;
      for (int m = s.n_out_modes - 1; m > 0; --m) {
        const uint32_t e = (uint32_t)modes[m].extent;
        const uint32_t q = rem / e, idx = rem - q * e;
        rem = q;
        qa += idx * (uint32_t)modes[m].stride[0];
        qb += idx * (uint32_t)modes[m].stride[1];
      }
      if (s.n_out_modes > 0) {
        qa += rem * (uint32_t)modes[0].stride[0];
        qb += rem * (uint32_t)modes[0].stride[1];
      }
      qn = 1;
      if (s.n_k_modes == 1) {
        const ModeDev& km = modes[s.n_out_modes];
        qn = (uint32_t)km.extent;
        qsa = (uint32_t)km.stride[0];
        qsb = (uint32_t)km.stride[1];
      }
    }
    // Wider short path (pairwise, <= 128 outputs, <= 2 summed modes): up to four outputs per lane, same idea.
    const bool quick4 = !quick && out_s != nullptr && s.tracked && numel <= 128 && s.n_k_modes <= 2;
    uint32_t wa[4] = {0, 0, 0, 0}, wb[4] = {0, 0, 0, 0};
    uint32_t w1n = 1, w1a = 0, w1b = 0, w2n = 1, w2a = 0, w2b = 0;
    if (quick4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t rem = (uint32_t)(lane + 32 * j);
        if ((int)rem < numel) {
          uint32_t oa = 0, ob = 0;
          for (int m = s.n_out_modes - 1; m > 0; --m) {
            const uint32_t e = (uint32_t)modes[m].extent;
            const uint32_t q = rem / e, idx = rem - q * e;
            rem = q;
            oa += idx * (uint32_t)modes[m].stride[0];
            ob += idx * (uint32_t)modes[m].stride[1];
          }
          if (s.n_out_modes > 0) {
            oa += rem * (uint32_t)modes[0].stride[0];
            ob += rem * (uint32_t)modes[0].stride[1];
          }
          wa[j] = oa;
          wb[j] = ob;
        }
      }
      const ModeDev* km = modes + s.n_out_modes;
      if (s.n_k_modes >= 1) {
        const ModeDev& in_m = km[s.n_k_modes - 1];
        w2n = (uint32_t)in_m.extent; w2a = (uint32_t)in_m.stride[0]; w2b = (uint32_t)in_m.stride[1];
      }
      if (s.n_k_modes == 2) {
        w1n = (uint32_t)km[0].extent; w1a = (uint32_t)km[0].stride[0]; w1b = (uint32_t)km[0].stride[1];
      }
    }
    // pending scales of inputs made by earlier launches: loaded before the wait, off the dependency chain
    // (the first two inputs are handled in registers: an 8-way predicated loop over the record costs
    // ~35 cycles per slot on the critical path)
    const int n_in = s.n_in;
    const int dep0 = s.dep[0], dep1 = n_in > 1 ? s.dep[1] : -2;
    const double numel_d = (double)s.out_numel;
    double alpha = 1.0;
    if (dep0 == -1) alpha *= __ldcg(&d.inv_kappa[s.in_slot[0]]);
    if (dep1 == -1) alpha *= __ldcg(&d.inv_kappa[s.in_slot[1]]);
    if (n_in > 2) {
#pragma unroll 1
      for (int i = 2; i < n_in; ++i)
        if (s.dep[i] == -1) alpha *= __ldcg(&d.inv_kappa[s.in_slot[i]]);
    }
    // wait for the producers of the inputs
    if (dep0 >= 0)
      while (ready[dep0] == 0) CTN_INTERP_BACKOFF;
    if (dep1 >= 0)
      while (ready[dep1] == 0) CTN_INTERP_BACKOFF;
    if (n_in > 2) {
#pragma unroll 1
      for (int i = 2; i < n_in; ++i)
        if (s.dep[i] >= 0)
          while (ready[s.dep[i]] == 0) CTN_INTERP_BACKOFF;
    }
    PROF(1);
    // pooled mode: everything made inside the launch is read from shared memory, written by a warp of
    // this CTA before its (volatile) ready flag -- a compiler barrier is enough on the reading side
    if (out_s) asm volatile("" ::: "memory");
    else __threadfence_block();
    if (dep0 >= 0) alpha *= inv_s[dep0];
    if (dep1 >= 0) alpha *= inv_s[dep1];
    if (n_in > 2) {
#pragma unroll 1
      for (int i = 2; i < n_in; ++i)
        if (s.dep[i] >= 0) alpha *= inv_s[s.dep[i]];
    }
    PROF(4);
    double local = 0.0;
    if (out_s) {
      // pooled mode: consumers inside this launch read the shared-memory copy, so publish that
      // first and keep everything else (global copy, scale arrays, the logarithm) off the
      // critical path of the dependency chain
      if (quick) {
        const T* a = in[0] + qa;
        const T* b = in[1] + qb;
        double acc = 0.0;
        uint32_t kk = 0;
        for (; kk + 4 <= qn; kk += 4) {
          const double a0 = (double)a[kk * qsa], a1 = (double)a[(kk + 1) * qsa], a2 = (double)a[(kk + 2) * qsa],
                       a3 = (double)a[(kk + 3) * qsa];
          const double b0 = (double)b[kk * qsb], b1 = (double)b[(kk + 1) * qsb], b2 = (double)b[(kk + 2) * qsb],
                       b3 = (double)b[(kk + 3) * qsb];
          acc = fma(a0, b0, acc); acc = fma(a1, b1, acc); acc = fma(a2, b2, acc); acc = fma(a3, b3, acc);
        }
        {  // tail of up to three terms: all loads first, then the FMA chain (x + 0*0 == x)
          const uint32_t r = qn - kk;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
          if (r > 0) { a0 = (double)a[kk * qsa]; b0 = (double)b[kk * qsb]; }
          if (r > 1) { a1 = (double)a[(kk + 1) * qsa]; b1 = (double)b[(kk + 1) * qsb]; }
          if (r > 2) { a2 = (double)a[(kk + 2) * qsa]; b2 = (double)b[(kk + 2) * qsb]; }
          acc = fma(a0, b0, acc); acc = fma(a1, b1, acc); acc = fma(a2, b2, acc);
        }
        if (lane < numel) {
          const T r = (T)(acc * alpha);
          out_s[lane] = r;
          local = fabs((double)r);
        }
      } else if (quick4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = lane + 32 * j;
          if (o < numel) {
            const T* a = in[0] + wa[j];
            const T* b = in[1] + wb[j];
            double acc = 0.0;
            for (uint32_t k1 = 0; k1 < w1n; ++k1) {
              const T* a1 = a + k1 * w1a;
              const T* b1 = b + k1 * w1b;
#pragma unroll 2
              for (uint32_t k2 = 0; k2 < w2n; ++k2) acc = fma((double)a1[k2 * w2a], (double)b1[k2 * w2b], acc);
            }
            const T r = (T)(acc * alpha);
            out_s[o] = r;
            local += fabs((double)r);
          }
        }
      } else if (s.tracked) {
        for (int o = lane; o < numel; o += 32) {
          const T r = (T)(pair_element<T>(s, modes, in[0], in[1], (uint32_t)o) * alpha);
          out_s[o] = r;
          local += fabs((double)r);
        }
      } else {
        for (int o = lane; o < numel; o += 32) {
          const double v = step_element<T, true, uint32_t>(s, modes, in, (uint32_t)o);
          const T r = (T)(v * alpha);
          out_s[o] = r;
          local += fabs((double)r);
        }
      }
      PROF(2);
      const double norm = warp_sum(local);
      PROF(5);
      const bool rescale = norm > kMinNorm;  // stabilize (einsum.py:89-107)
      // ONE division on the critical path: the applied factor is numel/norm; kappa and the log-scale
      // are derived from that same factor afterwards, so the pair stays exactly consistent
      const double inv = rescale ? numel_d / norm : 1.0;
      if (lane == 0) inv_s[li] = inv;
      __threadfence_block();
      __syncwarp();
      if (lane == 0) ready[li] = 1;
      PROF(3);
      for (int o = lane; o < numel; o += 32) out[o] = out_s[o];
      if (lane == 0) {
        const double kappa = rescale ? 1.0 / inv : 1.0;
        d.kappa[s.out_slot] = kappa;
        d.inv_kappa[s.out_slot] = inv;
        if (s.is_final) d.hdr->final_abs = rescale ? numel_d : norm;
        if (rescale) log_acc += log(kappa);
      }
    } else {
      for (int o = lane; o < numel; o += 32) {
        const double v = step_element<T, true, uint32_t>(s, modes, in, (uint32_t)o);
        const T r = (T)(v * alpha);
        out[o] = r;
        local += fabs((double)r);
      }
      const double norm = warp_sum(local);
      if (lane == 0) {
        const bool rescale = norm > kMinNorm;
        const double kappa = norm / (double)s.out_numel;
        const double inv = rescale ? 1.0 / kappa : 1.0;
        d.kappa[s.out_slot] = rescale ? kappa : 1.0;
        d.inv_kappa[s.out_slot] = inv;
        inv_s[li] = inv;
        if (s.is_final) d.hdr->final_abs = rescale ? (double)s.out_numel : norm;
        if (rescale) log_acc += log(kappa);
      }
      __threadfence_block();
      __syncwarp();
      if (lane == 0) ready[li] = 1;
    }
    __syncwarp();
  }
  if (lane == 0) warp_log[warp] = log_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int w = 0; w < kInterpWarps; ++w) total += warp_log[w];  // fixed order
    d.hdr->log_scale += total;
  }
  __syncthreads();
  const StepDev& s = steps[last];
  if (s.is_final) {
    T* out = (T*)d.ptrs[s.out_slot];
    const double kappa = __ldcg(&d.kappa[s.out_slot]);
    const double c = __ldcg(&d.hdr->log_scale);
    const double expc = dense ? exp(c) : 1.0;
    for (int i = threadIdx.x; i < (int)s.out_numel; i += blockDim.x)
      out[i] = finish_value<T>(__ldcg(&out[i]), kappa, dense != 0, expc);
    if (threadIdx.x == 0) {
      *log_scale_out = c;
      if (flags_out) *flags_out = d.hdr->flags;
    }
  }
}

template <typename T>
static cudaError_t launch_interp_t(const ExecCtx& c, const Launch& l) {
  const Plan& p = *c.plan;
  DevState d = dev_state(c);
  const int n_steps = l.last_step - l.first_step + 1;
  const size_t smem = (size_t)kInterpWarps * 2 * kInterpRecBytes + (((size_t)n_steps * 4 + 15) & ~(size_t)15) +
                      (size_t)n_steps * 8 + (size_t)l.smem_pool + 16;
  static AttrOnce attr;  // per instantiation
  {
    cudaError_t e = attr.ensure(interp_kernel<T>, p.device, 220 * 1024);
    if (e != cudaSuccess) return e;
  }
  interp_kernel<T><<<1, kInterpWarps * 32, smem, c.stream>>>(p.d_steps, p.d_interp_recs + (size_t)l.rec_off * kInterpRecBytes,
                                                            p.d_interp_pre + l.pre_off, l.pre_n, l.first_step, l.last_step, d,
                                                            c.out_mode == CTN_OUT_DENSE, c.log_scale_out, c.flags_out);
  return cudaGetLastError();
}

cudaError_t launch_interp(const ExecCtx& c, const Launch& l) {
  return c.plan->dtype == CTN_F32 ? launch_interp_t<float>(c, l) : launch_interp_t<double>(c, l);
}

// ---------------------------------------------------------------------------
// multi-GPU: rescale a batch shard to the global scale (one launch, scalars recomputed per thread)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) merge_shards_kernel(T* data, int64_t numel, const double* __restrict__ table,
                                                                int world, int rank, double* log_scale_out) {
  double c_star = -INFINITY;
  for (int i = 0; i < world; ++i) c_star = fmax(c_star, table[3 * i]);
  double S = 0.0, N = 0.0;
  for (int i = 0; i < world; ++i) {  // fixed order: every rank computes the same bits
    S += table[3 * i + 1] * exp(table[3 * i] - c_star);
    N += table[3 * i + 2];
  }
  const double kappa = S / N;
  const double scale = exp(table[3 * rank] - c_star) / kappa;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    data[i] = (T)((double)data[i] * scale);
  if (blockIdx.x == 0 && threadIdx.x == 0) *log_scale_out = c_star + log(kappa);
}

// ---------------------------------------------------------------------------
// multi-GPU: the whole merge of batch shards in ONE kernel, no host-launched collective.  Every rank
// posts its (log-scale, sum|T_shard|, numel_shard) triple -- read straight from the plan's state block
// -- into slot `rank` of every peer's mailbox (peer-mapped device memory, stores over NVLink), then
// waits until all `world` slots of its OWN mailbox carry this call's sequence number, recomputes the
// global scale in a fixed order (same bits on every rank) and rescales its shard in place.
// Mailbox: [2 parities][world][4 doubles] = (c, abs, numel, seq); a rank can be at most one call ahead
// of a peer (it needs the peer's post of call k to finish call k), so two parities suffice.
// ---------------------------------------------------------------------------
struct PeerBoxes {
  double* p[kMailboxMaxWorld];
};

__device__ __forceinline__ void st_release_sys(double* p, double v) {
  asm volatile("st.release.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_acquire_sys(const double* p) {
  double v;
  asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) merge_shards_peer_kernel(T* data, int64_t numel, const StateHdr* __restrict__ hdr,
                                                                     PeerBoxes boxes, int world, int rank,
                                                                     unsigned long long seq, double* log_scale_out) {
  __shared__ double tab[3 * kMailboxMaxWorld];
  __shared__ int give_up;
  const int par = (int)(seq & 1ull);
  const double seq_d = (double)seq;  // exact below 2^53
  if (blockIdx.x == 0 && threadIdx.x < world) {
    // post my triple into slot `rank` of peer threadIdx.x's mailbox: payload, then the flag with release order
    double* slot = boxes.p[threadIdx.x] + ((size_t)par * world + rank) * 4;
    slot[0] = hdr->log_scale;
    slot[1] = hdr->final_abs;
    slot[2] = (double)numel;
    __threadfence_system();
    st_release_sys(slot + 3, seq_d);
  }
  if (threadIdx.x == 0) give_up = 0;
  __syncthreads();
  if (threadIdx.x < world) {
    const double* slot = boxes.p[rank] + ((size_t)par * world + threadIdx.x) * 4;
    const long long t0 = clock64();
    while (ld_acquire_sys(slot + 3) != seq_d) {
      if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer never arrived; do not hang the GPU
        give_up = 1;
        break;
      }
    }
    tab[3 * threadIdx.x] = slot[0];
    tab[3 * threadIdx.x + 1] = slot[1];
    tab[3 * threadIdx.x + 2] = slot[2];
  }
  __syncthreads();
  double c_star = -INFINITY;
  for (int i = 0; i < world; ++i) c_star = fmax(c_star, tab[3 * i]);
  double S = 0.0, N = 0.0;
  for (int i = 0; i < world; ++i) {  // fixed order: every rank computes the same bits
    S += tab[3 * i + 1] * exp(tab[3 * i] - c_star);
    N += tab[3 * i + 2];
  }
  const double kappa = S / N;
  const double scale = give_up ? __longlong_as_double(0x7ff8000000000000LL) : exp(tab[3 * rank] - c_star) / kappa;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride)
    data[i] = (T)((double)data[i] * scale);
  if (blockIdx.x == 0 && threadIdx.x == 0) *log_scale_out = give_up ? scale : c_star + log(kappa);
}

cudaError_t launch_merge_shards_peer(void* data, int64_t numel, int dtype, const StateHdr* hdr, double* const* mailboxes,
                                     int world, int rank, unsigned long long seq, double* log_scale_out,
                                     cudaStream_t stream) {
  PeerBoxes boxes;
  for (int i = 0; i < kMailboxMaxWorld; ++i) boxes.p[i] = i < world ? mailboxes[i] : nullptr;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((numel + kThreads - 1) / kThreads, 148 * 8));
  if (dtype == CTN_F32)
    merge_shards_peer_kernel<float><<<grid, kThreads, 0, stream>>>((float*)data, numel, hdr, boxes, world, rank, seq, log_scale_out);
  else
    merge_shards_peer_kernel<double><<<grid, kThreads, 0, stream>>>((double*)data, numel, hdr, boxes, world, rank, seq, log_scale_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// multi-GPU: scale merge + all-to-all re-sharding of a boundary slab in ONE kernel over peer memory
// (contractn_b200/sharded.py).  Triple exchange as in merge_shards_peer_kernel; then every element is
// read ONCE from the local slab, multiplied by the common-scale factor (and by 1/kappa_raw when the
// producing plan skipped its finalize pass) and stored straight into the destination rank's receive
// buffer over NVLink: chunk j of [world][chunk] goes to recv[j] + rank * chunk.  blockIdx.y is the
// destination, so every link is busy at once and each CTA streams one contiguous range.  The last CTA
// to finish tells every peer "my stores have landed" (release flag after a system fence) and waits for
// the peers' flags: kernel completion on a rank means its receive buffer is complete.
// ---------------------------------------------------------------------------
struct PeerPtrs {
  void* p[kMailboxMaxWorld];
};

template <typename T>
struct Vec16;
template <>
struct Vec16<double> {
  using type = double2;
  static __device__ __forceinline__ double2 mul(double2 v, double s) { return make_double2(v.x * s, v.y * s); }
};
template <>
struct Vec16<float> {
  using type = float4;
  static __device__ __forceinline__ float4 mul(float4 v, double s) {
    const float f = (float)s;
    return make_float4(v.x * f, v.y * f, v.z * f, v.w * f);
  }
};

template <typename T>
__global__ void __launch_bounds__(kThreads) reshard_peer_kernel(const T* __restrict__ src, int64_t chunk,
                                                                const StateHdr* __restrict__ hdr,
                                                                const double* __restrict__ kappa_raw, PeerBoxes boxes,
                                                                PeerBoxes done, PeerPtrs recv, unsigned int* ticket,
                                                                int world, int rank, unsigned long long seq,
                                                                double* log_scale_out, int vec) {
  __shared__ double tab[3 * kMailboxMaxWorld];
  __shared__ int give_up, is_last;
  const int par = (int)(seq & 1ull);
  const double seq_d = (double)seq;
  const bool first = blockIdx.x == 0 && blockIdx.y == 0;
  if (first && threadIdx.x < world) {
    double* slot = boxes.p[threadIdx.x] + ((size_t)par * world + rank) * 4;
    slot[0] = hdr->log_scale;
    slot[1] = hdr->final_abs;
    slot[2] = (double)(chunk * world);
    __threadfence_system();
    st_release_sys(slot + 3, seq_d);
  }
  if (threadIdx.x == 0) give_up = 0;
  __syncthreads();
  if (threadIdx.x < world) {
    const double* slot = boxes.p[rank] + ((size_t)par * world + threadIdx.x) * 4;
    const long long t0 = clock64();
    while (ld_acquire_sys(slot + 3) != seq_d) {
      if (clock64() - t0 > 20000000000LL) {
        give_up = 1;
        break;
      }
    }
    tab[3 * threadIdx.x] = slot[0];
    tab[3 * threadIdx.x + 1] = slot[1];
    tab[3 * threadIdx.x + 2] = slot[2];
  }
  __syncthreads();
  double c_star = -INFINITY;
  for (int i = 0; i < world; ++i) c_star = fmax(c_star, tab[3 * i]);
  double S = 0.0, N = 0.0;
  for (int i = 0; i < world; ++i) {  // fixed order: every rank computes the same bits
    S += tab[3 * i + 1] * exp(tab[3 * i] - c_star);
    N += tab[3 * i + 2];
  }
  const double kappa = S / N;
  double scale = give_up ? __longlong_as_double(0x7ff8000000000000LL) : exp(tab[3 * rank] - c_star) / kappa;
  if (kappa_raw) scale /= *kappa_raw;
  const int j = blockIdx.y;  // destination rank
  const T* s = src + (int64_t)j * chunk;
  T* d = (T*)recv.p[j] + (int64_t)rank * chunk;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    using V = typename Vec16<T>::type;
    constexpr int E = 16 / (int)sizeof(T);
    const V* sv = (const V*)s;
    V* dv = (V*)d;
    const int64_t n = chunk / E;
    int64_t i = tid;
    for (; i + 3 * stride < n; i += 4 * stride) {  // four independent 16-byte loads in flight per thread
      V a = sv[i], b = sv[i + stride], c = sv[i + 2 * stride], e = sv[i + 3 * stride];
      dv[i] = Vec16<T>::mul(a, scale);
      dv[i + stride] = Vec16<T>::mul(b, scale);
      dv[i + 2 * stride] = Vec16<T>::mul(c, scale);
      dv[i + 3 * stride] = Vec16<T>::mul(e, scale);
    }
    for (; i < n; i += stride) dv[i] = Vec16<T>::mul(sv[i], scale);
  } else {
    for (int64_t i = tid; i < chunk; i += stride) d[i] = (T)((double)s[i] * scale);
  }
  if (first && threadIdx.x == 0) *log_scale_out = give_up ? scale : c_star + log(kappa);
  // completion: all my stores -> system fence -> ticket; the last CTA posts the flags and waits for the peers'
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence_system();
  if (threadIdx.x < world) {
    st_release_sys(done.p[threadIdx.x] + (size_t)par * world + rank, seq_d);
    const double* flag = done.p[rank] + (size_t)par * world + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) != seq_d)
      if (clock64() - t0 > 20000000000LL) break;
  }
  if (threadIdx.x == 0) *ticket = 0u;
}

cudaError_t launch_reshard_peer(const void* data, int64_t numel, int dtype, const StateHdr* hdr, const double* kappa_raw,
                                double* const* mailboxes, double* const* done_boxes, void* const* recv, uint32_t* ticket,
                                int world, int rank, unsigned long long seq, double* log_scale_out, cudaStream_t stream) {
  PeerBoxes boxes, done;
  PeerPtrs rp;
  for (int i = 0; i < kMailboxMaxWorld; ++i) {
    boxes.p[i] = i < world ? mailboxes[i] : nullptr;
    done.p[i] = i < world ? done_boxes[i] : nullptr;
    rp.p[i] = i < world ? recv[i] : nullptr;
  }
  const int64_t chunk = numel / world;
  const size_t elem = dtype == CTN_F32 ? 4 : 8;
  bool vec = (chunk * elem) % 16 == 0 && ((uintptr_t)data % 16) == 0;
  for (int i = 0; i < world; ++i) vec = vec && ((uintptr_t)recv[i] % 16) == 0;
  const int64_t per = vec ? 16 / (int64_t)elem : 1;
  const int64_t blocks = (chunk / per + kThreads - 1) / kThreads;
  // CTN_RESHARD_CTAS: experiment knob (read per call).  Fewer, longer store streams are FASTER over NVLink:
  // measured on 8 x B200, 134 MB slab: 1184 CTAs 349 GB/s per rank outgoing, 592 CTAs 586 GB/s
  const char* env = getenv("CTN_RESHARD_CTAS");
  const int env_ctas = env ? atoi(env) : 0;
  const int total_ctas = env_ctas > 0 ? env_ctas : 148 * 4;
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (total_ctas + world - 1) / world));
  dim3 grid(gx, world);
  if (dtype == CTN_F32)
    reshard_peer_kernel<float><<<grid, kThreads, 0, stream>>>((const float*)data, chunk, hdr, kappa_raw, boxes, done, rp,
                                                              ticket, world, rank, seq, log_scale_out, vec ? 1 : 0);
  else
    reshard_peer_kernel<double><<<grid, kThreads, 0, stream>>>((const double*)data, chunk, hdr, kappa_raw, boxes, done, rp,
                                                               ticket, world, rank, seq, log_scale_out, vec ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_merge_shards(void* data, int64_t numel, int dtype, const double* table, int world, int rank,
                                double* log_scale_out, cudaStream_t stream) {
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((numel + kThreads - 1) / kThreads, 148 * 8));
  if (dtype == CTN_F32)
    merge_shards_kernel<float><<<grid, kThreads, 0, stream>>>((float*)data, numel, table, world, rank, log_scale_out);
  else
    merge_shards_kernel<double><<<grid, kThreads, 0, stream>>>((double*)data, numel, table, world, rank, log_scale_out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Hadamard chain
// ---------------------------------------------------------------------------
template <typename T, bool VEC>
struct ChainCfg {
  static constexpr int kEpv = VEC ? 16 / (int)sizeof(T) : 1;  // elements per vector
  static constexpr int kVpt = VEC ? 2 : 4;                    // vectors per thread
  static constexpr int kEpt = kEpv * kVpt;
  static constexpr int kTile = kEpt * kThreads;
};

template <typename T, int EPV>
struct Vec;
template <>
struct Vec<double, 2> { using type = double2; };
template <>
struct Vec<float, 4> { using type = float4; };
template <>
struct Vec<double, 1> { using type = double; };
template <>
struct Vec<float, 1> { using type = float; };

template <typename T, int EPV>
__device__ __forceinline__ void load_vec(const T* p, double (&dst)[EPV]) {
  if constexpr (EPV == 1) {
    dst[0] = (double)__ldcs(p);
  } else if constexpr (sizeof(T) == 8) {
    const double2 v = __ldcs(reinterpret_cast<const double2*>(p));
    dst[0] = v.x; dst[1] = v.y;
  } else {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
}

template <typename T, int EPV>
__device__ __forceinline__ void store_vec(T* p, const double (&src)[EPV]) {
  if constexpr (EPV == 1) {
    *p = (T)src[0];
  } else if constexpr (sizeof(T) == 8) {
    *reinterpret_cast<double2*>(p) = make_double2(src[0], src[1]);
  } else {
    *reinterpret_cast<float4*>(p) = make_float4((float)src[0], (float)src[1], (float)src[2], (float)src[3]);
  }
}

// Pull the binary exponent of a normal double into an integer side-car.
__device__ __forceinline__ void split_exponent(double& p, int& x) {
  long long b = __double_as_longlong(p);
  const int ef = (int)((b >> 52) & 0x7ff);
  if (ef != 0 && ef != 0x7ff) {
    x += ef - 1023;
    b = (b & ~(0x7ffLL << 52)) | (1023LL << 52);
    p = __longlong_as_double(b);
  }
}

// merge (S, E) pairs meaning S * 2^E, keeping the larger exponent
__device__ __forceinline__ void merge_scaled(double& S, int& E, double s, int e) {
  if (s == 0.0) return;
  if (S == 0.0) { S = s; E = e; return; }
  if (e > E) { S = scalbn(S, E - e) + s; E = e; }
  else S += scalbn(s, e - E);
}

constexpr int kChainUnroll = 4;
constexpr int kRenormEvery = 4;

template <typename T, bool VEC, bool TRACKED>
__global__ void __launch_bounds__(kThreads) chain_kernel(const StepDev* __restrict__ steps,
                                                         const int32_t* __restrict__ slot_list, int step_idx,
                                                         DevState d) {
  using Cfg = ChainCfg<T, VEC>;
  constexpr int EPV = Cfg::kEpv, VPT = Cfg::kVpt, EPT = Cfg::kEpt;
  __shared__ const T* in_ptr[kChainMaxIn];
  __shared__ double red[33];
  __shared__ int redi[33];
  __shared__ bool last_block;
  const StepDev& s = steps[step_idx];
  const int n_in = s.n_in;
  const int64_t numel = s.out_numel;
  T* out = (T*)d.ptrs[s.out_slot];

  // stage the operand pointers and the combined pending scale of the inputs
  double a_part = 1.0;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) {
    const int slot = slot_list[s.slot_off + i];
    in_ptr[i] = (const T*)d.ptrs[slot];
    a_part *= d.inv_kappa[slot];
  }
  // product over the block (fixed order)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a_part *= __shfl_xor_sync(0xffffffffu, a_part, o);
    if (lane == 0) red[warp] = a_part;
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 1.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a *= red[w];
      red[32] = a;
    }
    __syncthreads();
  }
  const double alpha = red[32];
  __syncthreads();

  int32_t* tile_exp = TRACKED ? (int32_t*)(d.arena + s.scratch_off) : nullptr;
  const int64_t n_tiles = (numel + Cfg::kTile - 1) / Cfg::kTile;
  double blk_S = 0.0, blk_Sp = 0.0;  // running sums of this CTA (thread 0 only)
  int blk_E = 0, blk_Ep = 0;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t base = tile * Cfg::kTile;
    double p[VPT][EPV];
    int x[VPT][EPV];
    double pp[VPT][EPV];  // product before the last operand (TRACKED)
    int xp[VPT][EPV];
    bool live[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int64_t e0 = base + ((int64_t)v * kThreads + threadIdx.x) * EPV;
      live[v] = e0 < numel;  // numel % EPV == 0 is guaranteed by the launcher
#pragma unroll
      for (int e = 0; e < EPV; ++e) { p[v][e] = alpha; x[v][e] = 0; pp[v][e] = 0.0; xp[v][e] = 0; }
    }
    for (int i0 = 0; i0 < n_in; i0 += kChainUnroll) {
      double val[kChainUnroll][VPT][EPV];
#pragma unroll
      for (int u = 0; u < kChainUnroll; ++u) {
        const int i = i0 + u;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
          const int64_t e0 = base + ((int64_t)v * kThreads + threadIdx.x) * EPV;
          if (i < n_in && live[v]) load_vec<T, EPV>(in_ptr[i] + e0, val[u][v]);
          else {
#pragma unroll
            for (int e = 0; e < EPV; ++e) val[u][v][e] = 1.0;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kChainUnroll; ++u) {
        const int i = i0 + u;
        if (i < n_in) {
          if (TRACKED && i == n_in - 1) {
#pragma unroll
            for (int v = 0; v < VPT; ++v)
#pragma unroll
              for (int e = 0; e < EPV; ++e) { pp[v][e] = p[v][e]; xp[v][e] = x[v][e]; }
          }
#pragma unroll
          for (int v = 0; v < VPT; ++v)
#pragma unroll
            for (int e = 0; e < EPV; ++e) p[v][e] *= val[u][v][e];
        }
      }
      if (TRACKED) {
        // kChainUnroll == kRenormEvery: pull exponents out after every group
#pragma unroll
        for (int v = 0; v < VPT; ++v)
#pragma unroll
          for (int e = 0; e < EPV; ++e) split_exponent(p[v][e], x[v][e]);
      }
    }
    if (!TRACKED) {
      double local = 0.0;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        if (!live[v]) continue;
        const int64_t e0 = base + ((int64_t)v * kThreads + threadIdx.x) * EPV;
        double r[EPV];
#pragma unroll
        for (int e = 0; e < EPV; ++e) { r[e] = (double)(T)p[v][e]; local += fabs(r[e]); }
        store_vec<T, EPV>(out + e0, r);
      }
      const double tsum = block_sum(local, red);
      if (threadIdx.x == 0) blk_S += tsum;
    } else {
      // tile exponent = max exponent of its non-zero finite entries
      int ex = INT_MIN, exp_ = INT_MIN;
#pragma unroll
      for (int v = 0; v < VPT; ++v)
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
          split_exponent(pp[v][e], xp[v][e]);
          if (live[v] && p[v][e] != 0.0 && isfinite(p[v][e])) ex = max(ex, x[v][e]);
          if (live[v] && pp[v][e] != 0.0 && isfinite(pp[v][e])) exp_ = max(exp_, xp[v][e]);
        }
      int tE = block_max_int(ex, redi);
      int tEp = block_max_int(exp_, redi);
      if (tE == INT_MIN) tE = 0;
      if (tEp == INT_MIN) tEp = 0;
      double local = 0.0, localp = 0.0;
#pragma unroll
      for (int v = 0; v < VPT; ++v) {
        if (!live[v]) continue;
        const int64_t e0 = base + ((int64_t)v * kThreads + threadIdx.x) * EPV;
        double r[EPV];
#pragma unroll
        for (int e = 0; e < EPV; ++e) {
          r[e] = (double)(T)scalbn(p[v][e], max(x[v][e] - tE, -4000));
          local += fabs(r[e]);
          localp += fabs(scalbn(pp[v][e], max(xp[v][e] - tEp, -4000)));
        }
        store_vec<T, EPV>(out + e0, r);
      }
      const double tsum = block_sum(local, red);
      const double tsump = block_sum(localp, red);
      if (threadIdx.x == 0) {
        tile_exp[tile] = tE;
        merge_scaled(blk_S, blk_E, tsum, tE);
        merge_scaled(blk_Sp, blk_Ep, tsump, tEp);
      }
    }
  }

  // publish the CTA partials; the last CTA to arrive reduces them in index order
  if (threadIdx.x == 0) {
    d.partials[blockIdx.x] = blk_S;
    if (TRACKED) {
      d.partials[kMaxPartials + blockIdx.x] = blk_Sp;
      d.partial_exp[blockIdx.x] = blk_E;
      d.partial_exp[kMaxPartials + blockIdx.x] = blk_Ep;
    }
    __threadfence();
    const uint32_t t = atomicAdd(&d.tickets[s.ticket], 1u);
    last_block = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!last_block) return;
  __threadfence();
  if (!TRACKED) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) acc += __ldcg(&d.partials[i]);
    const double norm = block_sum(acc, red);
    if (threadIdx.x == 0) {
      record_scale(d, s.out_slot, norm, numel);
      d.tickets[s.ticket] = 0;
    }
  } else {
    // combine (S_i * 2^E_i) over the CTAs: common exponent first, then a
    // fixed-order block sum (deterministic, and parallel: grid can be ~2k)
    int e_loc = INT_MIN, ep_loc = INT_MIN;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      if (__ldcg(&d.partials[i]) != 0.0) e_loc = max(e_loc, __ldcg(&d.partial_exp[i]));
      if (__ldcg(&d.partials[kMaxPartials + i]) != 0.0) ep_loc = max(ep_loc, __ldcg(&d.partial_exp[kMaxPartials + i]));
    }
    int E = block_max_int(e_loc, redi);
    int Ep = block_max_int(ep_loc, redi);
    if (E == INT_MIN) E = 0;
    if (Ep == INT_MIN) Ep = 0;
    double s_loc = 0.0, sp_loc = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      s_loc += scalbn(__ldcg(&d.partials[i]), max(__ldcg(&d.partial_exp[i]) - E, -4000));
      sp_loc += scalbn(__ldcg(&d.partials[kMaxPartials + i]), max(__ldcg(&d.partial_exp[kMaxPartials + i]) - Ep, -4000));
    }
    const double S = block_sum(s_loc, red);
    const double Sp = block_sum(sp_loc, red);
    if (threadIdx.x == 0) {
      // reference decision for the LAST fused step (einsum.py:102): its input was
      // the normalised running product, so its norm is  sum|T_n| / mean|T_{n-1}|
      const double log_ratio = (log(S) + E * M_LN2) - (log(Sp) + Ep * M_LN2) + log((double)numel);
      const bool fine = isfinite(S) && isfinite(Sp) && S > 0.0 && Sp > 0.0 && log_ratio > log(kMinNorm) + 1e-6;
      const double mean = S / (double)numel;
      d.hdr->chain_exp = E;
      d.hdr->chain_mean = fine ? mean : 1.0;
      d.hdr->final_abs = (double)numel;  // the materialise pass leaves mean |T| = 1 (a degenerate run is re-executed)
      if (fine) d.hdr->log_scale += E * M_LN2 + log(mean);
      else d.hdr->flags |= CTN_RESULT_DEGENERATE;
      d.kappa[s.out_slot] = 1.0;      // the materialise pass leaves a normalised tensor
      d.inv_kappa[s.out_slot] = 1.0;
      d.tickets[s.ticket] = 0;
    }
  }
}

// Second pass of a tracked chain: bring every tile to the common exponent and
// divide by the mean (this IS the stabilize of the fused run); for the plan's
// final tensor also apply the output mode.
template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) chain_materialize_kernel(const StepDev* __restrict__ steps, int step_idx,
                                                                     DevState d, int dense, double* log_scale_out,
                                                                     int32_t* flags_out) {
  using Cfg = ChainCfg<T, VEC>;
  const StepDev& s = steps[step_idx];
  T* out = (T*)d.ptrs[s.out_slot];
  const int32_t* tile_exp = (const int32_t*)(d.arena + s.scratch_off);
  const int E = d.hdr->chain_exp;
  const double mean = d.hdr->chain_mean;
  const double c = d.hdr->log_scale;
  const bool final_dense = s.is_final && dense;
  const double expc = final_dense ? exp(c) : 1.0;
  const int64_t numel = s.out_numel;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
    const int64_t tile = i / Cfg::kTile;
    const double v = scalbn((double)out[i], max(tile_exp[tile] - E, -4000));
    T r = (T)v / (T)mean;
    if (final_dense) r = (T)((double)r * expc);
    out[i] = r;
  }
  if (s.is_final && blockIdx.x == 0 && threadIdx.x == 0) {
    *log_scale_out = c;
    if (flags_out) *flags_out = d.hdr->flags;
  }
}

template <typename T, bool VEC>
static cudaError_t launch_chain_t(const ExecCtx& c, const Launch& l) {
  using Cfg = ChainCfg<T, VEC>;
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  const int64_t n_tiles = (s.out_numel + Cfg::kTile - 1) / Cfg::kTile;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(n_tiles, kMaxPartials));
  DevState d = dev_state(c);
  if (s.tracked)
    chain_kernel<T, VEC, true><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_slot_list, l.first_step, d);
  else
    chain_kernel<T, VEC, false><<<grid, kThreads, 0, c.stream>>>(p.d_steps, p.d_slot_list, l.first_step, d);
  return cudaGetLastError();
}

template <typename T, bool VEC>
static cudaError_t launch_materialize_t(const ExecCtx& c, const Launch& l) {
  const Plan& p = *c.plan;
  const StepDev& s = p.steps[l.first_step];
  DevState d = dev_state(c);
  const int64_t blocks = (s.out_numel + kThreads - 1) / kThreads;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)p.sm_count * 8));
  chain_materialize_kernel<T, VEC><<<grid, kThreads, 0, c.stream>>>(p.d_steps, l.first_step, d,
                                                                   c.out_mode == CTN_OUT_DENSE, c.log_scale_out,
                                                                   c.flags_out);
  return cudaGetLastError();
}

static bool chain_vec(const Plan& p, const StepDev& s, bool vec_ok) {
  const int epv = p.dtype == CTN_F32 ? 4 : 2;
  return vec_ok && (s.out_numel % epv == 0);
}

cudaError_t launch_chain(const ExecCtx& c, const Launch& l, bool vec_ok) {
  const Plan& p = *c.plan;
  const bool vec = chain_vec(p, p.steps[l.first_step], vec_ok);
  if (p.dtype == CTN_F32) return vec ? launch_chain_t<float, true>(c, l) : launch_chain_t<float, false>(c, l);
  return vec ? launch_chain_t<double, true>(c, l) : launch_chain_t<double, false>(c, l);
}

// must pick the same tile geometry as the chain launch it follows
cudaError_t launch_materialize(const ExecCtx& c, const Launch& l, bool vec_ok) {
  const Plan& p = *c.plan;
  const bool vec = chain_vec(p, p.steps[l.first_step], vec_ok);
  if (p.dtype == CTN_F32) return vec ? launch_materialize_t<float, true>(c, l) : launch_materialize_t<float, false>(c, l);
  return vec ? launch_materialize_t<double, true>(c, l) : launch_materialize_t<double, false>(c, l);
}

}  // namespace ctn
