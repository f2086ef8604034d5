// Device helpers shared by kernels.cu and gemm.cu.
#pragma once
#include <climits>

#include "ctn_internal.h"

namespace ctn {

struct DevState {
  StateHdr* hdr;
  double* inv_kappa;
  double* kappa;
  uint32_t* tickets;
  double* partials;
  int32_t* partial_exp;
  void** ptrs;
  char* arena;
};

static inline DevState dev_state(const ExecCtx& c) {
  const StateLayout& st = c.plan->state;
  DevState d;
  d.hdr = (StateHdr*)(c.ws + st.hdr);
  d.inv_kappa = (double*)(c.ws + st.inv_kappa);
  d.kappa = (double*)(c.ws + st.kappa);
  d.tickets = (uint32_t*)(c.ws + st.tickets);
  d.partials = (double*)(c.ws + st.partials);
  d.partial_exp = (int32_t*)(c.ws + st.partial_exp);
  d.ptrs = (void**)(c.ws + st.ptrs);
  d.arena = c.ws + st.total;
  return d;
}

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block in a fixed order (deterministic). Result valid in all threads.
__device__ __forceinline__ double block_sum(double v, double* sm /* >= 33 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < (int)(blockDim.x >> 5) ? sm[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}

__device__ __forceinline__ int block_max_int(int v, int* sm /* >= 33 ints */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int t = lane < (int)(blockDim.x >> 5) ? sm[lane] : INT_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t = max(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (lane == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}

// stabilize (contractn/einsum.py:89-107) for a tensor whose sum|T| is `norm`:
// records kappa = norm/numel as the tensor's pending scale instead of dividing
// (the consumer folds 1/kappa into its alpha) and moves ln(kappa) to the register.
__device__ __forceinline__ void record_scale(const DevState& d, int slot, double norm, int64_t numel) {
  const bool rescale = norm > kMinNorm;  // einsum.py:102 (NaN compares false, as in numpy)
  const double kappa = norm / (double)numel;
  d.kappa[slot] = rescale ? kappa : 1.0;
  d.inv_kappa[slot] = rescale ? 1.0 / kappa : 1.0;
  d.hdr->final_abs = rescale ? (double)numel : norm;
  if (rescale) d.hdr->log_scale += log(kappa);
}


}  // namespace ctn
