// Internal structures shared by the plan compiler (host) and the kernels.
// Not part of the ABI; see include/ctn_b200.h for the public surface.
#pragma once
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "ctn_b200.h"

namespace ctn {

constexpr int kMaxIn = 8;          // inputs of one generic step
constexpr int kChainMaxIn = 512;   // inputs of one fused Hadamard-chain launch
constexpr int kMaxPartials = 148 * 16;
constexpr int kInterpMaxSteps = 4096;   // steps per interpreter launch (ready flags live in shared memory)
constexpr int kInterpPoolBytes = 128 * 1024;
constexpr int kInterpRecBytes = 1024;   // one packed record per interpreter step: StepDev at 0, its modes at 256
constexpr int kInterpRecModes = 10;     // (1024 - 256) / sizeof(ModeDev)  // shared-memory pool for the intermediates of one interpreter launch  // upper bound on the grid of any reducing kernel
constexpr double kMinNorm = 1e-7;  // contractn/einsum.py:94

// launch kinds (ctn_plan_launch_kind)
enum LaunchKind : int32_t {
  kGeneric = 0,
  kChain = 1,
  kInterp = 2,
  kGemm = 3,
  kFinalize = 4,
  kMaterialize = 5,  // second pass of a tracked Hadamard chain
  kGemmReduce = 6,   // second pass of a split-K GEMM
};

// One loop index of a step: extent and element stride in every input.
// Output strides are implicit: the output is C-contiguous over the out modes.
struct ModeDev {
  int64_t extent;
  int64_t stride[kMaxIn];
};

// GEMM view of a pairwise step: four groups of modes with strides into A, B, C.
struct GemmMode {
  int64_t extent;
  int64_t sa, sb, sc;
  int64_t sw;  // stride in the fused consumer's weight operand (0 unless fuse_pr > 0)
};
constexpr int kGemmMaxModes = 12;  // per group, after fusion

struct GemmDev {
  int32_t n_batch, n_m, n_n, n_k;
  GemmMode batch[kGemmMaxModes];
  GemmMode m[kGemmMaxModes];
  GemmMode n[kGemmMaxModes];
  GemmMode k[kGemmMaxModes];
  int64_t B, M, N, K;
  int32_t cfg;          // tile configuration (gemm.cu)
  int32_t splits;       // split-K factor (1 = single pass)
  int32_t la, lb;       // shared-memory staging layout of A and B
  int64_t scratch_off;  // split-K partial tiles, byte offset in the arena
  // fused consumer (gemm_tf32.cu): the GEMM result X[m, (r, n')] is never written; the next step
  // out[m, n'] = sum_r W[m, r] * X[m, (r, n')] is evaluated in the epilogue (in_slot[2] = W).
  // `n` then holds only the n' modes (sc = strides in `out`), N = prod(n'), and:
  int32_t fuse_pr;      // extent of r (0 = not fused; 2 or 4)
  int32_t fuse_pad;     // gemm_tf32.cu: bit 0 = the output admits 16-byte stores
  int64_t fuse_sb;      // stride of r in B
  int64_t fuse_sw;      // stride of r in W
};

// Device-side description of one step.
struct StepDev {
  int32_t kind;        // LaunchKind of the kernel that runs it
  int32_t n_in;
  int32_t in_slot[kMaxIn];  // generic / gemm steps
  int32_t slot_off;    // chain steps: offset into the plan's slot list
  int32_t out_slot;
  int32_t n_out_modes; // generic: the first n_out_modes modes index the output
  int32_t n_k_modes;   //          followed by n_k_modes summed modes
  int32_t mode_off;    // offset into the plan's mode table
  int32_t gemm_off;    // index into the plan's GemmDev table (-1 = none)
  int32_t ticket;      // index of this step's arrival counter
  int32_t is_final;    // output of the whole plan
  int32_t tracked;     // chain: per-element exponent tracking (n_in > 2); interp: pairwise 32-bit short path;
                       // generic: few outputs x long sum -> dot kernel (partials at scratch_off)
  int32_t dep[kMaxIn]; // interp: launch-local index of the step producing input i (-1: ready at launch)  (-1: made by an earlier launch, -2: operand)
  int32_t smem_out;    // interp: byte offset of this step's output in the launch's shared-memory pool (-1: none)
  int32_t smem_in[kMaxIn];  // interp: byte offset of input i in that pool (-1: read it from global memory)
  int64_t scratch_off; // chain: byte offset of the per-tile exponent array in the arena
  int64_t out_numel;
  int64_t k_numel;
};

// Per-execution mutable state, at the start of the caller's workspace.
struct StateHdr {
  double log_scale;   // contractn/einsum.py:338
  int32_t flags;      // CTN_RESULT_*
  int32_t chain_exp;  // global exponent of the last tracked chain
  double chain_mean;  // its mean |value| relative to 2^chain_exp
  double reserved;    // diagnostics of a timed-out pipeline wait (gemm_tf32.cu)
  double final_abs;   // sum |entries| of the tensor the LAST recorded step leaves after its rescale: numel when it
                      // was rescaled (mean |T| = 1), its raw norm when it was not (einsum.py:102); read by the
                      // multi-GPU merge (ctn_merge_batch_shards_peer) straight from the state block
};

struct StateLayout {
  size_t hdr = 0;          // StateHdr
  size_t inv_kappa = 0;    // double[n_slots]   pending 1/kappa per tensor
  size_t kappa = 0;        // double[n_slots]
  size_t tickets = 0;      // uint32[n_tickets]
  size_t partials = 0;     // double[4 * kMaxPartials]
  size_t partial_exp = 0;  // int32[2 * kMaxPartials]
  size_t ptrs = 0;         // void*[n_slots]
  size_t init_bytes = 0;   // [0, init_bytes) is (re)initialised by every execute
  size_t total = 0;        // start of the intermediates arena
};

struct Slot {
  std::vector<int32_t> labels;
  std::vector<int64_t> extents;
  std::vector<int64_t> strides;  // elements
  int64_t numel = 1;
  int32_t operand = -1;    // >= 0: caller operand index
  int64_t ws_offset = -1;  // >= 0: byte offset inside the intermediates arena
  bool is_final = false;
};

struct InterpPreload {
  int32_t slot;      // source tensor (operand or earlier intermediate)
  int32_t pool_off;  // byte offset in the launch's pool
  int32_t n_elems;   // address span in elements (1 + sum (extent-1)*stride): strided views keep their strides
  int32_t pad;
};

struct Launch {
  int32_t kind;
  int32_t first_step, last_step;  // inclusive range of StepDev entries
  int64_t bytes;                  // algorithmic bytes moved by this launch
  double flops;
  int32_t smem_pool = 0;          // interp: bytes of the shared-memory pool in use (0: global-memory mode)
  int32_t rec_off = 0;            // interp: index of the launch's first packed record
  int32_t pre_off = 0, pre_n = 0; // interp: inputs from outside the launch staged into the shared-memory pool
};

struct Plan {
  int32_t dtype = CTN_F64;
  int32_t flags = 0;
  int32_t n_operands = 0;
  size_t elem = 8;
  std::vector<Slot> slots;
  std::vector<StepDev> steps;
  std::vector<ModeDev> modes;
  std::vector<GemmDev> gemms;
  std::vector<int32_t> slot_list;  // chain inputs
  std::vector<Launch> launches;
  int32_t final_slot = -1;
  bool raw_final = false;  // CTN_PLAN_RAW_FINAL honoured: no finalize launch, output = T-hat * kappa[final_slot]
  int32_t n_tickets = 0;
  double flops = 0;
  StateLayout state;
  size_t arena_bytes = 0;
  // device copies (immutable after create)
  StepDev* d_steps = nullptr;
  ModeDev* d_modes = nullptr;
  GemmDev* d_gemms = nullptr;
  int32_t* d_slot_list = nullptr;
  std::vector<char> interp_recs;   // packed interpreter records (kInterpRecBytes each)
  char* d_interp_recs = nullptr;
  std::vector<InterpPreload> interp_pre;  // per interpreter launch: external inputs copied to the pool at kernel start
  InterpPreload* d_interp_pre = nullptr;
  int device = 0;
  int sm_count = 148;
  std::mutex host_mu;                     // serialises ctn_plan_execute_host (plan-owned staging below)
  // host-buffer path scratch (ctn_plan_execute_host)
  void* h_ws = nullptr;
  size_t h_ws_bytes = 0;
  std::vector<void*> h_operands;
  std::vector<size_t> h_operand_bytes;
  void* h_out = nullptr;
  double* h_log = nullptr;
  void* h_pack = nullptr;                 // one device buffer for all small operands
  std::vector<char> h_pack_host;          // their host-side staging image (one H2D copy)
  std::vector<int64_t> h_pack_off;        // offset of operand i inside the pack, -1 = copied on its own
};

// kernels.cu
struct ExecCtx {
  const Plan* plan;
  char* ws;             // workspace base (device)
  cudaStream_t stream;
  int out_mode;
  double* log_scale_out;
  int32_t* flags_out;
};
constexpr int kInitChunk = 448;  // slot pointers carried by one init launch (kernel parameters, < 4 KB)
struct PtrChunk {
  void* p[kInitChunk];
};
cudaError_t launch_init_state(const Plan& plan, char* ws, cudaStream_t stream, const void* const* slot_ptrs);
constexpr int kDotMaxOut = 16;          // dot kernel: at most this many outputs ...
constexpr int64_t kDotMinK = 1 << 15;   // ... each a sum over at least this many terms
constexpr int kMailboxMaxWorld = 64;   // ranks per peer mailbox (CTN_MAILBOX_WORDS = 2 * 4 * kMailboxMaxWorld doubles)
cudaError_t launch_merge_shards_peer(void* data, int64_t numel, int dtype, const StateHdr* hdr, double* const* mailboxes,
                                     int world, int rank, unsigned long long seq, double* log_scale_out,
                                     cudaStream_t stream);
cudaError_t launch_reshard_peer(const void* data, int64_t numel, int dtype, const StateHdr* hdr, const double* kappa_raw,
                                double* const* mailboxes, double* const* done_boxes, void* const* recv, uint32_t* ticket,
                                int world, int rank, unsigned long long seq, double* log_scale_out, cudaStream_t stream);
cudaError_t launch_merge_shards(void* data, int64_t numel, int dtype, const double* table, int world, int rank,
                                double* log_scale_out, cudaStream_t stream);
cudaError_t launch_generic(const ExecCtx& c, const Launch& l);
cudaError_t launch_chain(const ExecCtx& c, const Launch& l, bool vec_ok);
cudaError_t launch_interp(const ExecCtx& c, const Launch& l);
cudaError_t launch_gemm(const ExecCtx& c, const Launch& l, bool a_aligned, bool b_aligned);
cudaError_t launch_gemm_reduce(const ExecCtx& c, const Launch& l);
void gemm_configure(const Plan& p, GemmDev& g);
int skinny_config(const Plan& p, const StepDev& s, const GemmDev& g);
cudaError_t launch_skinny(const ExecCtx& c, const Launch& l);
bool gemm_tf32_supported(const Plan& p, const StepDev& s, const GemmDev& g);
void gemm_tf32_configure(const Plan& p, GemmDev& g);
cudaError_t launch_gemm_tf32(const ExecCtx& c, const Launch& l, bool a_aligned, bool b_aligned, bool c_aligned,
                             const void* b_ptr, const void* a_ptr);
void gemm_tf32_v1_configure(const Plan& p, GemmDev& g);  // round-1 kernel (gemm_tf32_v1.cu), CTN_TF32_V1=1
cudaError_t launch_gemm_tf32_v1(const ExecCtx& c, const Launch& l, bool a_aligned, bool b_aligned);
size_t gemm_scratch_bytes(const GemmDev& g);
cudaError_t launch_finalize(const ExecCtx& c, const Launch& l);
cudaError_t launch_materialize(const ExecCtx& c, const Launch& l, bool vec_ok);
bool gemm_supported(const Plan& p, const StepDev& s, const GemmDev& g);

void set_cuda_error(cudaError_t e, const char* where);

// cudaFuncSetAttribute is per device: remember per (kernel instantiation, device) that it was done.
struct AttrOnce {
  bool done[64] = {};
  template <class K>
  cudaError_t ensure(K kern, int device, int smem_bytes) {
    const int i = device >= 0 && device < 64 ? device : 63;
    if (done[i] && device < 63) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) done[i] = true;
    return e;
  }
};

}  // namespace ctn
