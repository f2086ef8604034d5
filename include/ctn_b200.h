/*
 * ctn_b200.h -- C ABI of the B200-native contraction engine for ContracTN.
 *
 * This is the drop-in boundary for ONE path of jemisjoky/ContracTN: the
 * stabilised pairwise contraction executor
 *
 *     _core_contract(operands, contract_list, backend)  contractn/einsum.py:326-393
 *     stabilize(tensor, log_scale, backend)             contractn/einsum.py:89-107
 *     destabilize(tensor, log_scale, backend)           contractn/einsum.py:110-114
 *
 * The reference is pure Python and has no FFI of its own; the entry points
 * below are what a ctypes binding placed at the `_core_contract` call site
 * (contractn/einsum.py:303) binds.  INTEGRATION.md shows that binding.
 *
 * Conventions
 *  - plain C, no C++/torch types; every function returns an int status
 *    (CTN_OK == 0) and never throws; ctn_status_string() explains a status.
 *  - the caller owns every data buffer passed to ctn_plan_execute().  For that
 *    entry point a plan is immutable metadata (host structs + a copy of the step
 *    descriptors on the device that was current at ctn_plan_create), so one plan
 *    may be executed concurrently on distinct streams with distinct workspaces,
 *    always with that same device current (CTN_ERR_INVALID otherwise).
 *    ctn_plan_execute_host() additionally keeps staging buffers inside the plan;
 *    calls on one plan are serialised by a mutex inside it.
 *  - execution is asynchronous on the given stream; no hidden synchronisation
 *    in ctn_plan_execute().  ctn_plan_execute_host() is synchronous.
 *  - element strides, not byte strides.  Outputs are C-contiguous in the order
 *    of the last step's output subscript.
 *  - index labels are arbitrary non-negative int32 (the host maps einsum
 *    symbols, which may be any unicode code point, to integers).
 */
#ifndef CTN_B200_H
#define CTN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTN_ABI_VERSION 1

/* status codes */
enum {
  CTN_OK = 0,
  CTN_ERR_INVALID = 1,     /* malformed description (ranks, labels, positions) */
  CTN_ERR_SHAPE = 2,       /* extents of a shared label disagree                */
  CTN_ERR_UNSUPPORTED = 3, /* valid but outside what the kernels handle         */
  CTN_ERR_CUDA = 4,        /* a CUDA runtime call failed; see ctn_last_cuda_error */
  CTN_ERR_NOMEM = 5,
  CTN_ERR_NO_DEVICE = 6,   /* no sm_100 device visible                          */
  CTN_ERR_INTERNAL = 7     /* a kernel gave up waiting on its own pipeline (bug) */
};

/* arithmetic type of the whole plan */
enum { CTN_F32 = 0, CTN_F64 = 1 };

/* plan flags */
enum {
  /* Execute every pairwise step on its own, with the reference's per-step
   * rescale decisions (contractn/einsum.py:102).  Without it, runs of
   * same-layout Hadamard steps (copy-node hubs) are fused into one n-ary
   * kernel; the final (tensor, log-scale) is identical (the per-step rescale is
   * a gauge) unless a norm on the way is <= 1e-7, which the fused kernel
   * detects and reports through CTN_RESULT_DEGENERATE. */
  CTN_PLAN_EXACT_STEPS = 1,
  /* Disable the single-CTA interpreter for runs of tiny steps. */
  CTN_PLAN_NO_INTERPRETER = 2,
  /* Disable the tiled tensor-core GEMM kernels (use the generic kernel). */
  CTN_PLAN_NO_GEMM = 4,
  /* Leave the result RAW: no finalize pass.  out = canonical T-hat times the pending scale kappa that
   * sits in the workspace at ctn_plan_final_kappa_offset(); the log-scale in the state block already
   * includes ln(kappa); *log_scale_out / *flags_out are NOT written.  For ctn_reshard_peer, which
   * folds 1/kappa into the scale it applies while the data crosses NVLink. */
  CTN_PLAN_RAW_FINAL = 8
};

/* output modes of ctn_plan_execute */
enum {
  CTN_OUT_SPLIT = 0, /* out = T / mean|T|, *log_scale = running log-scale      */
  CTN_OUT_DENSE = 1  /* out = (T / mean|T|) * exp(log_scale)  (destabilize)    */
};

/* result flags written next to the log-scale */
enum {
  CTN_RESULT_DEGENERATE = 1, /* a fused step met a norm <= 1e-7 or a non-finite
                                value: re-run with CTN_PLAN_EXACT_STEPS to get
                                the reference's exact (tensor, log-scale) pair */
  CTN_RESULT_INTERNAL = 2    /* a kernel's bounded pipeline wait expired: the result
                                is invalid (ctn_plan_execute_host returns
                                CTN_ERR_INTERNAL; device-path callers must test it) */
};

/* One entry of the contraction list (contractn/einsum.py:341-344): the
 * operands at `positions` (in the order given, i.e. descending) are popped from
 * the running operand list and their contraction is appended to it. */
typedef struct {
  int32_t n_inputs;            /* >= 1; 2 for every pairwise step               */
  const int32_t* positions;    /* [n_inputs]                                     */
  const int32_t* input_ranks;  /* [n_inputs] number of subscripts of each input  */
  const int32_t* input_labels; /* concatenated subscripts of the inputs          */
  int32_t out_rank;
  const int32_t* out_labels;   /* [out_rank] output subscript of the step        */
} ctn_step_desc;

typedef struct {
  int32_t abi_version;            /* CTN_ABI_VERSION                             */
  int32_t dtype;                  /* CTN_F32 | CTN_F64                           */
  int32_t n_operands;
  const int32_t* operand_ranks;   /* [n_operands]                                */
  const int64_t* operand_extents; /* concatenated                                */
  const int64_t* operand_strides; /* concatenated, elements; NULL = C-contiguous */
  int32_t n_steps;                /* 0 is invalid: a 1-operand network still has
                                     one step (contractn/einsum.py:341)          */
  const ctn_step_desc* steps;
  int32_t flags;                  /* CTN_PLAN_*                                  */
} ctn_plan_desc;

typedef struct ctn_plan ctn_plan;

/* Build a device-resident plan from a contraction list.  Replaces the Python
 * loop header of _core_contract (contractn/einsum.py:331-344): operand
 * bookkeeping, routing (einsum.py:347) and axis computation (einsum.py:357-368)
 * are resolved here once instead of on every call. */
int ctn_plan_create(const ctn_plan_desc* desc, ctn_plan** plan_out);
void ctn_plan_destroy(ctn_plan* plan);

/* Scratch the caller must provide to ctn_plan_execute (intermediates + the
 * per-execution state block); 256-byte aligned device memory. */
size_t ctn_plan_workspace_bytes(const ctn_plan* plan);

/* Shape of the result (C-contiguous). `extents` may be NULL to query the rank. */
int ctn_plan_output_rank(const ctn_plan* plan);
int ctn_plan_output_extents(const ctn_plan* plan, int64_t* extents);
int64_t ctn_plan_output_numel(const ctn_plan* plan);

/* Introspection used by the benchmark and the tests. */
int ctn_plan_num_launches(const ctn_plan* plan);      /* kernels per execute     */
double ctn_plan_flops(const ctn_plan* plan);          /* path FLOPs, opt_einsum convention */
double ctn_plan_step_bytes(const ctn_plan* plan);     /* sum over launches of operand+result bytes */
/* kind of launch i: 0 generic, 1 hadamard-chain, 2 interpreter, 3 gemm, 4 finalize,
 * 5 chain materialise pass, 6 split-K reduce pass; -1 past the end */
int ctn_plan_launch_kind(const ctn_plan* plan, int i);
/* One text line per launch (kernel class, GEMM view, tile configuration, physical strides of the
 * result) into buf[0..n): what the plan compiler decided, for tests and reports. */
int ctn_plan_describe(const ctn_plan* plan, char* buf, size_t n);

/* Run the plan: the body of _core_contract (contractn/einsum.py:341-393) plus
 * stabilize after every step (einsum.py:387, 89-107) and, for CTN_OUT_DENSE,
 * destabilize (einsum.py:310).  All pointers are DEVICE pointers.
 *   operands      [n_operands] device pointers (borrowed, never written)
 *   out           output tensor, ctn_plan_output_numel() elements
 *   log_scale_out one double: the log-scale register (contractn/einsum.py:338)
 *   flags_out     one int32 of CTN_RESULT_* (may be NULL)
 *   stream        cudaStream_t (NULL = default stream) */
int ctn_plan_execute(const ctn_plan* plan, const void* const* operands, void* out,
                     double* log_scale_out, int32_t* flags_out, void* workspace,
                     int out_mode, void* stream);

/* ctn_plan_execute with a CUDA event between consecutive kernels; synchronises
 * the stream and returns each kernel's device time in launch_ms[0 ..
 * ctn_plan_num_launches()).  Measurement aid for bench.py (per-kernel roofline). */
int ctn_plan_execute_timed(const ctn_plan* plan, const void* const* operands, void* out,
                           double* log_scale_out, int32_t* flags_out, void* workspace,
                           int out_mode, void* stream, float* launch_ms);

/* Same with HOST buffers: copies the operands to the device, runs the plan,
 * copies the result back and synchronises.  Device scratch is cached inside
 * the plan between calls; calls on the same plan are serialised by a mutex inside it.
 * This is the call a ctypes/cgo/JNI binding over host arrays makes. */
int ctn_plan_execute_host(ctn_plan* plan, const void* const* operands_host, void* out_host,
                          double* log_scale_host, int32_t* flags_host, int out_mode);

/* Batch-sharded evaluations (one process per GPU): bring this rank's shard of the split-format result
 * to the GLOBAL scale, i.e. the reference's single scale for the unsharded tensor
 * (contractn/einsum.py:89-107 applied to the concatenation of the shards).
 *
 * ctn_merge_batch_shards_peer does the whole merge in ONE kernel and without a host-launched
 * collective: `workspace` is the workspace of the ctn_plan_execute call that produced `data` (its
 * state block holds this shard's log-scale and sum|T_shard|); `mailboxes[i]` is rank i's mailbox,
 * CTN_MAILBOX_WORDS doubles of ZEROED device memory mapped into this process (peer access over
 * NVLink; mailboxes[rank] is local).  The kernel posts this rank's (log-scale, sum|T|, numel)
 * triple into every peer's mailbox, waits for all `world` triples of call number `seq` (1, 2, 3 ...
 * -- the same on every rank, incremented per call), and rescales `data` in place by
 * exp(c_rank - c*) / kappa, writing the common log-scale c* + ln(kappa) to *log_scale_out, where
 * c* = max_i c_i and kappa = sum_i(abs_i * exp(c_i - c*)) / sum_i numel_i.  world <= 64.
 *
 * ctn_merge_batch_shards is the second half alone, for callers that gathered the [world][3] table
 * of triples themselves (e.g. with an all-gather). */
#define CTN_MAILBOX_WORDS 512
int ctn_merge_batch_shards_peer(void* data, int64_t numel, int dtype, const void* workspace,
                                double* const* mailboxes, int world, int rank, uint64_t seq,
                                double* log_scale_out, void* stream);
int ctn_merge_batch_shards(void* data, int64_t numel, int dtype, const double* table, int world, int rank,
                           double* log_scale_out, void* stream);

/* Byte offset, inside the workspace, of the double holding the pending scale kappa of a
 * CTN_PLAN_RAW_FINAL plan's output; -1 when the output is already canonical (flag absent, or the last
 * launch finalises by itself). */
int64_t ctn_plan_final_kappa_offset(const ctn_plan* plan);

/* Boundary-sharded sweep (contractn_b200/sharded.py; no reference counterpart: the reference is
 * single-process): put this rank's slab on the common scale AND re-shard it across the ranks in ONE
 * kernel over peer memory.  `data` is the slab as the producing plan left it, laid out
 * [world][numel/world] with chunk j destined for rank j; `recv[i]` is rank i's receive buffer
 * (numel elements, peer-mapped); chunk j is stored to recv[j] + rank * (numel/world), multiplied on
 * the way by exp(c_rank - c*) / kappa (the ctn_merge_batch_shards_peer scale) and, when
 * kappa_offset >= 0, by 1 / kappa_raw read from workspace + kappa_offset (a CTN_PLAN_RAW_FINAL
 * plan).  `done_boxes[i]` is a second zeroed mailbox per rank (CTN_MAILBOX_WORDS doubles) through
 * which the ranks tell one another that their stores have landed: when the kernel has completed on
 * this rank, every peer's chunk is in recv[rank].  `ticket` is one zeroed uint32 in local device memory
 * (left zero).  `seq` is shared with ctn_merge_batch_shards_peer (one counter per group). */
int ctn_reshard_peer(const void* data, int64_t numel, int dtype, const void* workspace, int64_t kappa_offset,
                     double* const* mailboxes, double* const* done_boxes, void* const* recv, uint32_t* ticket,
                     int world, int rank, uint64_t seq, double* log_scale_out, void* stream);

const char* ctn_status_string(int status);
const char* ctn_last_cuda_error(void);
int ctn_abi_version(void);
/* Number of visible CUDA devices with compute capability 10.x (0 = none). */
int ctn_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CTN_B200_H */
