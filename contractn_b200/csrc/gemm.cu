// Tiled tensor-core GEMM kernels for pairwise steps (placeholder: filled in next).
#include "ctn_internal.h"

namespace ctn {

bool gemm_supported(const Plan&, const StepDev&, const GemmDev&) { return false; }

cudaError_t launch_gemm(const ExecCtx&, const Launch&) { return cudaErrorNotSupported; }

}  // namespace ctn
