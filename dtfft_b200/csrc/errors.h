// dtfft_error_t values (reference: include/dtfft_config.h.in:82-151) come from the public
// header; this file adds the mapping of CUDA failures to return codes.
#pragma once
#include <cuda_runtime.h>

#include "../../include/dtfft_b200_api.h"

// names used by the Fortran side of the reference (CONF_ macros) that differ from the C enum
#define DTFFT_ERROR_INVALID_EFFORT_FLAG DTFFT_ERROR_INVALID_EFFORT
#define DTFFT_ERROR_INVALID_EXECUTOR_TYPE DTFFT_ERROR_INVALID_EXECUTOR

namespace dtfftb {
inline int cuda_error(cudaError_t e) { return e == cudaSuccess ? 0 : DTFFTB_ERROR_CUDA_BASE - (int)e; }
}  // namespace dtfftb
