// Internal declarations shared by the GEMM translation units.
#pragma once
#include <cuda_runtime.h>

#include "../../include/videoblip_b200.h"

namespace vb {
bool gemm_tcgen05_eligible(const vb_gemm_args& a);
cudaError_t gemm_tcgen05_launch(const vb_gemm_args& a, cudaStream_t stream);
cudaError_t gemm_generic_launch(const vb_gemm_args& a, cudaStream_t stream);
}  // namespace vb
