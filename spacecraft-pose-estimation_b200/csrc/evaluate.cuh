// Internal interface between the C ABI (capi.cu) and the accuracy() kernel (evaluate.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spe {

cudaError_t launch_pck_counts(const float* pred, const float* target, int B, int J, double norm_x, double norm_y, double thr, int32_t* counts,
                              cudaStream_t stream);

}  // namespace spe
