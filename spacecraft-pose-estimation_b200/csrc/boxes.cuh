// Internal interface between the C ABI (capi.cu) and the box kernels (boxes.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spe {

cudaError_t launch_xywh2cs(const double* xywh, int B, float* center, float* scale, cudaStream_t stream);
cudaError_t launch_pick_boxes(const float* boxes, const float* scores, const int32_t* counts, int B, int K, double image_w, double image_h,
                              double* xywh, float* best_score, int32_t* best_index, float* center, float* scale, cudaStream_t stream);

}  // namespace spe
