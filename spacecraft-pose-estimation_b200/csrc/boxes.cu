// Detection boxes -> what the decode consumes, on the device: the last host hop between the
// detectron2 boxes and get_final_preds.
//
//   pick_boxes_kernel   object_detection/export_object_detection_bounding_boxes.py:313-329: an image with exactly one or
//                       two detections keeps the one with the highest score (np.argmax: first maximum, a NaN wins);
//                       any other count falls back to the whole image with score 0.  The box leaves as COCO
//                       [x, y, w, h] in float64 — the script converts the float32 corners with .tolist() before
//                       subtracting, and json keeps float64 exactly.
//   xywh2cs_kernel      landmark_regression/lib/dataset/PEdataset.py:98-113 (_xywh2cs, pixel_std = 200): center =
//                       float32(x + w * 0.5), scale = float32(w / 200) * 1.5 evaluated in float32 (a float32 array times
//                       a Python float stays float32), the factor 1.5 skipped when center[0] == -1.
// One thread per image; nothing here is worth more (a few bytes per image).
#include <cuda_runtime.h>
#include <stdint.h>

#include "boxes.cuh"

namespace spe {

namespace {

constexpr double kPixelStd = 200.0;  // lib/dataset/PEdataset.py:41

__device__ __forceinline__ void xywh_to_center_scale(double x, double y, double w, double h, float* center, float* scale) {
  const float cx = (float)(x + w * 0.5), cy = (float)(y + h * 0.5);
  float sx = (float)(w * 1.0 / kPixelStd), sy = (float)(h * 1.0 / kPixelStd);
  if (cx != -1.0f) {
    sx = __fmul_rn(sx, 1.5f);
    sy = __fmul_rn(sy, 1.5f);
  }
  center[0] = cx, center[1] = cy;
  scale[0] = sx, scale[1] = sy;
}

__global__ void xywh2cs_kernel(const double* __restrict__ xywh, int B, float* __restrict__ center, float* __restrict__ scale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double* q = xywh + (size_t)b * 4;
  xywh_to_center_scale(q[0], q[1], q[2], q[3], center + (size_t)b * 2, scale + (size_t)b * 2);
}

__global__ void pick_boxes_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int32_t* __restrict__ counts, int B, int K,
                                  double image_w, double image_h, double* __restrict__ xywh, float* __restrict__ best_score, int32_t* __restrict__ best_index,
                                  float* __restrict__ center, float* __restrict__ scale) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = counts ? counts[b] : K;
  double x1 = 0.0, y1 = 0.0, x2 = image_w, y2 = image_h;
  float score = 0.f;
  int pick = -1;
  if (n == 1 || n == 2) {
    const float* s = scores + (size_t)b * K;
    pick = 0;
    score = s[0];
    // np.argmax: the first maximum; a NaN compares as the maximum
    if (n == 2 && !(score != score) && (s[1] > score || s[1] != s[1])) pick = 1, score = s[1];
    const float* q = boxes + ((size_t)b * K + pick) * 4;
    x1 = (double)q[0], y1 = (double)q[1], x2 = (double)q[2], y2 = (double)q[3];
  }
  const double w = x2 - x1, h = y2 - y1;
  if (xywh) {
    double* o = xywh + (size_t)b * 4;
    o[0] = x1, o[1] = y1, o[2] = w, o[3] = h;
  }
  if (best_score) best_score[b] = score;
  if (best_index) best_index[b] = pick;
  if (center && scale) xywh_to_center_scale(x1, y1, w, h, center + (size_t)b * 2, scale + (size_t)b * 2);
}

}  // namespace

cudaError_t launch_xywh2cs(const double* xywh, int B, float* center, float* scale, cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  xywh2cs_kernel<<<(B + 127) / 128, 128, 0, stream>>>(xywh, B, center, scale);
  return cudaGetLastError();
}

cudaError_t launch_pick_boxes(const float* boxes, const float* scores, const int32_t* counts, int B, int K, double image_w, double image_h,
                              double* xywh, float* best_score, int32_t* best_index, float* center, float* scale, cudaStream_t stream) {
  if (B == 0) return cudaSuccess;
  pick_boxes_kernel<<<(B + 127) / 128, 128, 0, stream>>>(boxes, scores, counts, B, K, image_w, image_h, xywh, best_score, best_index, center, scale);
  return cudaGetLastError();
}

}  // namespace spe
