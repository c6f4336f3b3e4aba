// The counting half of accuracy() on the device: landmark_regression/lib/core/evaluate.py:16-39
// (calc_dists, dist_acc) over the argmax coordinates that spe_max_preds_f32 leaves in HBM for the network output and for
// the target heatmaps, so that the training / validation loops (lib/core/function.py:61-62, :395-396) no longer copy
// both heatmap tensors to the host for their PCK number.
//
// Per (frame, joint): counted only if both target coordinates are > 1; distance = | pred / norm - target / norm |_2 in
// float64 with norm = (H / 10, W / 10) applied to (x, y) — the reference's order — every product and sum rounded
// separately (NumPy does not contract), then `< thr`.  One CTA per joint, no atomics: counts[j] = (valid, below).
#include <cuda_runtime.h>
#include <stdint.h>

#include "evaluate.cuh"

namespace spe {

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) pck_counts_kernel(const float* __restrict__ pred, const float* __restrict__ target, int B, int J, double norm_x,
                                                              double norm_y, double thr, int32_t* __restrict__ counts) {
  const int j = blockIdx.x;
  int valid = 0, below = 0;
  for (int n = threadIdx.x; n < B; n += kThreads) {
    const float2 p = reinterpret_cast<const float2*>(pred)[(size_t)n * J + j];
    const float2 t = reinterpret_cast<const float2*>(target)[(size_t)n * J + j];
    if (t.x > 1.0f && t.y > 1.0f) {
      const double dx = (double)p.x / norm_x - (double)t.x / norm_x, dy = (double)p.y / norm_y - (double)t.y / norm_y;
      const double dist = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      ++valid;
      below += dist < thr ? 1 : 0;
    }
  }
  __shared__ int s_valid[kThreads / 32], s_below[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    valid += __shfl_xor_sync(0xffffffffu, valid, o);
    below += __shfl_xor_sync(0xffffffffu, below, o);
  }
  if ((threadIdx.x & 31) == 0) s_valid[threadIdx.x >> 5] = valid, s_below[threadIdx.x >> 5] = below;
  __syncthreads();
  if (threadIdx.x == 0) {
    int v = 0, b = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) v += s_valid[w], b += s_below[w];
    counts[2 * j] = v;
    counts[2 * j + 1] = b;
  }
}

}  // namespace

cudaError_t launch_pck_counts(const float* pred, const float* target, int B, int J, double norm_x, double norm_y, double thr, int32_t* counts,
                              cudaStream_t stream) {
  pck_counts_kernel<<<J, kThreads, 0, stream>>>(pred, target, B, J, norm_x, norm_y, thr, counts);
  return cudaGetLastError();
}

}  // namespace spe
