// Internal interface between the C ABI (capi.cu) and the decode kernels (decode.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spe {

// Shared-memory carveout preferred by every kernel of the stage (percent of 228 KB -> 132 KB).
#ifndef SPE_SMEM_CARVEOUT
#define SPE_SMEM_CARVEOUT 58
#endif
constexpr int kSmemCarveoutPct = SPE_SMEM_CARVEOUT;

struct DecodeArgs {
  const float* hm;      // [n_maps, H, W]
  int n_maps, J, H, W;  // n_maps = B * J
  const float* center;  // [B,2] or nullptr (get_max_preds: no affine, no refine)
  const float* scale;   // [B,2]
  int post_process;
  float* preds;     // [n_maps,2]   (used when kpts == nullptr)
  float* maxvals;   // [n_maps]
  float* kpts;      // [n_maps,3]   (x, y, maxval) or nullptr
  int32_t* argmax;  // [n_maps] or nullptr
  int background;   // SPE_DECODE_BACKGROUND: small CTAs meant to run UNDER a compute-bound kernel of another stream
};

cudaError_t launch_decode(const DecodeArgs& a, cudaStream_t stream);

// decode of a combination of heatmap tensors that is never materialised.
constexpr int kMaxCombine = 8;
enum CombineMode { kCombineMean = 0, kCombineFlip = 1 };
struct CombineArgs {
  const float* src[kMaxCombine];  // K tensors [n_maps, H, W]
  int K, mode;
  const int32_t* flip_perm;  // DEVICE [J]: joint whose flipped map lands on joint j (nullptr = identity)
  int shift_heatmap;
  DecodeArgs out;  // hm unused; everything else as for launch_decode
};
cudaError_t launch_decode_combined(const CombineArgs& a, cudaStream_t stream);

}  // namespace spe
