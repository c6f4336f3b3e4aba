// Internal interface between the C ABI (capi.cu) and the RANSAC-EPnP kernels (ransac_epnp.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace spe {

constexpr int kMaxLandmarks = 32;
constexpr int kModelPoints = 5;  // EPnP minimal set used by cv2.solvePnPRansac
// one control-point table entry: alpha[5][3] (points in ascending landmark order; alpha_k0 = 1 - sum),
// k_i^2 [3] (squared control-point distances from the centroid: they give rho), pad[2].  The control
// points themselves are not needed by the hypothesis kernel.
constexpr int kCtrlEntryFloats = 20;

struct Camera {
  double fx, fy, cx, cy;
  double k1, k2, p1, p2, k3;
};

// Host-side model: immutable after creation.
struct Model {
  int J = 0;
  int max_hyp = 0;
  int device = 0;
  Camera cam{};
  float landmarks_f32[kMaxLandmarks * 3] = {};  // cv2 rounds object points to float32 on entry
  float* d_landmarks = nullptr;                 // [J,3] float32
  uint8_t* d_subsets = nullptr;                 // [J-5][max_hyp][5]: minimal sets for n = 6..J
  std::vector<uint8_t> h_subsets;               // host copy of the same table
  // Control points / barycentric coordinates of EVERY 5-subset of the J landmarks (they depend on the
  // object points only): C(J,5) entries of kCtrlEntryFloats floats, indexed by the combinatorial rank
  // of the sorted landmark ids (ransac_epnp.cu: build_control_table / hypothesis_kernel_t1).
  float* d_ctrl = nullptr;
  size_t ctrl_entries = 0;
};

// Workspace carve-up for B frames x H hypotheses (all offsets 16-byte aligned).
struct RansacWorkspace {
  double2* und;      // [B,J] undistorted normalised coordinates (float64, for the final refit)
  float2* us_hyp;    // [B,J] ideal pixel coordinates as float32 (for the hypotheses)
  float2* img;       // [B,J] raw (distorted) pixel coordinates, for the optional LM refinement
  int32_t* n;        // [B] number of landmarks that passed the confidence filter
  uint32_t* vis;     // [B] bit j = landmark j takes part
  uint32_t* masks;   // [B,H] inlier mask of every hypothesis over the J landmarks
  uint8_t* counts;   // [B,H] popcount of the above
  int32_t* need;     // [B] adaptive mode: hypotheses cv2 could still look at after the first pass
  int frames;        // B
  size_t bytes;
};

size_t ransac_workspace_bytes(int J, int B, int H);
RansacWorkspace carve_workspace(void* base, int J, int B, int H);

struct RansacArgs {
  const float* kpts;  // [B,J,3] (x, y, conf)
  int B, H;
  float reproj_err;
  double confidence;
  float conf_floor;  // < 0: the reference's adaptive filter
  int jacobi_sweeps;  // Jacobi sweeps (variants 1, 2) / inverse-iteration steps (variant 0)
  int refine_lm;              // SPE_FLAG_REFINE_LM
  int adaptive;               // SPE_FLAG_ADAPTIVE: score only the hypotheses cv2 could look at
  int refit_background;       // the tail runs under other kernels: keep its shared-memory footprint at zero
  int tail_warps;             // warps per CTA of the background select/refit kernel (1..8)
  int t1_warps;               // warps per CTA of the thread-per-hypothesis kernel (1..4)
  int kernel_variant;  // 0: thread per hypothesis, QR + inverse iteration (default); 1: 4 lanes per hypothesis; 2: thread per hypothesis, Jacobi SVD
  float* pose7;           // [B,7]
  uint32_t* inlier_mask;  // [B]
  int32_t* status;        // [B]
  int32_t* winner;        // [B] or nullptr
  double* rt;             // [B,12] or nullptr
};

cudaError_t model_upload(Model& m);
void model_free(Model& m);
cudaError_t launch_ransac_epnp(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
// the two halves of the above: frame prep + hypothesis scoring, then selection + float64 refit
cudaError_t launch_ransac_score(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
cudaError_t launch_ransac_select_refit(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
cudaError_t launch_debug_scores(const RansacWorkspace& ws, int B, int H, int32_t* counts, uint32_t* masks, cudaStream_t stream);

// Host side of the control-point table (ransac_epnp.cu): one entry, and the rank of a sorted 5-subset.
void control_table_entry(const float* landmarks_f32, const int (&ids)[5], float* entry /* [kCtrlEntryFloats] */);
size_t control_table_rank(const int (&sorted_ids)[5]);

// OpenCV's RANSAC RNG: minimal sets for `count` points, draw order preserved.
void opencv_minimal_sets(int count, int num, uint8_t* out /* [num][5] */);

}  // namespace spe
