// Internal interface between the C ABI (capi.cu) and the RANSAC-EPnP kernels:
//   ransac_model.cu  host side: OpenCV's minimal sets, duplicate-free hypothesis lists, control-point table, workspace
//   ransac_score.cu  frame preparation + the FP32 hypothesis kernel (every distinct minimal set scored once)
//   ransac_exact.cu  float64 replay of cv2's sequential RANSAC loop (SPE_FLAG_EXACT)
//   ransac_refit.cu  selection over the FP32 scores (fast mode) + the float64 EPnP refit on the winner's inliers
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
// float64 replay (ransac_exact.cu): phases of width 8, 32, 128, 512, 2048, 2048, ...
constexpr int kReplayMaxWidth = 2048;
constexpr int kReplayMaxPhases = 16;
constexpr int kReplayPlanMax = 32;      // widest phase 0 (hypothesis index fits 5 bits of an item)
constexpr int kReplayStateBytes = 32;
constexpr int kReplayFrameBytes = kMaxLandmarks * (3 * 8 + 2 * 8 + 2 * 4 + 1);  // sizeof(FramePoints)
constexpr int kClaimActive = 0, kClaimPlanItems = kReplayMaxPhases + 1, kClaimWords = kReplayMaxPhases + 3;

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
  uint8_t* d_subsets = nullptr;                 // [J-5][max_hyp][5]: minimal sets for n = 6..J, draw order
  std::vector<uint8_t> h_subsets;               // host copy of the same table
  // Duplicate-free hypothesis lists.  OpenCV's RNG draws the same 5-subset (as a set) again and again: for n = 11 only
  // C(11,5) = 462 sets exist, and 25 % of the first 256 draws, 60 % of the first 1024 repeat an earlier one.  The FP32
  // kernel handles the five points in ascending landmark order, so a repeated set is bit-identical work: it is scored
  // once.  Per n: uniq[u] = draw index of the first occurrence of the u-th distinct set (ascending), slot[h] = u of the
  // set hypothesis h draws.  The distinct sets among the first H draws are exactly the prefix uniq[0 .. U(n,H)).
  uint16_t* d_uniq = nullptr;  // [J-5][max_hyp]
  uint16_t* d_slot = nullptr;  // [J-5][max_hyp]
  std::vector<uint16_t> h_uniq, h_slot;
  std::vector<int> h_uniq_count;  // [J-5]: number of distinct sets among all max_hyp draws
  // Control points / barycentric coordinates of EVERY 5-subset of the J landmarks (they depend on the
  // object points only): C(J,5) entries of kCtrlEntryFloats floats, indexed by the combinatorial rank
  // of the sorted landmark ids (ransac_model.cu: build_control_table / ransac_score.cu: hypothesis_kernel_t1).
  float* d_ctrl = nullptr;
  size_t ctrl_entries = 0;
};

// number of distinct minimal sets among the first H draws for n points (host)
int unique_sets(const Model& m, int n, int H);
// duplicate-free lists of `num` draws: fills uniq[num] (first-occurrence draw index per distinct set, ascending, 0xffff-padded)
// and slot[num] (distinct-set number of every draw); returns the number of distinct sets
int build_unique(const uint8_t* subsets, int num, uint16_t* uniq, uint16_t* slot);

// What the kernels see of the model (passed by value).
struct DevModel {
  const float* landmarks;   // [J,3]
  const uint8_t* subsets;   // [J-5][max_hyp][5]
  const uint16_t* uniq;     // [J-5][max_hyp]
  const uint16_t* slot;     // [J-5][max_hyp]
  int J, max_hyp;
  Camera cam;
  const float4* ctrl;  // [C(J,5)][kCtrlEntryFloats / 4] control-point table
};
inline DevModel dev_model(const Model& m) {
  return DevModel{m.d_landmarks, m.d_subsets, m.d_uniq, m.d_slot, m.J, m.max_hyp, m.cam, reinterpret_cast<const float4*>(m.d_ctrl)};
}

// Workspace carve-up for B frames x H hypotheses (all offsets 16-byte aligned).
struct RansacWorkspace {
  double2* und;      // [B,J] undistorted normalised coordinates (float64, for the final refit)
  float2* us_hyp;    // [B,J] ideal pixel coordinates as float32 (for the hypotheses)
  float2* img;       // [B,J] raw (distorted) pixel coordinates, for the optional LM refinement
  int32_t* n;        // [B] number of landmarks that passed the confidence filter
  uint32_t* vis;     // [B] bit j = landmark j takes part
  uint32_t* masks;   // [B,H] inlier mask of every DISTINCT minimal set, at its slot (see Model::d_slot)
  uint8_t* counts;   // [B,H] popcount of the above
  int32_t* need;     // [B] adaptive mode: hypotheses cv2 could still look at after the first pass
  // float64 replay (SPE_FLAG_EXACT): what cv2's own loop ends with
  int32_t* x_winner;   // [B] accepted hypothesis (-1 none)
  uint32_t* x_mask;    // [B] its inlier mask over the J landmarks
  int32_t* x_visited;  // [B] hypotheses cv2 evaluates before its budget runs out
  uint32_t* claim;     // [kClaimWords] per-phase work-list lengths and item counters, zeroed by the launcher on the call's stream
  void* x_state;       // [B] kReplayStateBytes per-frame state of cv2's loop between phases (ransac_exact.cu: ReplayState)
  void* x_frames;      // [B] compacted visible landmarks of every frame (ransac_exact_eval.cuh: FramePoints)
  int32_t* x_width;    // [B] hypotheses of phase 0 (the predicted length of cv2's loop)
  uint32_t* x_items;   // [B * kReplayPlanMax] phase-0 work list: frame << 5 | hypothesis
  uint32_t* x_done;    // [B] blocks of the current phase completed per frame
  int32_t* x_active;   // [2][B] work lists of frames still in cv2's loop (ping-pong between phases)
  uint32_t* x_masks;   // [B][kReplayMaxWidth] inlier masks of the current phase, one row per work-list entry
  int frames;          // B
  size_t bytes;
};

size_t ransac_workspace_bytes(int J, int B, int H);
RansacWorkspace carve_workspace(void* base, int J, int B, int H);

struct RansacArgs {
  const float* kpts;  // [B,J,3] (x, y, conf)
  int B, H;
  int iterations;  // cv2's iterationsCount for the float64 replay (<= max_hyp)
  float reproj_err;
  double confidence;
  float conf_floor;  // < 0: the reference's adaptive filter
  int eig_iters;     // inverse-iteration steps of the FP32 eigen stage (Jacobi sweeps for the dev variants)
  int refine_lm;              // SPE_FLAG_REFINE_LM
  int adaptive;               // SPE_FLAG_ADAPTIVE: score only the hypotheses cv2 could look at
  int exact;                  // SPE_FLAG_EXACT: selection = float64 replay of cv2's loop
  int refit_background;       // the tail runs under other kernels: whole-SM CTAs on few SMs
  int tail_warps;             // dev knob: warps per CTA of the background select/refit kernel (0 = default)
  int t1_warps;               // dev knob: warps per CTA of the thread-per-hypothesis kernel (0 = default)
  int kernel_variant;  // 0: thread per hypothesis, QR + inverse iteration; dev builds: 1 = 4 lanes per hypothesis, 2 = Jacobi SVD
  float* pose7;           // [B,7]
  uint32_t* inlier_mask;  // [B]
  int32_t* status;        // [B]
  int32_t* winner;        // [B] or nullptr
  double* rt;             // [B,12] or nullptr
  int32_t* budget;        // [B] or nullptr: cv2's iteration budget when its loop ends (fast mode: capped scan; exact: visited)
};

cudaError_t model_upload(Model& m);
void model_free(Model& m);
// frame prep + FP32 hypothesis scoring; float64 replay; selection + float64 refit
cudaError_t launch_ransac_score(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
cudaError_t launch_frame_prep(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
cudaError_t launch_ransac_replay(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
cudaError_t launch_ransac_select_refit(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream);
cudaError_t launch_debug_scores(const Model& m, const RansacWorkspace& ws, int B, int H, int32_t* counts, uint32_t* masks, cudaStream_t stream);

// Host side of the control-point table (ransac_model.cu): one entry, and the rank of a sorted 5-subset.
void control_table_entry(const float* landmarks_f32, const int (&ids)[5], float* entry /* [kCtrlEntryFloats] */);
size_t control_table_rank(const int (&sorted_ids)[5]);

// OpenCV's RANSAC RNG: minimal sets for `count` points, draw order preserved.
void opencv_minimal_sets(int count, int num, uint8_t* out /* [num][5] */);

}  // namespace spe
