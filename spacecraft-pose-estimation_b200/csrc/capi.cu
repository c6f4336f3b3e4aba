// C ABI of the heatmap -> pose stage (include/spe_b200.h).  Argument checking and launches only.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/spe_b200.h"
#include "boxes.cuh"
#include "evaluate.cuh"
#include "decode.cuh"
#include "ransac.cuh"

#include <new>
#include <stdexcept>

namespace {

thread_local cudaError_t t_last_cuda = cudaSuccess;

int cuda_fail(cudaError_t e) {
  t_last_cuda = e;
  return SPE_ERR_CUDA;
}

int run_decode(const float* hm, int B, int J, int H, int W, const float* center, const float* scale, int post_process,
               float* preds, float* maxvals, float* kpts, int32_t* argmax, void* stream, int background = 0) {
  if (B < 0 || J <= 0 || H <= 0 || W <= 0) return SPE_ERR_INVALID_ARGUMENT;
  if ((long long)H * W > 0x7fffffffLL || (long long)B * J > 0x7fffffffLL) return SPE_ERR_INVALID_ARGUMENT;
  if (B == 0) return SPE_OK;
  if (hm == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  if (kpts == nullptr && (preds == nullptr || maxvals == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  if ((center == nullptr) != (scale == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  spe::DecodeArgs a{};
  a.hm = hm;
  a.n_maps = B * J;
  a.J = J;
  a.H = H;
  a.W = W;
  a.center = center;
  a.scale = scale;
  a.post_process = post_process;
  a.preds = preds;
  a.maxvals = maxvals;
  a.kpts = kpts;
  a.argmax = argmax;
  a.background = background;
  const cudaError_t e = spe::launch_decode(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

}  // namespace

extern "C" {

int spe_abi_version(void) { return SPE_ABI_VERSION; }

const char* spe_status_string(int status) {
  switch (status) {
    case SPE_OK: return "ok";
    case SPE_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SPE_ERR_CUDA: return "CUDA runtime error";
    case SPE_ERR_WORKSPACE: return "workspace too small or misaligned";
    case SPE_ERR_UNSUPPORTED: return "unsupported configuration";
    case SPE_ERR_OUT_OF_MEMORY: return "out of host memory";
    default: return "unknown status";
  }
}

const char* spe_last_cuda_error(void) { return cudaGetErrorString(t_last_cuda); }

int spe_max_preds_f32(const float* hm, int B, int J, int H, int W, float* preds, float* maxvals, int32_t* argmax, void* stream) {
  return run_decode(hm, B, J, H, W, nullptr, nullptr, 0, preds, maxvals, nullptr, argmax, stream);
}

int spe_decode_f32(const float* hm, int B, int J, int H, int W, const float* center, const float* scale, int post_process,
                   float* preds, float* maxvals, int32_t* argmax, void* stream) {
  if (center == nullptr || scale == nullptr) return B == 0 ? SPE_OK : SPE_ERR_INVALID_ARGUMENT;
  return run_decode(hm, B, J, H, W, center, scale, post_process, preds, maxvals, nullptr, argmax, stream);
}

int spe_decode_kpts_f32(const float* hm, int B, int J, int H, int W, const float* center, const float* scale, int post_process,
                        float* kpts, int32_t* argmax, void* stream) {
  return spe_decode_kpts_ex_f32(hm, B, J, H, W, center, scale, post_process, kpts, argmax, 0, stream);
}

int spe_decode_kpts_ex_f32(const float* hm, int B, int J, int H, int W, const float* center, const float* scale, int post_process,
                           float* kpts, int32_t* argmax, int flags, void* stream) {
  if (center == nullptr || scale == nullptr || kpts == nullptr) return B == 0 ? SPE_OK : SPE_ERR_INVALID_ARGUMENT;
  return run_decode(hm, B, J, H, W, center, scale, post_process, nullptr, nullptr, kpts, argmax, stream, (flags & SPE_DECODE_BACKGROUND) ? 1 : 0);
}

int spe_decode_combined_kpts_f32(const float* const* srcs, int K, int mode, const int32_t* flip_perm, int shift_heatmap, int B, int J, int H, int W,
                                 const float* center, const float* scale, int post_process, float* kpts, int32_t* argmax, void* stream) {
  if (B < 0 || J <= 0 || H <= 0 || W <= 0 || srcs == nullptr || K < 1 || K > spe::kMaxCombine) return SPE_ERR_INVALID_ARGUMENT;
  if (mode != SPE_COMBINE_MEAN && mode != SPE_COMBINE_FLIP) return SPE_ERR_INVALID_ARGUMENT;
  if (mode == SPE_COMBINE_FLIP && K != 2) return SPE_ERR_INVALID_ARGUMENT;
  if ((long long)H * W > 0x7fffffffLL || (long long)B * J > 0x7fffffffLL) return SPE_ERR_INVALID_ARGUMENT;
  if (B == 0) return SPE_OK;
  if (center == nullptr || scale == nullptr || kpts == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  spe::CombineArgs a{};
  for (int k = 0; k < K; ++k) {
    if (srcs[k] == nullptr) return SPE_ERR_INVALID_ARGUMENT;
    a.src[k] = srcs[k];
  }
  a.K = K;
  a.mode = mode == SPE_COMBINE_MEAN ? spe::kCombineMean : spe::kCombineFlip;
  a.flip_perm = flip_perm;
  a.shift_heatmap = shift_heatmap;
  a.out.n_maps = B * J;
  a.out.J = J;
  a.out.H = H;
  a.out.W = W;
  a.out.center = center;
  a.out.scale = scale;
  a.out.post_process = post_process;
  a.out.kpts = kpts;
  a.out.argmax = argmax;
  const cudaError_t e = spe::launch_decode_combined(a, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

// ---- detection boxes -> (center, scale) --------------------------------------------------------
int spe_boxes_to_center_scale_f64(const double* xywh, int B, float* center, float* scale, void* stream) {
  if (B < 0) return SPE_ERR_INVALID_ARGUMENT;
  if (B == 0) return SPE_OK;
  if (xywh == nullptr || center == nullptr || scale == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  const cudaError_t e = spe::launch_xywh2cs(xywh, B, center, scale, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_pick_boxes_f32(const float* boxes, const float* scores, const int32_t* counts, int B, int K, double image_w, double image_h, double* xywh,
                       float* best_score, int32_t* best_index, float* center, float* scale, void* stream) {
  if (B < 0 || K < 0) return SPE_ERR_INVALID_ARGUMENT;
  if (B == 0) return SPE_OK;
  if ((center == nullptr) != (scale == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  if (K > 0 && (boxes == nullptr || scores == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  if (K == 0 && counts != nullptr) return SPE_ERR_INVALID_ARGUMENT;
  const cudaError_t e = spe::launch_pick_boxes(boxes, scores, counts, B, K, image_w, image_h, xywh, best_score, best_index, center, scale,
                                               static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

// ---- accuracy() ------------------------------------------------------------------------------
int spe_pck_counts_f32(const float* pred, const float* target, int B, int J, double norm_x, double norm_y, double thr, int32_t* counts, void* stream) {
  if (B < 0 || J <= 0 || !(norm_x > 0.0) || !(norm_y > 0.0) || counts == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  if (B > 0 && (pred == nullptr || target == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  // the kernel reads (x, y) pairs as float2
  if ((reinterpret_cast<uintptr_t>(pred) & 7u) || (reinterpret_cast<uintptr_t>(target) & 7u)) return SPE_ERR_INVALID_ARGUMENT;
  const cudaError_t e = spe::launch_pck_counts(pred, target, B, J, norm_x, norm_y, thr, counts, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

// ---- pose ----------------------------------------------------------------------------------
struct spe_model {
  spe::Model m;
};

int spe_pnp_model_create(const double* landmarks, int J, const double* K, const double* dist, int max_hypotheses, spe_model_t** out) {
  if (out == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (landmarks == nullptr || K == nullptr || J < 4 || J > SPE_MAX_LANDMARKS) return SPE_ERR_INVALID_ARGUMENT;
  if (max_hypotheses < 1 || max_hypotheses > SPE_MAX_HYPOTHESES) return SPE_ERR_INVALID_ARGUMENT;
  // the pinhole + (k1,k2,p1,p2,k3) model the reference's calibration.json holds: no skew, a plain last row
  if (K[1] != 0.0 || K[3] != 0.0 || K[6] != 0.0 || K[7] != 0.0 || K[8] != 1.0 || !(K[0] > 0.0) || !(K[4] > 0.0)) return SPE_ERR_UNSUPPORTED;
  spe_model* h = new (std::nothrow) spe_model();
  if (h == nullptr) return SPE_ERR_OUT_OF_MEMORY;
  spe::Model& m = h->m;
  m.J = J;
  m.max_hyp = max_hypotheses;
  for (int i = 0; i < 3 * J; ++i) m.landmarks_f32[i] = (float)landmarks[i];
  m.cam.fx = K[0], m.cam.cx = K[2], m.cam.fy = K[4], m.cam.cy = K[5];
  m.cam.k1 = dist ? dist[0] : 0.0, m.cam.k2 = dist ? dist[1] : 0.0, m.cam.p1 = dist ? dist[2] : 0.0;
  m.cam.p2 = dist ? dist[3] : 0.0, m.cam.k3 = dist ? dist[4] : 0.0;
  cudaError_t e = cudaSuccess;
  try {  // the tables are std::vectors (the control-point table alone is 16 MB for J = 32): nothing may throw across the ABI
    e = spe::model_upload(m);
  } catch (const std::bad_alloc&) {
    spe::model_free(m);
    delete h;
    return SPE_ERR_OUT_OF_MEMORY;
  } catch (...) {
    spe::model_free(m);
    delete h;
    return SPE_ERR_UNSUPPORTED;
  }
  if (e != cudaSuccess) {
    spe::model_free(m);
    delete h;
    return cuda_fail(e);
  }
  *out = h;
  return SPE_OK;
}

int spe_pnp_model_destroy(spe_model_t* model) {
  if (model == nullptr) return SPE_OK;
  spe::model_free(model->m);
  delete model;
  return SPE_OK;
}

int spe_pnp_model_num_landmarks(const spe_model_t* model) { return model ? model->m.J : SPE_ERR_INVALID_ARGUMENT; }

int spe_pnp_model_minimal_sets(const spe_model_t* model, int n, int count, int32_t* out) {
  if (model == nullptr || out == nullptr || n < 6 || n > model->m.J || count < 0 || count > model->m.max_hyp) return SPE_ERR_INVALID_ARGUMENT;
  const uint8_t* t = model->m.h_subsets.data() + (size_t)(n - 6) * model->m.max_hyp * spe::kModelPoints;
  for (int i = 0; i < count * spe::kModelPoints; ++i) out[i] = t[i];
  return SPE_OK;
}

int spe_pnp_minimal_sets_host(int n, int count, int32_t* sets, int32_t* slot, int32_t* uniq, int32_t* num_unique) {
  if (n < 6 || n > SPE_MAX_LANDMARKS || count < 1 || count > SPE_MAX_HYPOTHESES) return SPE_ERR_INVALID_ARGUMENT;
  try {
    std::vector<uint8_t> sub((size_t)count * spe::kModelPoints);
    std::vector<uint16_t> u(count), s(count);
    spe::opencv_minimal_sets(n, count, sub.data());
    const int nu = spe::build_unique(sub.data(), count, u.data(), s.data());
    for (int i = 0; i < count; ++i) {
      if (sets)
        for (int k = 0; k < spe::kModelPoints; ++k) sets[i * spe::kModelPoints + k] = sub[(size_t)i * spe::kModelPoints + k];
      if (slot) slot[i] = s[i];
      if (uniq) uniq[i] = i < nu ? (int32_t)u[i] : -1;
    }
    if (num_unique) *num_unique = nu;
  } catch (...) {
    return SPE_ERR_OUT_OF_MEMORY;
  }
  return SPE_OK;
}

int spe_pnp_control_entry(const double* landmarks, int J, const int32_t* ids, float* entry, int64_t* rank) {
  if (landmarks == nullptr || ids == nullptr || entry == nullptr || J < 5 || J > SPE_MAX_LANDMARKS) return SPE_ERR_INVALID_ARGUMENT;
  int sorted[5];
  for (int k = 0; k < 5; ++k) {
    if (ids[k] < 0 || ids[k] >= J || (k > 0 && ids[k] <= ids[k - 1])) return SPE_ERR_INVALID_ARGUMENT;  // ascending, distinct
    sorted[k] = ids[k];
  }
  float lm[SPE_MAX_LANDMARKS * 3];
  for (int i = 0; i < 3 * J; ++i) lm[i] = (float)landmarks[i];
  spe::control_table_entry(lm, sorted, entry);
  if (rank) *rank = (int64_t)spe::control_table_rank(sorted);
  return SPE_OK;
}

#ifdef SPE_DEV
// Development builds only (-DSPE_DEV): environment knobs for A/B runs.  The shipped library reads no environment
// variables and keeps no mutable global state.
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static void dev_knobs(spe::RansacArgs& a) {
  static const int variant = [] {
    const char* v = getenv("SPE_HYP_KERNEL");  // "g4" = 4 lanes per hypothesis, "jacobi" = full one-sided Jacobi SVD of M^T
    return (v && v[0] == 'g') ? 1 : (v && v[0] == 'j') ? 2 : 0;
  }();
  static const int iters = env_int(variant == 0 ? "SPE_EIG_ITERS" : "SPE_JACOBI_SWEEPS", variant == 1 ? 5 : 6);
  static const int t1_warps = env_int("SPE_T1_WARPS", 0), tail_warps = env_int("SPE_TAIL_WARPS", 0);
  a.kernel_variant = variant;
  a.eig_iters = iters > 0 && iters <= 30 ? iters : 6;
  a.t1_warps = t1_warps;
  a.tail_warps = tail_warps;
}
#endif

static int fill_ransac_args(const spe_model_t* model, int B, int hypotheses, void* workspace, size_t workspace_bytes, int flags, spe::RansacArgs& a,
                            spe::RansacWorkspace& ws) {
  if (model == nullptr || B < 0 || hypotheses < 0 || hypotheses > model->m.max_hyp) return SPE_ERR_INVALID_ARGUMENT;
  if (hypotheses == 0 && !(flags & SPE_FLAG_EXACT)) return SPE_ERR_INVALID_ARGUMENT;  // nothing to select from
  if ((flags & SPE_FLAG_JACOBI_SVD)) {
#ifndef SPE_DEV
    return SPE_ERR_UNSUPPORTED;  // development variant, not in the shipped library
#endif
  }
  if (B > 0) {  // the model's tables live on the device it was created on
    int dev = -1;
    const cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    if (dev != model->m.device) return SPE_ERR_INVALID_ARGUMENT;
  }
  const size_t need = spe::ransac_workspace_bytes(model->m.J, B, hypotheses);
  if (B > 0 && (workspace == nullptr || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15u))) return SPE_ERR_WORKSPACE;
  a = spe::RansacArgs{};
  a.B = B;
  a.H = hypotheses;
  a.iterations = model->m.max_hyp;  // cv2's iterationsCount
  a.eig_iters = 6;
  a.exact = (flags & SPE_FLAG_EXACT) ? 1 : 0;
#ifdef SPE_DEV
  dev_knobs(a);
  if ((flags & SPE_FLAG_JACOBI_SVD) && a.kernel_variant == 0) {
    a.kernel_variant = 2;
    a.eig_iters = 6;
  }
#endif
  ws = spe::carve_workspace(workspace, model->m.J, B, hypotheses);
  return SPE_OK;
}

size_t spe_ransac_workspace_bytes(const spe_model_t* model, int B, int hypotheses) {
  if (model == nullptr || B < 0 || hypotheses < 0) return 0;
  return spe::ransac_workspace_bytes(model->m.J, B, hypotheses);
}

int spe_ransac_score_f32(const spe_model_t* model, const float* kpts, int B, int hypotheses, float reproj_err, double confidence,
                         float conf_floor, void* workspace, size_t workspace_bytes, int flags, void* stream) {
  spe::RansacArgs a;
  spe::RansacWorkspace ws;
  const int rc = fill_ransac_args(model, B, hypotheses, workspace, workspace_bytes, flags, a, ws);
  if (rc != SPE_OK || B == 0) return rc;
  if (kpts == nullptr || !(reproj_err > 0.f)) return SPE_ERR_INVALID_ARGUMENT;
  a.kpts = kpts;
  a.reproj_err = reproj_err;
  a.confidence = confidence;
  a.conf_floor = conf_floor;
  a.adaptive = (flags & SPE_FLAG_ADAPTIVE) ? 1 : 0;
  const cudaError_t e = spe::launch_ransac_score(model->m, a, ws, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_ransac_replay_f64(const spe_model_t* model, int B, int hypotheses, float reproj_err, double confidence, void* workspace,
                          size_t workspace_bytes, void* stream) {
  spe::RansacArgs a;
  spe::RansacWorkspace ws;
  const int rc = fill_ransac_args(model, B, hypotheses, workspace, workspace_bytes, SPE_FLAG_EXACT, a, ws);
  if (rc != SPE_OK || B == 0) return rc;
  if (!(reproj_err > 0.f)) return SPE_ERR_INVALID_ARGUMENT;
  a.reproj_err = reproj_err;
  a.confidence = confidence;
  const cudaError_t e = spe::launch_ransac_replay(model->m, a, ws, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_ransac_select_refit_f32(const spe_model_t* model, int B, int hypotheses, double confidence, float* pose7, uint32_t* inlier_mask,
                                int32_t* status, int32_t* winner_hyp, double* rt, void* workspace, size_t workspace_bytes, int flags,
                                void* stream) {
  spe::RansacArgs a;
  spe::RansacWorkspace ws;
  const int rc = fill_ransac_args(model, B, hypotheses, workspace, workspace_bytes, flags, a, ws);
  if (rc != SPE_OK || B == 0) return rc;
  if (pose7 == nullptr || inlier_mask == nullptr || status == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  a.confidence = confidence;
  a.pose7 = pose7;
  a.inlier_mask = inlier_mask;
  a.status = status;
  a.winner = winner_hyp;
  a.rt = rt;
  a.budget = ws.x_visited;  // kept in the workspace for spe_ransac_read_budget
  a.refine_lm = (flags & SPE_FLAG_REFINE_LM) ? 1 : 0;
  a.refit_background = (flags & SPE_FLAG_BACKGROUND_TAIL) ? 1 : 0;
  const cudaError_t e = spe::launch_ransac_select_refit(model->m, a, ws, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_ransac_epnp_f32(const spe_model_t* model, const float* kpts, int B, int hypotheses, float reproj_err, double confidence,
                        float conf_floor, float* pose7, uint32_t* inlier_mask, int32_t* status, int32_t* winner_hyp, double* rt,
                        void* workspace, size_t workspace_bytes, int flags, void* stream) {
  if (B > 0 && (pose7 == nullptr || inlier_mask == nullptr || status == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  int rc = spe_ransac_score_f32(model, kpts, B, hypotheses, reproj_err, confidence, conf_floor, workspace, workspace_bytes, flags, stream);
  if (rc != SPE_OK) return rc;
  if (flags & SPE_FLAG_EXACT) {
    rc = spe_ransac_replay_f64(model, B, hypotheses, reproj_err, confidence, workspace, workspace_bytes, stream);
    if (rc != SPE_OK) return rc;
  }
  return spe_ransac_select_refit_f32(model, B, hypotheses, confidence, pose7, inlier_mask, status, winner_hyp, rt, workspace, workspace_bytes,
                                     flags, stream);
}

int spe_ransac_read_budget(const spe_model_t* model, const void* workspace, int B, int hypotheses, int32_t* budget, void* stream) {
  if (model == nullptr || workspace == nullptr || budget == nullptr || B < 0 || hypotheses < 0) return SPE_ERR_INVALID_ARGUMENT;
  if (B == 0) return SPE_OK;
  const spe::RansacWorkspace ws = spe::carve_workspace(const_cast<void*>(workspace), model->m.J, B, hypotheses);
  const cudaError_t e = cudaMemcpyAsync(budget, ws.x_visited, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_ransac_debug_scores(const spe_model_t* model, const void* workspace, int B, int hypotheses, int32_t* counts, uint32_t* masks,
                            void* stream) {
  if (model == nullptr || workspace == nullptr || B < 0 || hypotheses < 1) return SPE_ERR_INVALID_ARGUMENT;
  const spe::RansacWorkspace ws = spe::carve_workspace(const_cast<void*>(workspace), model->m.J, B, hypotheses);
  const cudaError_t e = spe::launch_debug_scores(model->m, ws, B, hypotheses, counts, masks, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

size_t spe_pipeline_workspace_bytes(const spe_model_t* model, int B, int J, int hypotheses) {
  if (model == nullptr || B < 0 || hypotheses < 0 || J != model->m.J) return 0;
  const size_t kp = ((size_t)B * J * 3 * sizeof(float) + 15) & ~(size_t)15;
  return kp + spe::ransac_workspace_bytes(J, B, hypotheses);
}

int spe_heatmap_to_pose_f32(const spe_model_t* model, const float* hm, int B, int J, int H, int W, const float* center,
                            const float* scale, int post_process, int hypotheses, float reproj_err, double confidence,
                            float conf_floor, float* pose7, uint32_t* inlier_mask, int32_t* status, float* kpts_out, void* workspace,
                            size_t workspace_bytes, int flags, void* stream) {
  if (model == nullptr || J != model->m.J) return SPE_ERR_INVALID_ARGUMENT;
  if (B == 0) return SPE_OK;
  const size_t need = spe_pipeline_workspace_bytes(model, B, J, hypotheses);
  if (workspace == nullptr || need == 0 || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15u)) return SPE_ERR_WORKSPACE;
  const size_t kp = ((size_t)B * J * 3 * sizeof(float) + 15) & ~(size_t)15;
  float* kpts = kpts_out ? kpts_out : static_cast<float*>(workspace);
  int rc = spe_decode_kpts_f32(hm, B, J, H, W, center, scale, post_process, kpts, nullptr, stream);
  if (rc != SPE_OK) return rc;
  return spe_ransac_epnp_f32(model, kpts, B, hypotheses, reproj_err, confidence, conf_floor, pose7, inlier_mask, status, nullptr, nullptr,
                             static_cast<unsigned char*>(workspace) + kp, workspace_bytes - kp, flags, stream);
}

}  // extern "C"
