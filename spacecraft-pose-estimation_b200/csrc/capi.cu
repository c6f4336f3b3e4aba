// C ABI of the heatmap -> pose stage (include/spe_b200.h).  Argument checking and launches only.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"
#include "ransac.cuh"

#include <new>

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

// ---- pose ----------------------------------------------------------------------------------
struct spe_model {
  spe::Model m;
};

int spe_pnp_model_create(const double* landmarks, int J, const double* K, const double* dist, int max_hypotheses, spe_model_t** out) {
  if (out == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (landmarks == nullptr || K == nullptr || J < 4 || J > SPE_MAX_LANDMARKS) return SPE_ERR_INVALID_ARGUMENT;
  if (max_hypotheses < 1 || max_hypotheses > SPE_MAX_HYPOTHESES) return SPE_ERR_INVALID_ARGUMENT;
  spe_model* h = new (std::nothrow) spe_model();
  if (h == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  spe::Model& m = h->m;
  m.J = J;
  m.max_hyp = max_hypotheses;
  for (int i = 0; i < 3 * J; ++i) m.landmarks_f32[i] = (float)landmarks[i];
  m.cam.fx = K[0], m.cam.cx = K[2], m.cam.fy = K[4], m.cam.cy = K[5];
  m.cam.k1 = dist ? dist[0] : 0.0, m.cam.k2 = dist ? dist[1] : 0.0, m.cam.p1 = dist ? dist[2] : 0.0;
  m.cam.p2 = dist ? dist[3] : 0.0, m.cam.k3 = dist ? dist[4] : 0.0;
  const cudaError_t e = spe::model_upload(m);
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

size_t spe_ransac_workspace_bytes(const spe_model_t* model, int B, int hypotheses) {
  if (model == nullptr || B < 0 || hypotheses < 1) return 0;
  return spe::ransac_workspace_bytes(model->m.J, B, hypotheses);
}

static int hyp_kernel_setting() {
  static int variant = [] {
    // dev knob for A/B runs: "g4" = 4 lanes per hypothesis, "jacobi" = thread per hypothesis with the full
    // one-sided Jacobi SVD of M^T; default = thread per hypothesis, Householder QR + inverse iteration
    const char* v = getenv("SPE_HYP_KERNEL");
    return (v && v[0] == 'g') ? 1 : (v && v[0] == 'j') ? 2 : 0;
  }();
  return variant;
}

static int jacobi_sweeps_setting() {
  static int sweeps = [] {
    // dev knob: Jacobi sweeps (variants 1, 2) or inverse-iteration steps (variant 0)
    const char* v = getenv(hyp_kernel_setting() == 0 ? "SPE_EIG_ITERS" : "SPE_JACOBI_SWEEPS");
    const int dflt = hyp_kernel_setting() == 1 ? 5 : 6;
    const int s = v ? atoi(v) : dflt;
    return s > 0 && s <= 30 ? s : dflt;
  }();
  return sweeps;
}

static int t1_warps_setting() {
  static int warps = [] {
    const char* v = getenv("SPE_T1_WARPS");  // dev knob: warps per CTA of the hypothesis kernel
    const int w = v ? atoi(v) : 12;  // 12 warps x 168 registers = one CTA per SM: 0.673 vs 0.681 ms per pipelined step (4 warps)
    return w >= 1 && w <= 12 ? w : 12;
  }();
  return warps;
}

static int tail_warps_setting() {
  static int warps = [] {
    const char* v = getenv("SPE_TAIL_WARPS");  // dev knob: warps per CTA of the background select/refit kernel
    const int w = v ? atoi(v) : 0;  // 0: as many as fit one SM
    return w >= 1 && w <= 32 ? w : 0;
  }();
  return warps;
}

static int fill_ransac_args(const spe_model_t* model, int B, int hypotheses, void* workspace, size_t workspace_bytes, spe::RansacArgs& a,
                            spe::RansacWorkspace& ws) {
  if (model == nullptr || B < 0 || hypotheses < 1 || hypotheses > model->m.max_hyp) return SPE_ERR_INVALID_ARGUMENT;
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
  a.jacobi_sweeps = jacobi_sweeps_setting();
  a.kernel_variant = hyp_kernel_setting();
  a.t1_warps = t1_warps_setting();
  a.tail_warps = tail_warps_setting();
  ws = spe::carve_workspace(workspace, model->m.J, B, hypotheses);
  return SPE_OK;
}

int spe_ransac_score_f32(const spe_model_t* model, const float* kpts, int B, int hypotheses, float reproj_err, double confidence,
                         float conf_floor, void* workspace, size_t workspace_bytes, int flags, void* stream) {
  spe::RansacArgs a;
  spe::RansacWorkspace ws;
  const int rc = fill_ransac_args(model, B, hypotheses, workspace, workspace_bytes, a, ws);
  if (rc != SPE_OK || B == 0) return rc;
  if (kpts == nullptr || !(reproj_err > 0.f)) return SPE_ERR_INVALID_ARGUMENT;
  a.kpts = kpts;
  a.reproj_err = reproj_err;
  a.confidence = confidence;
  a.conf_floor = conf_floor;
  a.adaptive = (flags & SPE_FLAG_ADAPTIVE) ? 1 : 0;
  if ((flags & SPE_FLAG_JACOBI_SVD) && a.kernel_variant == 0) {
    a.kernel_variant = 2;
    a.jacobi_sweeps = 6;
  }
  const cudaError_t e = spe::launch_ransac_score(model->m, a, ws, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_ransac_select_refit_f32(const spe_model_t* model, int B, int hypotheses, double confidence, float* pose7, uint32_t* inlier_mask,
                                int32_t* status, int32_t* winner_hyp, double* rt, void* workspace, size_t workspace_bytes, int flags,
                                void* stream) {
  spe::RansacArgs a;
  spe::RansacWorkspace ws;
  const int rc = fill_ransac_args(model, B, hypotheses, workspace, workspace_bytes, a, ws);
  if (rc != SPE_OK || B == 0) return rc;
  if (pose7 == nullptr || inlier_mask == nullptr || status == nullptr) return SPE_ERR_INVALID_ARGUMENT;
  a.confidence = confidence;
  a.pose7 = pose7;
  a.inlier_mask = inlier_mask;
  a.status = status;
  a.winner = winner_hyp;
  a.rt = rt;
  a.refine_lm = (flags & SPE_FLAG_REFINE_LM) ? 1 : 0;
  a.refit_background = (flags & SPE_FLAG_BACKGROUND_TAIL) ? 1 : 0;
  const cudaError_t e = spe::launch_ransac_select_refit(model->m, a, ws, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

int spe_ransac_epnp_f32(const spe_model_t* model, const float* kpts, int B, int hypotheses, float reproj_err, double confidence,
                        float conf_floor, float* pose7, uint32_t* inlier_mask, int32_t* status, int32_t* winner_hyp, double* rt,
                        void* workspace, size_t workspace_bytes, int flags, void* stream) {
  if (B > 0 && (pose7 == nullptr || inlier_mask == nullptr || status == nullptr)) return SPE_ERR_INVALID_ARGUMENT;
  const int rc = spe_ransac_score_f32(model, kpts, B, hypotheses, reproj_err, confidence, conf_floor, workspace, workspace_bytes, flags, stream);
  if (rc != SPE_OK) return rc;
  return spe_ransac_select_refit_f32(model, B, hypotheses, confidence, pose7, inlier_mask, status, winner_hyp, rt, workspace, workspace_bytes,
                                     flags, stream);
}

int spe_ransac_debug_scores(const spe_model_t* model, const void* workspace, int B, int hypotheses, int32_t* counts, uint32_t* masks,
                            void* stream) {
  if (model == nullptr || workspace == nullptr || B < 0 || hypotheses < 1) return SPE_ERR_INVALID_ARGUMENT;
  const spe::RansacWorkspace ws = spe::carve_workspace(const_cast<void*>(workspace), model->m.J, B, hypotheses);
  const cudaError_t e = spe::launch_debug_scores(ws, B, hypotheses, counts, masks, static_cast<cudaStream_t>(stream));
  return e == cudaSuccess ? SPE_OK : cuda_fail(e);
}

size_t spe_pipeline_workspace_bytes(const spe_model_t* model, int B, int J, int hypotheses) {
  if (model == nullptr || B < 0 || hypotheses < 1 || J != model->m.J) return 0;
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
