// C ABI of the heatmap -> pose stage (include/spe_b200.h).  Argument checking and launches only.
#include <cuda_runtime.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"

namespace {

thread_local cudaError_t t_last_cuda = cudaSuccess;

int cuda_fail(cudaError_t e) {
  t_last_cuda = e;
  return SPE_ERR_CUDA;
}

int run_decode(const float* hm, int B, int J, int H, int W, const float* center, const float* scale, int post_process,
               float* preds, float* maxvals, float* kpts, int32_t* argmax, void* stream) {
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
  if (center == nullptr || scale == nullptr || kpts == nullptr) return B == 0 ? SPE_OK : SPE_ERR_INVALID_ARGUMENT;
  return run_decode(hm, B, J, H, W, center, scale, post_process, nullptr, nullptr, kpts, argmax, stream);
}

}  // extern "C"
