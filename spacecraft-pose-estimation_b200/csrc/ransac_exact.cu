// Batched RANSAC-EPnP for sm_100a, part 2: float64 replay of cv2's own RANSAC loop (SPE_FLAG_EXACT).
//
// cv2.solvePnPRansac (pose_estimation/export_predicted_poses_real.py:199-201) is sequential and adaptive: hypothesis h
// is looked at only while h < niters, and every accepted model (first strictly better inlier count) shrinks niters
// (SURVEY App. B.6).  On the benchmark data it looks at 6.5 hypotheses per frame on average (90 % of the frames: <= 12)
// and decides on float64 arithmetic.  This kernel does exactly that, frame by frame, in float64:
//
//   * a group of 8 lanes owns one frame at a time; the four groups of a warp draw their frames from a counter in
//     the caller's workspace (zeroed by the launcher on the call's stream), so a frame that needs many rounds does not
//     hold the other three groups' work back;
//   * a round = the next 8 hypotheses of the frame, one per lane: thread-serial EPnP on the five points of the
//     minimal set in float64 (control points from OpenCV's sign-defining Jacobi, M^T -> Householder QR + block inverse
//     iteration for the four vectors, three beta initialisations, five Gauss-Newton steps each with Householder QR,
//     Procrustes, best of three by mean reprojection error — OpenCV's sequence, oracle/epnp_ref.py), then
//     cv2.projectPoints of all n visible points in float64, rounded to float32, squared error in float32 <= thr^2;
//   * the 8 results are walked in order with cv2's acceptance rule and RANSACUpdateNumIters; the frame is finished when
//     the next hypothesis index reaches niters (or `iterations`, cv2's iterationsCount: 10000 in the reference).
//
// The FP32 scores of ransac_score.cu are not read here: which hypotheses decide is cv2's rule, and those are
// re-evaluated in float64 whatever FP32 said about them.
// Why a replay and not per-hypothesis bit-parity: a 5-point EPnP has an exact 2-D null space, so ~8 % of cv2's own
// hypothesis results are decided by rounding noise (measured: a float64 NumPy restatement reproduces cv2's
// per-hypothesis inlier count on 92 % of minimal sets whether it uses a port of OpenCV's Jacobi SVD or LAPACK's
// eigensolver).  What is reproducible is the outcome of the loop — the winner's inlier set — and that needs float64 on
// the hypotheses cv2 looks at: 99.0 % of close-range frames against 77.6 % with FP32 scores (profiles/parity_r2.md).
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdint.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"
#include "device_util.cuh"
#include "epnp_f64.cuh"
#include "epnp_math.cuh"
#include "ransac.cuh"
#include "ransac_common.cuh"
#include "ransac_exact_eval.cuh"

namespace spe {

namespace {

constexpr int kLanesPerFrame = 8;  // hypotheses evaluated per round of a frame
constexpr int kGroupsPerWarp = 32 / kLanesPerFrame;
constexpr int kReplayWarps = 4;  // 128 threads x 255 registers: two CTAs per SM

__global__ void __launch_bounds__(kReplayWarps * 32) replay_kernel(DevModel m, RansacArgs a, RansacWorkspace ws) {
  __shared__ FramePoints s_frames[kReplayWarps * kGroupsPerWarp];
  const int lane = threadIdx.x & 31, sub = lane & (kLanesPerFrame - 1), gbase = lane & ~(kLanesPerFrame - 1);
  const unsigned gmask = ((1u << kLanesPerFrame) - 1u) << gbase;
  FramePoints& f = s_frames[threadIdx.x / kLanesPerFrame];
  const float thr2 = a.reproj_err * a.reproj_err;
  // group state, replicated on the group's lanes
  int b = -1, n = 0, base = 0, niters = 0, max_good = 0, winner = -1;
  unsigned best_mask = 0;
  bool exhausted = false;  // no frames left for this group
  while (true) {
    if (b < 0 && !exhausted) {
      // claim the next frame that needs RANSAC (n >= 6); the others are settled by select_refit_kernel
      for (;;) {
        int nb = 0;
        if (sub == 0) nb = (int)atomicAdd(ws.claim, 1u);
        nb = __shfl_sync(gmask, nb, gbase);
        if (nb >= a.B) {
          exhausted = true;
          break;
        }
        n = ws.n[nb];
        if (n > kModelPoints) {
          b = nb;
          break;
        }
        if (sub == 0) ws.x_winner[nb] = -1, ws.x_mask[nb] = 0, ws.x_visited[nb] = 0;
      }
      if (b >= 0) {
        const unsigned vis = ws.vis[b];
        __syncwarp(gmask);  // the previous frame's readers are done with f
        for (int k = sub; k < n; k += kLanesPerFrame) {
          const int j = __fns(vis, 0, k + 1);
#pragma unroll
          for (int c = 0; c < 3; ++c) f.pw[k][c] = (double)m.landmarks[3 * j + c];
          const double2 q = ws.und[(size_t)b * m.J + j];
          // hypotheses see the float32-rounded undistorted point (cv2 keeps the input dtype), App. B.3a
          f.us[k][0] = (double)(float)q.x * m.cam.fx + m.cam.cx;
          f.us[k][1] = (double)(float)q.y * m.cam.fy + m.cam.cy;
          const float2 px = ws.img[(size_t)b * m.J + j];
          f.img[k][0] = px.x, f.img[k][1] = px.y;
          f.id[k] = (uint8_t)j;
        }
        __syncwarp(gmask);
        base = 0, niters = a.iterations, max_good = 0, winner = -1, best_mask = 0;
      }
    }
    if (__all_sync(kFullMask, b < 0)) break;  // every group of the warp has run out of frames
    const int h = base + sub;
    unsigned bits = 0;
    if (b >= 0 && h < niters) {
      const uint8_t* subset = m.subsets + ((size_t)(n - 6) * m.max_hyp + h) * kModelPoints;
      bits = hypothesis_f64(m.cam, f, n, subset, thr2);
    }
    if (b >= 0) {
      // cv2's acceptance loop over this round, in order (every lane of the group runs it on the same values)
#pragma unroll 1
      for (int k = 0; k < kLanesPerFrame; ++k) {
        const unsigned mk = __shfl_sync(gmask, bits, gbase + k);
        const int g = __popc(mk);
        if (base + k < niters && g > max(max_good, kModelPoints - 1)) {
          winner = base + k;
          max_good = g;
          best_mask = mk;
          niters = update_num_iters(a.confidence, (double)(n - g) / n, niters);
        }
      }
      base += kLanesPerFrame;
      if (base >= niters) {
        if (sub == 0) {
          ws.x_winner[b] = winner;
          ws.x_mask[b] = best_mask;
          ws.x_visited[b] = niters;
        }
        b = -1;
      }
    }
  }
}

}  // namespace

cudaError_t launch_ransac_replay(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0 || m.J <= kModelPoints) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(ws.claim, 0, sizeof(uint32_t) * 4, stream);
  if (e != cudaSuccess) return e;
  static PerDeviceOnce once;
  e = once.run(m.device, [] { return cudaFuncSetAttribute(replay_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct); });
  if (e != cudaSuccess) return e;
  int dev = 0, num_sms = 0;
  e = current_device(dev, num_sms);
  if (e != cudaSuccess) return e;
  // persistent groups: enough CTAs to fill the GPU (2 per SM at 255 registers), never more groups than frames
  const int groups_per_cta = kReplayWarps * kGroupsPerWarp;
  const int ctas = std::max(1, std::min(2 * num_sms, (a.B + groups_per_cta - 1) / groups_per_cta));
  replay_kernel<<<ctas, kReplayWarps * 32, 0, stream>>>(dev_model(m), a, ws);
  return cudaGetLastError();
}

}  // namespace spe
