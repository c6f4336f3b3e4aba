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

namespace spe {

namespace {

constexpr int kLanesPerFrame = 8;  // hypotheses evaluated per round of a frame
constexpr int kGroupsPerWarp = 32 / kLanesPerFrame;
constexpr int kReplayWarps = 4;  // 128 threads x 255 registers: two CTAs per SM
constexpr int kExactEigIters = 12;  // block inverse-iteration steps of the float64 eigen stage

struct FramePoints {  // one frame's visible landmarks, compacted (shared memory, one per group)
  double pw[kMaxLandmarks][3];   // object points (the float32-rounded landmarks)
  double us[kMaxLandmarks][2];   // ideal pixel coordinates of the float32-rounded undistorted points (hypothesis input)
  float img[kMaxLandmarks][2];   // raw pixel coordinates (scoring)
  uint8_t id[kMaxLandmarks];     // landmark number of every compacted point
};

// EPnP on the five points `sub` (indices into the compacted frame, draw order) in float64, then the inlier mask of the
// resulting pose over the frame's n points.  One thread, everything in registers / local memory.
__device__ __noinline__ unsigned hypothesis_f64(const Camera& cam, const FramePoints& f, int n, const uint8_t* __restrict__ sub, float thr2) {
  const double fu = cam.fx, fv = cam.fy, uc = cam.cx, vc = cam.cy;
  double pw[5][3], us[5][2];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int i = sub[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) pw[k][c] = f.pw[i][c];
    us[k][0] = f.us[i][0], us[k][1] = f.us[i][1];
  }
  // control points: centroid + PCA axes from OpenCV's Jacobi (the signs place the control points), App. B.3c
  double cws[4][3], pw0[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) pw0[c] += pw[k][c];
#pragma unroll
  for (int c = 0; c < 3; ++c) cws[0][c] = pw0[c] = pw0[c] / 5.0;
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, dc[3];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const double q[3] = {pw[k][0] - cws[0][0], pw[k][1] - cws[0][1], pw[k][2] - cws[0][2]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) cov[3 * r + c] += q[r] * q[c];
  }
  cv_jacobi_rows(cov, dc, 3);  // cov now holds the axes as rows
  double inv_k[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double k = sqrt(dc[i] / 5.0);
    inv_k[i] = k > 0 ? 1.0 / k : 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) cws[i + 1][c] = cws[0][c] + k * cov[3 * i + c];
  }
  double rho[6];
  build_rho<double>(cws, rho);
  // barycentric coordinates (CC = [k_i u_i] has orthogonal columns: CC^-1 = diag(1/k) U^T), then A = M^T in the row /
  // column order of eig_qr_inverse_iteration
  double al[5][4];
  double v[4][12];
  {
    double A[12][10];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double q[3] = {pw[k][0] - cws[0][0], pw[k][1] - cws[0][1], pw[k][2] - cws[0][2]};
#pragma unroll
      for (int j = 0; j < 3; ++j) al[k][1 + j] = (cov[3 * j] * q[0] + cov[3 * j + 1] * q[1] + cov[3 * j + 2] * q[2]) * inv_k[j];
      al[k][0] = 1.0 - al[k][1] - al[k][2] - al[k][3];
      const double du = uc - us[k][0], dv = vc - us[k][1];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double a = al[k][j];
        A[2 * j][k] = a * fu, A[2 * j + 1][k] = a * du, A[8 + j][k] = 0.0;
        A[2 * j][5 + k] = 0.0, A[2 * j + 1][5 + k] = a * dv, A[8 + j][5 + k] = a * fv;
      }
    }
    eig_qr_inverse_iteration<double>(A, &v[0][0], kExactEigIters);
  }
  double L[6][10];
  build_L<double>(v, L);
  double Rb[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, tb[3] = {0, 0, 0}, eb = 0.0;
#pragma unroll 1
  for (int variant = 1; variant <= 3; ++variant) {
    double be[4];
    approx_betas<double>(L, rho, variant, be);
    gauss_newton<double>(L, rho, be);
    double ccs[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) ccs[j][c] = be[0] * v[0][3 * j + c] + be[1] * v[1][3 * j + c] + be[2] * v[2][3 * j + c] + be[3] * v[3][3 * j + c];
    double pcs[5][3], pc0[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) pcs[k][c] = al[k][0] * ccs[0][c] + al[k][1] * ccs[1][c] + al[k][2] * ccs[2][c] + al[k][3] * ccs[3][c];
    const double sgn = pcs[0][2] < 0 ? -1.0 : 1.0;  // solve_for_sign: the first point lies in front of the camera
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        pcs[k][c] *= sgn;
        pc0[c] += pcs[k][c];
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) pc0[c] /= 5.0;
    double abt[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) abt[r][c] += (pcs[k][r] - pc0[r]) * (pw[k][c] - pw0[c]);
    double R[3][3], t[3];
    procrustes_uvt<double>(abt, R);
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = pc0[r] - (R[r][0] * pw0[0] + R[r][1] * pw0[1] + R[r][2] * pw0[2]);
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double Xc = R[0][0] * pw[k][0] + R[0][1] * pw[k][1] + R[0][2] * pw[k][2] + t[0];
      const double Yc = R[1][0] * pw[k][0] + R[1][1] * pw[k][1] + R[1][2] * pw[k][2] + t[1];
      const double iz = 1.0 / (R[2][0] * pw[k][0] + R[2][1] * pw[k][1] + R[2][2] * pw[k][2] + t[2]);
      const double du = us[k][0] - (uc + fu * Xc * iz), dv = us[k][1] - (vc + fv * Yc * iz);
      sum += sqrt(du * du + dv * dv);
    }
    const double err = sum / 5.0;
    if (variant == 1 || err < eb) {  // N = 1; if (e2 < e1) N = 2; if (e3 < e[N]) N = 3
      eb = err;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rb[r][c] = R[r][c];
        tb[r] = t[r];
      }
    }
  }
  // cv2.projectPoints in float64 -> float32 image points; squared error in float32 (App. B.5)
  unsigned bits = 0;
  for (int k = 0; k < n; ++k) {
    const double X = f.pw[k][0], Y = f.pw[k][1], Z = f.pw[k][2];
    const double xc = Rb[0][0] * X + Rb[0][1] * Y + Rb[0][2] * Z + tb[0];
    const double yc = Rb[1][0] * X + Rb[1][1] * Y + Rb[1][2] * Z + tb[1];
    const double zc = Rb[2][0] * X + Rb[2][1] * Y + Rb[2][2] * Z + tb[2];
    const double iz = zc != 0.0 ? 1.0 / zc : 1.0;
    const double x = xc * iz, y = yc * iz;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2.0 * x * y, a2 = r2 + 2.0 * x * x, a3 = r2 + 2.0 * y * y;
    const double cd = 1.0 + cam.k1 * r2 + cam.k2 * r4 + cam.k3 * r6;
    const double xd = x * cd + cam.p1 * a1 + cam.p2 * a2;
    const double yd = y * cd + cam.p1 * a3 + cam.p2 * a1;
    const float pu = (float)(xd * fu + uc), pv = (float)(yd * fv + vc);
    const float du = f.img[k][0] - pu, dv = f.img[k][1] - pv;
    const float e = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
    if (e <= thr2) bits |= 1u << f.id[k];
  }
  return bits;
}

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
