// Batched RANSAC-EPnP for sm_100a, part 1: frame preparation and the FP32 hypothesis kernel.
//
// Replaces the front of the per-frame loop of pose_estimation/export_predicted_poses_real.py:177-204 (confidence
// filter :186-197, the hypothesis generation inside cv2.solvePnPRansac(flags=SOLVEPNP_EPNP) :199-201).  OpenCV calib3d
// is an un-vendored dependency of the reference; its algorithm is restated in SURVEY.md App. B and
// oracle/{ocv_rng,epnp_ref,pnp_ref}.py.
//
//   frame_prep_kernel     warp per frame: confidence filter -> visible set, 5-iteration undistortion of every landmark
//                         (float64, cv2.undistortPoints)
//   hypothesis_kernel_t1  FP32, one THREAD per (frame, distinct minimal set).  Every distinct 5-subset among the first H
//                         draws of OpenCV's fixed-seed RNG is scored once (Model::d_uniq / d_slot); there is no early
//                         exit on the GPU.  EPnP's four vectors come from a Householder QR of M^T held in 120 registers
//                         + block inverse iteration, then three beta initialisations, Gauss-Newton, Procrustes, and the
//                         reprojection test of all n visible points -> 32-bit inlier mask + count.
//   budget_kernel         optional (SPE_FLAG_ADAPTIVE): replays cv2's acceptance loop over the first 32 hypotheses
//
// Compute-bound on FP32 CUDA cores (no tensor cores: nothing here is a dense contraction); HBM traffic is ~100 B per
// hypothesis.  Selection and the float64 refit are in ransac_refit.cu, the float64 replay in ransac_exact.cu.
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdint.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"
#include "device_util.cuh"
#include "epnp_math.cuh"
#include "ransac.cuh"
#include "ransac_common.cuh"

namespace spe {

namespace {

constexpr unsigned kFull = kFullMask;

// ------------------------------------------------------------------------------------------
// 1. per-frame preparation: warp per frame, lane j = landmark j
__global__ void __launch_bounds__(128) frame_prep_kernel(DevModel m, const float* __restrict__ kpts, int B, float conf_floor,
                                                          RansacWorkspace ws) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const bool has = lane < m.J;
  float u = 0.f, v = 0.f, conf = -1.f;
  if (has) {
    const float* k = kpts + ((size_t)b * m.J + lane) * 3;
    u = k[0], v = k[1], conf = k[2];
  }
  unsigned vis;
  if (conf_floor >= 0.f) {
    vis = __ballot_sync(kFull, has && conf > conf_floor);
  } else {
    // export_predicted_poses_real.py:186-197: thr = 0.95, *= 0.8 while fewer than 15 pass (<= 100 times).
    // thr lives in float64 (a Python float); the comparison happens in float32 (NumPy array dtype).
    double thr = 0.95;
    vis = __ballot_sync(kFull, has && conf > (float)thr);
    for (int it = 0; it < 100 && __popc(vis) < 15; ++it) {
      thr *= 0.8;
      vis = __ballot_sync(kFull, has && conf > (float)thr);
    }
  }
  if (has) {
    // cv2.undistortPoints: exactly 5 fixed-point iterations in float64 (App. B.3a)
    const Camera& c = m.cam;
    const double x0 = ((double)u - c.cx) / c.fx, y0 = ((double)v - c.cy) / c.fy;
    double x = x0, y = y0;
#pragma unroll 1
    for (int it = 0; it < 5; ++it) {
      const double r2 = x * x + y * y;
      const double icd = 1.0 / (1.0 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
      if (icd < 0) {  // cv2.undistortPoints gives up on a point whose radial factor turns negative
        x = x0, y = y0;
        break;
      }
      const double dx = 2.0 * c.p1 * x * y + c.p2 * (r2 + 2.0 * x * x);
      const double dy = c.p1 * (r2 + 2.0 * y * y) + 2.0 * c.p2 * x * y;
      x = (x0 - dx) * icd;
      y = (y0 - dy) * icd;
    }
    ws.und[(size_t)b * m.J + lane] = make_double2(x, y);
    ws.img[(size_t)b * m.J + lane] = make_float2(u, v);
    // hypotheses see the float32-rounded normalised point (cv2 keeps the input dtype), mapped to
    // ideal pixels in float64 by EPnP's init_points, then held in float32 by this implementation
    const double xf = (double)(float)x, yf = (double)(float)y;
    ws.us_hyp[(size_t)b * m.J + lane] = make_float2((float)(xf * c.fx + c.cx), (float)(yf * c.fy + c.cy));
  }
  if (lane == 0) {
    ws.vis[b] = vis;
    ws.n[b] = __popc(vis);
  }
}

#ifdef SPE_DEV
#include "dev_variants.cuh"
#endif


// ------------------------------------------------------------------------------------------
// 2b. hypothesis kernel, one THREAD per (frame, hypothesis)
//
// Same mathematics, different side of the SVD: instead of rotating the 12 columns of M (and
// accumulating V), the 10 columns of M^T (12 x 10) are orthogonalised.  After convergence the
// normalised columns ARE the right singular vectors of M with non-zero singular value (no
// accumulator needed: 120 registers hold the whole problem), the two smallest give EPnP's v2, v3,
// and the 2-D null space (v0, v1) is the orthogonal complement, built from two columns of the
// projector I - sum v_i v_i^T.  45 pairs x 12 rows per sweep instead of 66 pairs x 22 rows, no
// shuffles, no work replicated across lanes; the three beta variants run one after the other.
#ifndef SPE_T1_REGS
#define SPE_T1_REGS 168
#endif
__device__ __forceinline__ int ws_frames(const RansacWorkspace& ws) { return ws.frames; }
constexpr int kT1Stride = 4 * 12 + 20 + 1;  // per-thread scratch: v[4][12], alphas[5][4] (+1: odd stride, conflict-free)
// shared memory of one warp: landmarks [32][3], ideal + raw pixels [32] float2 each, per-thread scratch
constexpr int kT1WarpBytes = (int)(sizeof(float) * 3 * kMaxLandmarks + 2 * sizeof(float2) * kMaxLandmarks + sizeof(float) * 32 * kT1Stride +
                                   kMaxLandmarks /* landmark id of every compacted point */);
constexpr int kT1MaxWarps = 65536 / (32 * ((SPE_T1_REGS + 7) / 8 * 8));  // one CTA = one SM's register file: 12 warps x 168 registers
static_assert(kT1WarpBytes % 16 == 0, "per-warp shared-memory slice must stay 16-byte aligned");

// U(n, h_begin) and U(n, h_end) for every point count (the distinct-set slots a launch covers), from the host tables
struct UniqueRange {
  uint16_t begin[kMaxLandmarks + 1], end[kMaxLandmarks + 1];
};

template <int kEig>  // 0: Householder QR + inverse iteration (default), 1: one-sided Jacobi SVD of M^T
__global__ void __maxnreg__(SPE_T1_REGS)
hypothesis_kernel_t1(DevModel m, const float* __restrict__ kpts, int H, int h_begin, int h_end, int hblocks, UniqueRange ur, const int32_t* __restrict__ need,
                     float thr2, int sweeps, RansacWorkspace ws) {
  // One warp = one work item: 32 consecutive DISTINCT minimal sets [u_begin + 32*hb, +32) of frame b, in the order of
  // their first draw (Model::d_uniq).  Warps of a CTA are independent (different frames in general), each with its own
  // slice of (dynamic) shared memory, so the CTA size is a launch-time choice (launch_ransac_score).
  extern __shared__ __align__(16) unsigned char t1_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kWarps = blockDim.x >> 5;
  unsigned char* const wbase = t1_smem + (size_t)warp * kT1WarpBytes;
  float(*s_pw)[3] = reinterpret_cast<float(*)[3]>(wbase);
  float2* s_us = reinterpret_cast<float2*>(wbase + sizeof(float) * 3 * kMaxLandmarks);
  float2* s_img = s_us + kMaxLandmarks;
  float* work = reinterpret_cast<float*>(s_img + kMaxLandmarks) + lane * kT1Stride;
  uint8_t* s_id = reinterpret_cast<uint8_t*>(reinterpret_cast<float*>(s_img + kMaxLandmarks) + 32 * kT1Stride);
  const long long item = (long long)blockIdx.x * kWarps + warp;
  const int b = (int)(item / hblocks), hb = (int)(item - (long long)b * hblocks);
  if (b >= ws_frames(ws)) return;
  const int n = ws.n[b];
  if (n <= kModelPoints) return;
  // hypotheses [h_begin, limit) of this frame are wanted (adaptive second pass: limit = the frame's remaining budget);
  // the distinct sets first drawn in that range are the slots [U(n, h_begin), U(n, limit))
  const int limit = need ? min(need[b], h_end) : h_end;
  if (limit <= h_begin) return;
  const int u_begin = ur.begin[n];  // = U(n, h_begin), computed on the host
  const int u = u_begin + hb * 32 + lane;
  if (u - lane >= limit) return;  // slot u draws at index >= u: nothing of this block lies below the limit
  const int u_end = need ? unique_below(m, n, limit) : (int)ur.end[n];  // a per-frame limit needs the search, a common one does not
  if (u - lane >= u_end) return;
  const unsigned vis = ws.vis[b];
  if (lane < n) {
    const int j = __fns(vis, 0, lane + 1);
    s_pw[lane][0] = m.landmarks[3 * j], s_pw[lane][1] = m.landmarks[3 * j + 1], s_pw[lane][2] = m.landmarks[3 * j + 2];
    s_us[lane] = ws.us_hyp[(size_t)b * m.J + j];
    const float* k = kpts + ((size_t)b * m.J + j) * 3;
    s_img[lane] = make_float2(k[0], k[1]);
    s_id[lane] = (uint8_t)j;
  }
  __syncwarp();
  if (u >= u_end) return;
  const int h = m.uniq[(size_t)(n - 6) * m.max_hyp + u];  // the draw that introduces this set
  const uint8_t* sub = m.subsets + ((size_t)(n - 6) * m.max_hyp + h) * kModelPoints;
  // The five points are handled in ascending landmark order (EPnP does not depend on the order of
  // its points beyond rounding): that is the order of the control-point table's alphas.  Key =
  // landmark id * 32 + compacted index, sorted with a 9-exchange network.
  int si[5];
  unsigned ctrl_rank;
  {
    int key[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int idx = sub[k];
      key[k] = ((int)s_id[idx] << 5) | idx;
    }
    auto cx = [&](int a, int b) {
      const int lo = min(key[a], key[b]), hi = max(key[a], key[b]);
      key[a] = lo, key[b] = hi;
    };
    cx(0, 1), cx(3, 4), cx(2, 4), cx(2, 3), cx(0, 3), cx(0, 2), cx(1, 4), cx(1, 3), cx(1, 2);
    unsigned j[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) si[k] = key[k] & 31, j[k] = (unsigned)key[k] >> 5;
    // C(j0,1) + C(j1,2) + C(j2,3) + C(j3,4) + C(j4,5); every product is exactly divisible
    ctrl_rank = j[0] + j[1] * (j[1] - 1) / 2 + j[2] * (j[2] - 1) * (j[2] - 2) / 6 + j[3] * (j[3] - 1) * (j[3] - 2) * (j[3] - 3) / 24 +
                j[4] * (j[4] - 1) * (j[4] - 2) * (j[4] - 3) / 24 * (j[4] - 4) / 5;
  }
  const float fu = (float)m.cam.fx, fv = (float)m.cam.fy, uc = (float)m.cam.cx, vc = (float)m.cam.cy;

  // ---- control points, alphas, M^T ---------------------------------------------------------
  float rho[6];
  float A[12][10];
#ifdef SPE_DEV
  float d[10];
#endif
  {
    float al[5][4];
    {
      float e[kCtrlEntryFloats];
      const float4* src = m.ctrl + (size_t)ctrl_rank * (kCtrlEntryFloats / 4);
#pragma unroll
      for (int q = 0; q < kCtrlEntryFloats / 4; ++q) {
        const float4 v = __ldg(src + q);
        e[4 * q] = v.x, e[4 * q + 1] = v.y, e[4 * q + 2] = v.z, e[4 * q + 3] = v.w;
      }
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        al[k][1] = e[3 * k], al[k][2] = e[3 * k + 1], al[k][3] = e[3 * k + 2];
        al[k][0] = 1.0f - al[k][1] - al[k][2] - al[k][3];
      }
      // rho over the control-point pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3): the axes are orthogonal
      rho[0] = e[15], rho[1] = e[16], rho[2] = e[17];
      rho[3] = e[15] + e[16], rho[4] = e[15] + e[17], rho[5] = e[16] + e[17];
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float du = uc - s_us[si[k]].x, dv = vc - s_us[si[k]].y;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = al[k][j];
        work[48 + 4 * k + j] = a;
        if constexpr (kEig == 0) {  // rows (x, z) of control point j, then the y rows; x-equations first
          A[2 * j][k] = a * fu, A[2 * j + 1][k] = a * du, A[8 + j][k] = 0.f;
          A[2 * j][5 + k] = 0.f, A[2 * j + 1][5 + k] = a * dv, A[8 + j][5 + k] = a * fv;
        } else {
          A[3 * j][2 * k] = a * fu, A[3 * j + 1][2 * k] = 0.f, A[3 * j + 2][2 * k] = a * du;
          A[3 * j][2 * k + 1] = 0.f, A[3 * j + 1][2 * k + 1] = a * fv, A[3 * j + 2][2 * k + 1] = a * dv;
        }
      }
    }
  }
#ifdef SPE_DEV
  if constexpr (kEig == 0) {
    eig_qr_inverse_iteration(A, work, sweeps);
  } else {
  jacobi_mt(A, d, sweeps);

  // ---- v2, v3 = the two smallest singular directions; v0, v1 = null space (complement) ---------
  {
    int j2 = 0, j3 = 0;
    float b2 = INFINITY, b3 = INFINITY;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const bool lt2 = d[j] < b2, lt3 = d[j] < b3;
      j3 = lt2 ? j2 : (lt3 ? j : j3);
      b3 = lt2 ? b2 : (lt3 ? d[j] : b3);
      j2 = lt2 ? j : j2;
      b2 = lt2 ? d[j] : b2;
    }
    float diag[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) diag[r] = 1.0f;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      const float inv = rsqrt_approx(fmaxf(d[j], 1e-30f));
      float v2r, v3r;
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        A[r][j] *= inv;
        diag[r] = fmaf(-A[r][j], A[r][j], diag[r]);
        v2r = A[r][j];
        if (j == j2) work[24 + r] = v2r;
        if (j == j3) work[36 + r] = v2r;
      }
      (void)v3r;
    }
    // first null vector: the projector column with the largest norm
    float n0[12], n1[12];
    {
      int a = 0;
      float best = diag[0];
#pragma unroll
      for (int r = 1; r < 12; ++r) {
        const bool gt = diag[r] > best;
        best = gt ? diag[r] : best;
        a = gt ? r : a;
      }
#pragma unroll
      for (int r = 0; r < 12; ++r) n0[r] = r == a ? 1.0f : 0.0f;
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        float coef = 0.f;
#pragma unroll
        for (int r = 0; r < 12; ++r) coef = r == a ? A[r][j] : coef;
#pragma unroll
        for (int r = 0; r < 12; ++r) n0[r] = fmaf(-coef, A[r][j], n0[r]);
      }
      float nn = 0.f;
#pragma unroll
      for (int r = 0; r < 12; ++r) nn = fmaf(n0[r], n0[r], nn);
      const float inv = rsqrt_approx(fmaxf(nn, 1e-30f));
#pragma unroll
      for (int r = 0; r < 12; ++r) n0[r] *= inv;
    }
    {
      int bsel = 0;
      float best = -1.f;
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        const float res = diag[r] - n0[r] * n0[r];
        const bool gt = res > best;
        best = gt ? res : best;
        bsel = gt ? r : bsel;
      }
#pragma unroll
      for (int r = 0; r < 12; ++r) n1[r] = r == bsel ? 1.0f : 0.0f;
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        float coef = 0.f;
#pragma unroll
        for (int r = 0; r < 12; ++r) coef = r == bsel ? A[r][j] : coef;
#pragma unroll
        for (int r = 0; r < 12; ++r) n1[r] = fmaf(-coef, A[r][j], n1[r]);
      }
      float dot = 0.f;
#pragma unroll
      for (int r = 0; r < 12; ++r) dot = fmaf(n1[r], n0[r], dot);
      float nn = 0.f;
#pragma unroll
      for (int r = 0; r < 12; ++r) {
        n1[r] = fmaf(-dot, n0[r], n1[r]);
        nn = fmaf(n1[r], n1[r], nn);
      }
      const float inv = rsqrt_approx(fmaxf(nn, 1e-30f));
#pragma unroll
      for (int r = 0; r < 12; ++r) n1[r] *= inv;
    }
#pragma unroll
    for (int r = 0; r < 12; ++r) work[r] = n0[r], work[12 + r] = n1[r];
  }
  }  // kEig
#else
  static_assert(kEig == 0, "the Jacobi SVD eigen stage is a development variant (-DSPE_DEV)");
  eig_qr_inverse_iteration(A, work, sweeps);
#endif

  // ---- the three beta variants, one after the other; keep the best by OpenCV's rule -----------
  float Rb[3][3], tb[3], eb = 0.f;
  {
    float L[6][10];
    {
      float v[4][12];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 12; ++j) v[i][j] = work[12 * i + j];
      build_L<float, true>(v, L);
      // computed ONCE: without the barrier the compiler sinks the 60 entries into the variant loop
      // below and recomputes them three times (ncu source page, profiles/solver_r1.md)
#pragma unroll
      for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int i = 0; i < 10; ++i) asm volatile("" : "+f"(L[k][i]));
    }
    float pw[5][3], pw0[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        pw[k][c] = s_pw[si[k]][c];
        pw0[c] += pw[k][c];
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) pw0[c] *= 0.2f;
    // The three beta variants side by side: initialisation, five Gauss-Newton steps (interleaved), then the
    // camera-frame control points and the 3x3 cross-covariance of each, for the interleaved Procrustes.
    float betas[3][4];
#pragma unroll
    for (int v = 0; v < 3; ++v) approx_betas<float, true, true>(L, rho, v + 1, betas[v]);
    gauss_newton_doubled_batch<3>(L, rho, betas);
    float abt_all[3][3][3], pc0_all[3][3];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      float ccs[4][3];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          ccs[j][c] = betas[v][0] * work[3 * j + c] + betas[v][1] * work[12 + 3 * j + c] + betas[v][2] * work[24 + 3 * j + c] +
                      betas[v][3] * work[36 + 3 * j + c];
      float pcs[5][3];
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          pcs[k][c] = work[48 + 4 * k] * ccs[0][c] + work[48 + 4 * k + 1] * ccs[1][c] + work[48 + 4 * k + 2] * ccs[2][c] +
                      work[48 + 4 * k + 3] * ccs[3][c];
      const float sgn = pcs[0][2] < 0.f ? -1.f : 1.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) pc0_all[v][c] = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          pcs[k][c] *= sgn;
          pc0_all[v][c] += pcs[k][c];
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) pc0_all[v][c] *= 0.2f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) abt_all[v][r][c] = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int c = 0; c < 3; ++c) abt_all[v][r][c] = fmaf(pcs[k][r] - pc0_all[v][r], pw[k][c] - pw0[c], abt_all[v][r][c]);
    }
    float R_all[3][3][3];
    procrustes_uvt_batch<3>(abt_all, R_all);
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      float t[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) t[r] = pc0_all[v][r] - (R_all[v][r][0] * pw0[0] + R_all[v][r][1] * pw0[1] + R_all[v][r][2] * pw0[2]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float Xc = R_all[v][0][0] * pw[k][0] + R_all[v][0][1] * pw[k][1] + R_all[v][0][2] * pw[k][2] + t[0];
        const float Yc = R_all[v][1][0] * pw[k][0] + R_all[v][1][1] * pw[k][1] + R_all[v][1][2] * pw[k][2] + t[1];
        const float iz = rcp_approx(R_all[v][2][0] * pw[k][0] + R_all[v][2][1] * pw[k][1] + R_all[v][2][2] * pw[k][2] + t[2]);
        const float du = s_us[si[k]].x - (uc + fu * Xc * iz), dv = s_us[si[k]].y - (vc + fv * Yc * iz);
        sum += sqrt_approx(du * du + dv * dv);
      }
      const float err = sum * 0.2f;
      if (v == 0 || err < eb) {  // N = 1; if (e2 < e1) N = 2; if (e3 < e[N]) N = 3
        eb = err;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) Rb[r][c] = R_all[v][r][c];
          tb[r] = t[r];
        }
      }
    }
  }

  // ---- score all n points ------------------------------------------------------------------
  unsigned bits = 0;
  {
    const float k1 = (float)m.cam.k1, k2 = (float)m.cam.k2, p1 = (float)m.cam.p1, p2 = (float)m.cam.p2, k3 = (float)m.cam.k3;
    unsigned rest = vis;
    for (int k = 0; k < n; ++k) {
      const int j = __ffs(rest) - 1;  // landmark number of the k-th visible point
      rest &= rest - 1;
      const float X = s_pw[k][0], Y = s_pw[k][1], Z = s_pw[k][2];
      const float xc = Rb[0][0] * X + Rb[0][1] * Y + Rb[0][2] * Z + tb[0];
      const float yc = Rb[1][0] * X + Rb[1][1] * Y + Rb[1][2] * Z + tb[1];
      const float zc = Rb[2][0] * X + Rb[2][1] * Y + Rb[2][2] * Z + tb[2];
      const float iz = rcp_approx(zc);
      const float x = xc * iz, y = yc * iz;
      const float r2 = x * x + y * y;
      const float cd = 1.0f + ((k3 * r2 + k2) * r2 + k1) * r2;
      const float xd = x * cd + 2.0f * p1 * x * y + p2 * (r2 + 2.0f * x * x);
      const float yd = y * cd + p1 * (r2 + 2.0f * y * y) + 2.0f * p2 * x * y;
      const float du = s_img[k].x - (fu * xd + uc), dv = s_img[k].y - (fv * yd + vc);
      const float e = du * du + dv * dv;
      if (e <= thr2) bits |= 1u << j;
    }
  }
  ws.masks[(size_t)b * H + u] = bits;
  ws.counts[(size_t)b * H + u] = (uint8_t)__popc(bits);
}

// ------------------------------------------------------------------------------------------
// 2c. adaptive budget (optional): after the first kFirstPass hypotheses of every frame have been
// scored, replay cv2's acceptance loop over them.  need[b] = how many hypotheses cv2 could still
// look at (its shrinking iteration budget, capped at H), or 0 when its loop has already ended.
// The second pass scores only those; select_refit then reads exactly the entries cv2 would read,
// so the result is identical to scoring all H.
constexpr int kFirstPass = 32;

__global__ void __launch_bounds__(128) budget_kernel(DevModel m, RansacWorkspace ws, int B, int H, double confidence, int32_t* need) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = ws.n[b];
  int out = 0;
  if (n > kModelPoints) {
    const uint8_t* counts = ws.counts + (size_t)b * H;
    const uint16_t* slot = m.slot + (size_t)(n - 6) * m.max_hyp;
    int niters = H, max_good = 0;
    const int first = min(kFirstPass, H);
    for (int h = 0; h < min(niters, first); ++h) {
      const int g = counts[slot[h]];
      if (g > max(max_good, kModelPoints - 1)) {
        max_good = g;
        niters = update_num_iters(confidence, (double)(n - g) / n, niters);
      }
    }
    out = niters > first ? min(niters, H) : 0;
  }
  need[b] = out;
}

__global__ void debug_scores_kernel(DevModel m, RansacWorkspace ws, int B, int H, int32_t* counts, uint32_t* masks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * H) return;
  const int b = (int)(i / H), h = (int)(i - (long long)b * H);
  const int n = ws.n[b];
  if (n <= kModelPoints) {  // no RANSAC for this frame: nothing was scored
    if (counts) counts[i] = 0;
    if (masks) masks[i] = 0;
    return;
  }
  const size_t at = (size_t)b * H + m.slot[(size_t)(n - 6) * m.max_hyp + h];
  if (counts) counts[i] = ws.counts[at];
  if (masks) masks[i] = ws.masks[at];
}

}  // namespace

cudaError_t launch_frame_prep(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0) return cudaSuccess;
  const int wpb = 4;
  frame_prep_kernel<<<(a.B + wpb - 1) / wpb, wpb * 32, 0, stream>>>(dev_model(m), a.kpts, a.B, a.conf_floor, ws);
  return cudaGetLastError();
}

cudaError_t launch_ransac_score(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0) return cudaSuccess;
  const DevModel dm = dev_model(m);
  cudaError_t e = launch_frame_prep(m, a, ws, stream);
  if (e != cudaSuccess) return e;
  if (m.J <= kModelPoints || a.H < 1) return cudaSuccess;
  const float thr2 = a.reproj_err * a.reproj_err;
#ifdef SPE_DEV
  if (a.kernel_variant == 1) {  // 4 lanes per hypothesis: no duplicate elimination, so it fills the slots by replaying every draw
    const int hblocks = (a.H + kHypPerCta - 1) / kHypPerCta;
    const long long ctas = (long long)a.B * hblocks;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    hypothesis_kernel<<<(unsigned)ctas, kHypPerCta * kGroup, 0, stream>>>(dm, a.kpts, a.H, hblocks, thr2, a.eig_iters, ws);
    return cudaGetLastError();
  }
#endif
  static PerDeviceOnce once;  // same shared-memory/L1 split as the decode kernel (decode.cuh)
  e = once.run(m.device, [] {
    cudaError_t r = cudaFuncSetAttribute(hypothesis_kernel_t1<0>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(hypothesis_kernel_t1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT1MaxWarps * kT1WarpBytes);
#ifdef SPE_DEV
    if (r == cudaSuccess) r = cudaFuncSetAttribute(hypothesis_kernel_t1<1>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(hypothesis_kernel_t1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT1MaxWarps * kT1WarpBytes);
#endif
    return r;
  });
  if (e != cudaSuccess) return e;
  const int max_warps = a.t1_warps >= 1 && a.t1_warps <= kT1MaxWarps ? a.t1_warps : kT1MaxWarps;
  int dev = 0, num_sms = 0;
  e = current_device(dev, num_sms);
  if (e != cudaSuccess) return e;
  // Hypotheses [h_begin, h_end) of every frame.  The grid covers the largest number of distinct sets any point count
  // has in that range; warps whose frame has fewer (or a smaller adaptive budget) leave at once.
  auto launch = [&](int h_begin, int h_end, const int32_t* need) -> cudaError_t {
    int most = 0;
    UniqueRange ur{};
    for (int n = kModelPoints + 1; n <= m.J; ++n) {
      ur.begin[n] = (uint16_t)unique_sets(m, n, h_begin);
      ur.end[n] = (uint16_t)unique_sets(m, n, h_end);
      most = std::max(most, (int)ur.end[n] - (int)ur.begin[n]);
    }
    const int hblocks = (most + 31) / 32;
    const long long items = (long long)a.B * hblocks;
    // small batches: spread the warps over the SMs instead of packing 12 of them into one CTA
    const int kWarps = (int)std::min<long long>(max_warps, std::max<long long>(1, (items + num_sms - 1) / num_sms));
    const size_t smem = (size_t)kWarps * kT1WarpBytes;
    const long long ctas = (items + kWarps - 1) / kWarps;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (ctas == 0) return cudaSuccess;
#ifdef SPE_DEV
    if (a.kernel_variant == 2) {  // full SVD of M^T
      hypothesis_kernel_t1<1><<<(unsigned)ctas, kWarps * 32, smem, stream>>>(dm, a.kpts, a.H, h_begin, h_end, hblocks, ur, need, thr2, a.eig_iters, ws);
      return cudaGetLastError();
    }
#endif
    hypothesis_kernel_t1<0><<<(unsigned)ctas, kWarps * 32, smem, stream>>>(dm, a.kpts, a.H, h_begin, h_end, hblocks, ur, need, thr2, a.eig_iters, ws);
    return cudaGetLastError();
  };
  if (a.adaptive && a.H > kFirstPass) {
    e = launch(0, kFirstPass, nullptr);
    if (e != cudaSuccess) return e;
    budget_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(dm, ws, a.B, a.H, a.confidence, ws.need);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = launch(kFirstPass, a.H, ws.need);
  } else {
    e = launch(0, a.H, nullptr);
  }
  return e;
}

cudaError_t launch_debug_scores(const Model& m, const RansacWorkspace& ws, int B, int H, int32_t* counts, uint32_t* masks, cudaStream_t stream) {
  const long long total = (long long)B * H;
  if (total == 0) return cudaSuccess;
  debug_scores_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(dev_model(m), ws, B, H, counts, masks);
  return cudaGetLastError();
}

}  // namespace spe
