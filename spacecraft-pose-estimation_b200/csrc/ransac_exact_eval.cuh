// The arithmetic of the float64 replay (ransac_exact.cu): one 5-point EPnP hypothesis in float64 + the reprojection test
// of all visible points, as a __host__ __device__ function so that tests/host/exact_eval_host.cu can run the very same
// code on the CPU against cv2 (tests/test_exact_eval_host.py, no GPU needed).  Not a CPU path of the product: the
// library only ever calls it from replay_kernel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "epnp_f64.cuh"
#include "epnp_math.cuh"
#include "ransac.cuh"

#ifdef __CUDA_ARCH__
#define SPE_NOINLINE __noinline__
#else
#define SPE_NOINLINE
#endif

// The small least-squares solves of the hypothesis evaluation (beta initialisations, Gauss-Newton steps) go through the
// normal equations + Cholesky: in float64 the squared conditioning is harmless (agreement with cv2 identical to the
// Householder version on every dataset of tests/test_exact_eval_host.py) and it costs ~40 % less.  The final refit
// (ransac_refit.cu) keeps Householder QR.
#ifndef SPE_EXACT_NORMAL_EQ
#define SPE_EXACT_NORMAL_EQ true
#endif

namespace spe {

constexpr int kExactEigIters = 10;  // inverse-iteration steps of the float64 eigen stage (block of four vectors)

struct FramePoints {  // one frame's visible landmarks, compacted (workspace, written once per call by replay_plan_kernel)
  double pw[kMaxLandmarks][3];   // object points (the float32-rounded landmarks)
  double us[kMaxLandmarks][2];   // ideal pixel coordinates of the float32-rounded undistorted points (hypothesis input)
  float img[kMaxLandmarks][2];   // raw pixel coordinates (scoring)
  uint8_t id[kMaxLandmarks];     // landmark number of every compacted point
};
static_assert(sizeof(FramePoints) == kReplayFrameBytes, "FramePoints is carved as kReplayFrameBytes per frame");

// EPnP on the five points `sub` (indices into the compacted frame, draw order) in float64, then the inlier mask of the
// resulting pose over the frame's n points.  One thread, everything in registers / local memory.
// dbg (host harness only, may be null): v[48], then per variant (err, betas[4]) x 3, then R[9], t[3]
SPE_HD SPE_NOINLINE unsigned hypothesis_f64(const Camera& cam, const FramePoints& f, int n, const uint8_t* __restrict__ sub, float thr2,
                                            double* dbg = nullptr) {
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
    eig_qr_subspace4<double>(A, &v[0][0], kExactEigIters);
  }
#ifndef __CUDA_ARCH__
  if (dbg)
    for (int i = 0; i < 48; ++i) dbg[i] = (&v[0][0])[i];
#endif
  double L[6][10];
#ifndef SPE_NO_NULL_BASIS_ROTATION
  // Canonical basis of the 2-D null space.  Any orthonormal (v0, v1) is a legal outcome of OpenCV's SVD (the two
  // singular values are exactly zero: its own basis is decided by rounding noise), but EPnP's initialisations are not
  // invariant under the choice: they read beta0^2 and beta1^2 off the linearised solution b = (b00, b01, b11) and drop
  // the cross term.  The solution itself IS covariant — as a symmetric 2 x 2 form B it just rotates with the basis — so
  // the basis is turned to B's eigenvectors, where the dropped cross term is zero and the read-off is exact for the
  // rank-1 part.  A fixed QR basis is systematically bad for some landmark subsets (4 nearly coplanar points + 1: the
  // initialisation lands in the wrong basin on every frame), which a noise-chosen basis is not.
  {
    double A3[6][3], r6[6], b3[3];
    {
      constexpr int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
      for (int k = 0; k < 6; ++k) {  // the three columns of L that involve v0 and v1 only
        double d0[3], d1[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) d0[c] = v[0][3 * pa[k] + c] - v[0][3 * pb[k] + c], d1[c] = v[1][3 * pa[k] + c] - v[1][3 * pb[k] + c];
        A3[k][0] = d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2];
        A3[k][1] = 2.0 * (d0[0] * d1[0] + d0[1] * d1[1] + d0[2] * d1[2]);
        A3[k][2] = d1[0] * d1[0] + d1[1] * d1[1] + d1[2] * d1[2];
        r6[k] = rho[k];
      }
    }
    lsq_householder<double, 6, 3>(A3, r6, b3);
    // B = [[b0, b1/2], [b1/2, b2]]; Jacobi angle that diagonalises it, larger |eigenvalue| first
    const double off = 0.5 * b3[1], h = b3[2] - b3[0];
    double c = 1.0, sn = 0.0;
    if (off != 0.0) {
      const double q = sqrt(h * h + 4.0 * off * off);
      const double t = 2.0 * off / (h + copysign(q, h != 0.0 ? h : 1.0));
      c = 1.0 / sqrt(1.0 + t * t);
      sn = c * t;
    }
    // rotated basis: v0' = c v0 - s v1, v1' = s v0 + c v1 diagonalises B; eigenvalues
    const double l0 = b3[0] - (sn / c) * off, l1 = b3[2] + (sn / c) * off;
    const bool swap = fabs(l1) > fabs(l0);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const double a = c * v[0][i] - sn * v[1][i], b = sn * v[0][i] + c * v[1][i];
      v[0][i] = swap ? b : a;
      v[1][i] = swap ? -a : b;
    }
  }
#endif
  build_L<double>(v, L);
  double Rb[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, tb[3] = {0, 0, 0}, eb = 0.0;
#pragma unroll 1
  for (int variant = 1; variant <= 3; ++variant) {
    double be[4];
    approx_betas<double, SPE_EXACT_NORMAL_EQ>(L, rho, variant, be);
    gauss_newton<double, SPE_EXACT_NORMAL_EQ>(L, rho, be);
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
#ifndef __CUDA_ARCH__
    if (dbg) {
      dbg[48 + 5 * (variant - 1)] = err;
      for (int i = 0; i < 4; ++i) dbg[49 + 5 * (variant - 1) + i] = be[i];
    }
#endif
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
#ifndef __CUDA_ARCH__
  if (dbg) {
    for (int i = 0; i < 9; ++i) dbg[63 + i] = Rb[i / 3][i % 3];
    for (int i = 0; i < 3; ++i) dbg[72 + i] = tb[i];
  }
#endif
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
#ifdef __CUDA_ARCH__
    const float e = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));  // no FMA contraction: cv2 rounds both squares
#else
    const float e = du * du + dv * dv;  // host harness: compiled with -ffp-contract=off
#endif
    if (e <= thr2) bits |= 1u << f.id[k];
  }
  return bits;
}

}  // namespace spe
