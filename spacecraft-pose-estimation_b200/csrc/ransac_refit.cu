// Batched RANSAC-EPnP for sm_100a, part 3: selection and the float64 refit.
//
// Replaces the back of cv2.solvePnPRansac as the reference calls it (pose_estimation/export_predicted_poses_real.py:199-201)
// plus cv2.Rodrigues (:203) and cv_rotation_matrix_to_quat (:22-57):
//   select_refit_kernel  fast mode: replays cv2's sequential acceptance rule (first strictly better count, adaptively
//                        shrinking iteration budget) over the FP32 inlier counts; exact mode: takes the
//                        winner of the float64 replay (ransac_exact.cu).  Then the final EPnP on the winner's inliers in
//                        float64, following OpenCV operation by operation where signs depend on it (PCA axes), and the
//                        optional Levenberg-Marquardt step (SPE_FLAG_REFINE_LM).
#include <cuda_runtime.h>

#include <math.h>
#include <stdint.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"
#include "device_util.cuh"
#include "epnp_f64.cuh"
#include "epnp_math.cuh"
#include "p3p_f64.cuh"
#include "ransac.cuh"
#include "ransac_common.cuh"

#ifndef SPE_REFIT_REGS
#define SPE_REFIT_REGS 255
#endif

namespace spe {

namespace {

// ONE copy per frame in shared memory, used by all four lanes of the frame's group.  After the
// accumulation every lane would hold the same matrix and the tridiagonalisation / QL steps are the same
// instructions on the same data in lockstep, so the lanes read the same address (a broadcast) and
// write the same value.  9.3 KB per warp instead of 36.9 KB (per-thread copies in shared memory) or
// 36.9 KB of local memory streaming through L1: eight such warps fit one SM.  The frame stride of
// 145 doubles keeps the eight frames of a warp on different banks.
constexpr int kFrameMatStride = 145;
struct FrameMat12 {
  static constexpr bool kSharedPerFrame = true;
  double* base;  // &storage[frame slot * kFrameMatStride]
  __device__ __forceinline__ double& operator()(int r, int c) const { return base[r * 12 + c]; }
};
template <typename Mat>
struct MatTraits {
  static constexpr bool kSharedPerFrame = false;
};
template <>
struct MatTraits<FrameMat12> {
  static constexpr bool kSharedPerFrame = true;
};


// Eigenvectors of the four SMALLEST eigenvalues of a symmetric N x N matrix (float64), for MtM in
// the final refit: Householder tridiagonalisation (the classic tred2 reduction, reflectors kept,
// Q never formed), eigenvalues by implicit QL without vectors (tql1), the four wanted
// eigenvectors of the tridiagonal matrix by inverse iteration (pivoted tridiagonal LU, three
// iterations, orthogonalised against the vectors already found) and back-transformation through
// the stored reflectors.  ~5x fewer instructions than accumulating all 12 eigenvectors through QL,
// ~20x fewer than the one-sided Jacobi OpenCV runs here; residuals |A v - w v| / |A| ~ 4e-16 on EPnP
// matrices (tools prototype).  Eigenvector SIGNS are irrelevant for MtM (they matter for the 3x3
// PCA, which keeps cv_jacobi_rows).  V is destroyed.  out[k] <-> k-th smallest eigenvalue.
// Called by the 4 lanes of a frame's group with identical inputs (sub = lane within the group,
// gmask = the group's lanes): the reduction and the eigenvalues are computed redundantly, then lane
// k runs the inverse iteration and the back-transformation of eigenvector k; the iterates are
// exchanged with shuffles after every iteration and orthonormalised in order by every lane.
template <int N, typename Mat>
__device__ void sym_eig_smallest4(Mat V, double (&out)[4][N], int sub, unsigned gmask) {
  const int gbase = (threadIdx.x & 31) & ~3;
  double d[N], e[N], hs[N], diag[N];
  // A matrix shared by the four lanes of the group (FrameMat12): the lanes hold d, e replicated in registers and
  // read the matrix; lane 0 writes it, except in the rank-2 update where the lanes split the rows of a column.
  // Every write is separated from the other lanes' reads of the same entry by a __syncwarp of the group: once per
  // outer step, and once per column of the rank-2 update (every lane needs V(i-1, j) back, which it computes
  // itself from the value read before the barrier).
  constexpr bool kShared = MatTraits<Mat>::kSharedPerFrame;
  const bool writer = !kShared || sub == 0;
  // --- reduction to tridiagonal form
  for (int j = 0; j < N; ++j) d[j] = V(N - 1, j);
  for (int i = N - 1; i > 0; --i) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; ++k) scale += fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) {
        d[j] = V(i - 1, j);
        if (writer) {
          V(i, j) = 0.0;
          V(j, i) = 0.0;
        }
      }
    } else {
      const double inv_scale = 1.0 / scale;
      for (int k = 0; k < i; ++k) {
        d[k] *= inv_scale;
        h += d[k] * d[k];
      }
      double f = d[i - 1];
      double g = sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.0;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        if (writer) V(j, i) = f;
        g = e[j] + V(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) {
          g += V(k, j) * d[k];
          e[k] += V(k, j) * f;
        }
        e[j] = g;
      }
      f = 0.0;
      const double inv_h = 1.0 / h;
      for (int j = 0; j < i; ++j) {
        e[j] *= inv_h;
        f += e[j] * d[j];
      }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j];
        g = e[j];
        // the one updated entry every lane needs back; read before lane 0 rewrites the column
        const double last = V(i - 1, j) - (f * e[i - 1] + g * d[i - 1]);
        if constexpr (kShared) __syncwarp(gmask);
        // shared matrix: the four lanes split the rows of the column (d, e are replicated in their registers)
        for (int k = j + (kShared ? sub : 0); k <= i - 1; k += (kShared ? 4 : 1)) V(k, j) -= (f * e[k] + g * d[k]);
        if (writer) V(i, j) = 0.0;
        d[j] = last;  // = V(i - 1, j)
      }
    }
    d[i] = h;
    if constexpr (kShared) __syncwarp(gmask);
  }
  // reflector m (m >= 1) acts on coordinates 0..m-1: u = V(0..m-1, m), h = hs[m]; T = tridiag(diag, e[1..])
  double tnorm = 0.0;
  for (int j = 0; j < N; ++j) {
    hs[j] = d[j];
    diag[j] = V(j, j);
    tnorm = fmax(tnorm, fmax(fabs(diag[j]), j > 0 ? fabs(e[j]) : 0.0));
  }
  e[0] = 0.0;
  const double eps = 2.220446049250313e-16;
  // --- eigenvalues: implicit QL on a copy (d2, e2)
  double d2[N], e2[N];
  for (int j = 0; j < N; ++j) d2[j] = diag[j];
  for (int j = 1; j < N; ++j) e2[j - 1] = e[j];
  e2[N - 1] = 0.0;
  {
    double f = 0.0, tst1 = 0.0;
    for (int l = 0; l < N; ++l) {
      tst1 = fmax(tst1, fabs(d2[l]) + fabs(e2[l]));
      int m = l;
      while (m < N) {
        if (fabs(e2[m]) <= eps * tst1) break;
        ++m;
      }
      if (m > l) {
        int iter = 0;
        do {
          ++iter;
          double g = d2[l];
          double p = (d2[l + 1] - g) / (2.0 * e2[l]);
          double r = sqrt(fma(p, p, 1.0));
          if (p < 0) r = -r;
          d2[l] = e2[l] / (p + r);
          d2[l + 1] = e2[l] * (p + r);
          const double dl1 = d2[l + 1];
          double h = g - d2[l];
          for (int i = l + 2; i < N; ++i) d2[i] -= h;
          f += h;
          p = d2[m];
          double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
          const double el1 = e2[l + 1];
          for (int i = m - 1; i >= l; --i) {
            c3 = c2;
            c2 = c;
            s2 = s;
            g = c * e2[i];
            h = c * p;
            const double rr = fma(p, p, e2[i] * e2[i]);
            const double rinv = rr > 0.0 ? rsqrt(rr) : 0.0;
            r = rr * rinv;
            e2[i + 1] = s * r;
            s = e2[i] * rinv;
            c = p * rinv;
            p = c * d2[i] - s * g;
            d2[i + 1] = h + s * (c * g + s * d2[i]);
          }
          p = -s * s2 * c3 * el1 * e2[l] / dl1;
          e2[l] = s * p;
          d2[l] = c * p;
        } while (fabs(e2[l]) > eps * tst1 && iter < 60);
      }
      d2[l] += f;
      e2[l] = 0.0;
    }
  }
  // the four smallest, ascending
  double lam[4];
  for (int k = 0; k < 4; ++k) {
    int jm = 0;
    for (int j = 1; j < N; ++j)
      if (d2[j] < d2[jm]) jm = j;
    lam[k] = d2[jm];
    d2[jm] = 1.7976931348623157e308;
  }
  // --- inverse iteration on the tridiagonal matrix (lane k <-> eigenvector k) + back-transformation
  double l = lam[0];
  {
    double prev = 0.0;
    for (int k = 0; k < 4; ++k) {
      double lk = lam[k];
      if (k > 0 && lk - prev < 10.0 * eps * tnorm) lk = prev + 10.0 * eps * tnorm;  // keep numerically equal eigenvalues apart
      prev = lk;
      if (k == sub) l = lk;
    }
  }
  double dd[N], dl[N], du[N], du2[N];
  bool piv[N];
  for (int j = 0; j < N; ++j) dd[j] = diag[j] - l;
  for (int j = 0; j < N - 1; ++j) dl[j] = du[j] = e[j + 1], du2[j] = 0.0, piv[j] = false;
  for (int i = 0; i < N - 1; ++i) {  // pivoted LU of the shifted tridiagonal matrix
    if (fabs(dd[i]) >= fabs(dl[i])) {
      if (dd[i] == 0.0) dd[i] = eps * tnorm;
      const double fact = dl[i] / dd[i];
      dl[i] = fact;
      dd[i + 1] -= fact * du[i];
    } else {
      const double fact = dd[i] / dl[i];
      dd[i] = dl[i];
      dl[i] = fact;
      const double tmp = du[i];
      du[i] = dd[i + 1];
      dd[i + 1] = tmp - fact * dd[i + 1];
      if (i < N - 2) {
        du2[i] = du[i + 1];
        du[i + 1] = -fact * du[i + 1];
      }
      piv[i] = true;
    }
  }
  if (dd[N - 1] == 0.0) dd[N - 1] = eps * tnorm;
  double x[N];
  {
    double nn = 0.0;
    for (int j = 0; j < N; ++j) {
      x[j] = (double)((j * 7 + sub * 3) % 5 - 2) + 0.37 * (sub + 1);
      nn += x[j] * x[j];
    }
    const double inv = rsqrt(nn);
    for (int j = 0; j < N; ++j) x[j] *= inv;
  }
  for (int it = 0; it < 3; ++it) {
    for (int i = 0; i < N - 1; ++i) {
      if (!piv[i]) {
        x[i + 1] -= dl[i] * x[i];
      } else {
        const double tmp = x[i];
        x[i] = x[i + 1];
        x[i + 1] = tmp - dl[i] * x[i];
      }
    }
    x[N - 1] /= dd[N - 1];
    x[N - 2] = (x[N - 2] - du[N - 2] * x[N - 1]) / dd[N - 2];
    for (int i = N - 3; i >= 0; --i) x[i] = (x[i] - du[i] * x[i + 1] - du2[i] * x[i + 2]) / dd[i];
    {  // scale before the exchange (the solve amplifies by ~1/|lambda - l|)
      double nn = 0.0;
      for (int i = 0; i < N; ++i) nn = fmax(nn, fabs(x[i]));
      const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
      for (int i = 0; i < N; ++i) x[i] *= inv;
    }
    __syncwarp(gmask);
    for (int j = 0; j < 4; ++j)
      for (int i = 0; i < N; ++i) out[j][i] = __shfl_sync(gmask, x[i], gbase + j);
    for (int k = 0; k < 4; ++k) {  // modified Gram-Schmidt in eigenvalue order
      for (int j = 0; j < k; ++j) {
        double dot = 0.0;
        for (int i = 0; i < N; ++i) dot += out[k][i] * out[j][i];
        for (int i = 0; i < N; ++i) out[k][i] -= dot * out[j][i];
      }
      double nn = 0.0;
      for (int i = 0; i < N; ++i) nn += out[k][i] * out[k][i];
      const double inv = nn > 0.0 ? rsqrt(nn) : 0.0;
      for (int i = 0; i < N; ++i) out[k][i] *= inv;
    }
    for (int k = 0; k < 4; ++k)
      if (k == sub)
        for (int i = 0; i < N; ++i) x[i] = out[k][i];
  }
  for (int m = 1; m < N; ++m) {  // v = Q y = H_{N-1} ... H_1 y, own vector only
    if (hs[m] != 0.0) {
      double dot = 0.0;
      for (int i = 0; i < m; ++i) dot += V(i, m) * x[i];
      dot /= hs[m];
      for (int i = 0; i < m; ++i) x[i] -= dot * V(i, m);
    }
  }
  __syncwarp(gmask);
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < N; ++i) out[j][i] = __shfl_sync(gmask, x[i], gbase + j);
}

// EPnP on n points in float64, OpenCV's sequence (epnp::compute_pose).
// Runs on the 4 lanes of a frame's group (identical inputs on every lane): MtM is accumulated over
// the lane's share of the points and summed with shuffles, the eigenvectors are split over the
// lanes, and the three beta variants run one per lane (lane 3 repeats variant 3).
template <typename Mat>
__device__ void epnp_f64(int n, const double (*pw)[3], const double (*und)[2], const Camera& cam, double (&Rbest)[3][3],
                         double (&tbest)[3], int sub, unsigned gmask, Mat mtm) {
  const int gbase = (threadIdx.x & 31) & ~3;
  const double fu = cam.fx, fv = cam.fy, uc = cam.cx, vc = cam.cy;
  double us[kMaxLandmarks][2], al[kMaxLandmarks][4];
  for (int i = 0; i < n; ++i) {
    us[i][0] = und[i][0] * fu + uc;
    us[i][1] = und[i][1] * fv + vc;
  }
  // control points: centroid + PCA axes from OpenCV's Jacobi (signs matter)
  double cws[4][3] = {};
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) cws[0][c] += pw[i][c];
  for (int c = 0; c < 3; ++c) cws[0][c] /= n;
  double cov[9] = {}, dc[3];
  for (int i = 0; i < n; ++i) {
    const double q[3] = {pw[i][0] - cws[0][0], pw[i][1] - cws[0][1], pw[i][2] - cws[0][2]};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) cov[3 * r + c] += q[r] * q[c];
  }
  cv_jacobi_rows(cov, dc, 3);  // cov now holds uct
  double inv_k[3];
  for (int i = 0; i < 3; ++i) {
    const double k = sqrt(dc[i] / n);
    inv_k[i] = k > 0 ? 1.0 / k : 0.0;
    for (int c = 0; c < 3; ++c) cws[i + 1][c] = cws[0][c] + k * cov[3 * i + c];
  }
  // barycentric coordinates: CC = [k_i u_i] has orthogonal columns, so CC^-1 = diag(1/k) U^T
  for (int i = 0; i < n; ++i) {
    const double q[3] = {pw[i][0] - cws[0][0], pw[i][1] - cws[0][1], pw[i][2] - cws[0][2]};
    for (int j = 0; j < 3; ++j) al[i][1 + j] = (cov[3 * j] * q[0] + cov[3 * j + 1] * q[1] + cov[3 * j + 2] * q[2]) * inv_k[j];
    al[i][0] = 1.0 - al[i][1] - al[i][2] - al[i][3];
  }
  static_assert(MatTraits<Mat>::kSharedPerFrame, "the refit keeps one working matrix per frame in shared memory");
  {
    // MtM (12x12), one copy per frame.  Row 2i of M = alpha_i (x) (fu, 0, du_i), row 2i+1 = alpha_i (x) (0, fv, dv_i)
    // with du_i = uc - u_i, dv_i = vc - v_i, so every entry is fu^2, fv^2, fu, fv or 1 times one of the 40 sums
    //   S_m[jr][jc] = sum_i alpha_i[jr] alpha_i[jc] w_m(i),  w = 1, du_i, dv_i, du_i^2 + dv_i^2,  jr <= jc.
    // Lane `sub` computes the sums t = sub, sub + 4, ... and scatters them (with their transposes).
    for (int idx = sub; idx < 144; idx += 4) mtm(idx / 12, idx - 12 * (idx / 12)) = 0.0;  // (x, y) cross entries stay 0
    __syncwarp(gmask);
    for (int t = sub; t < 40; t += 4) {
      const int mi = t / 10, pr = t - 10 * mi;  // pairs (0,0)(0,1)(0,2)(0,3)(1,1)(1,2)(1,3)(2,2)(2,3)(3,3)
      const int jr = pr < 4 ? 0 : (pr < 7 ? 1 : (pr < 9 ? 2 : 3));
      const int jc = pr < 4 ? pr : (pr < 7 ? pr - 3 : (pr < 9 ? pr - 5 : 3));
      double acc = 0.0;
      for (int i = 0; i < n; ++i) {
        const double du = uc - us[i][0], dv = vc - us[i][1];
        const double w = mi == 0 ? 1.0 : (mi == 1 ? du : (mi == 2 ? dv : du * du + dv * dv));
        acc += al[i][jr] * al[i][jc] * w;
      }
      auto set = [&](int r, int c, double v) {
        mtm(r, c) = v;
        mtm(c, r) = v;
      };
      if (mi == 0) {
        set(3 * jr, 3 * jc, fu * fu * acc);
        set(3 * jr + 1, 3 * jc + 1, fv * fv * acc);
      } else if (mi == 1) {
        set(3 * jr, 3 * jc + 2, fu * acc);
        set(3 * jr + 2, 3 * jc, fu * acc);
      } else if (mi == 2) {
        set(3 * jr + 1, 3 * jc + 2, fv * acc);
        set(3 * jr + 2, 3 * jc + 1, fv * acc);
      } else {
        set(3 * jr + 2, 3 * jc + 2, acc);
      }
    }
    __syncwarp(gmask);
  }
  // eigenvectors of MtM for the four smallest eigenvalues: v0 = smallest (OpenCV's ut[11]) ... v3
  double v[4][12];
  sym_eig_smallest4<12, Mat>(mtm, v, sub, gmask);
  double L[6][10], rho[6];
  build_L<double>(v, L);
  build_rho<double>(cws, rho);

  double pw0[3] = {};
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) pw0[c] += pw[i][c];
  for (int c = 0; c < 3; ++c) pw0[c] /= n;

  double R[3][3], t[3], err;
  {
    const int variant = sub < 3 ? sub + 1 : 3;
    double be[4];
    approx_betas<double>(L, rho, variant, be);
    gauss_newton<double>(L, rho, be);
    double ccs[4][3];
    for (int j = 0; j < 4; ++j)
      for (int c = 0; c < 3; ++c) ccs[j][c] = be[0] * v[0][3 * j + c] + be[1] * v[1][3 * j + c] + be[2] * v[2][3 * j + c] + be[3] * v[3][3 * j + c];
    // sign from the first point's depth
    double z0 = 0;
    for (int j = 0; j < 4; ++j) z0 += al[0][j] * ccs[j][2];
    const double sgn = z0 < 0 ? -1.0 : 1.0;
    double pc0[3] = {};
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) pc0[c] += sgn * (al[i][0] * ccs[0][c] + al[i][1] * ccs[1][c] + al[i][2] * ccs[2][c] + al[i][3] * ccs[3][c]);
    for (int c = 0; c < 3; ++c) pc0[c] /= n;
    double abt[3][3] = {};
    for (int i = 0; i < n; ++i) {
      double pc[3];
      for (int c = 0; c < 3; ++c) pc[c] = sgn * (al[i][0] * ccs[0][c] + al[i][1] * ccs[1][c] + al[i][2] * ccs[2][c] + al[i][3] * ccs[3][c]);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) abt[r][c] += (pc[r] - pc0[r]) * (pw[i][c] - pw0[c]);
    }
    procrustes_uvt<double>(abt, R);
    for (int r = 0; r < 3; ++r) t[r] = pc0[r] - (R[r][0] * pw0[0] + R[r][1] * pw0[1] + R[r][2] * pw0[2]);
    double sum = 0;
    for (int i = 0; i < n; ++i) {
      const double Xc = R[0][0] * pw[i][0] + R[0][1] * pw[i][1] + R[0][2] * pw[i][2] + t[0];
      const double Yc = R[1][0] * pw[i][0] + R[1][1] * pw[i][1] + R[1][2] * pw[i][2] + t[1];
      const double iz = 1.0 / (R[2][0] * pw[i][0] + R[2][1] * pw[i][1] + R[2][2] * pw[i][2] + t[2]);
      const double du = us[i][0] - (uc + fu * Xc * iz), dv = us[i][1] - (vc + fv * Yc * iz);
      sum += sqrt(du * du + dv * dv);
    }
    err = sum / n;
  }
  // N = 1; if (err2 < err1) N = 2; if (err3 < err[N]) N = 3 — then everybody takes lane N-1's pose
  __syncwarp(gmask);
  const double e1 = __shfl_sync(gmask, err, gbase), e2 = __shfl_sync(gmask, err, gbase + 1), e3 = __shfl_sync(gmask, err, gbase + 2);
  int N = 0;
  if (e2 < e1) N = 1;
  if (e3 < (N == 1 ? e2 : e1)) N = 2;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Rbest[r][c] = __shfl_sync(gmask, R[r][c], gbase + N);
    tbest[r] = __shfl_sync(gmask, t[r], gbase + N);
  }
}

// Optional reprojection-error refinement of (R, t) over the inliers, the counterpart of
// cv2.solvePnPRefineLM(obj[inl], img[inl], K, dist, rvec, tvec): Levenberg-Marquardt on
// sum |project(X_i) - x_i|^2 with the 5-coefficient distortion model, float64.  OpenCV's solver
// stops within 1e-13 deg of the minimiser (20 iterations, eps = FLT_EPSILON), so a converged
// minimisation matches it; the rotation is updated by left-multiplied so(3) increments.
// NOT part of the reference's call (its SOLVEPNP_EPNP path has no LM step).
__device__ void refine_lm_f64(int n, const double (*pw)[3], const double (*img)[2], const Camera& cam, double (&R)[3][3], double (&t)[3]) {
  auto cost_and_normal = [&](const double (&Rc)[3][3], const double (&tc)[3], double (*JtJ)[6], double* Jtr) {
    double S = 0.0;
    if (JtJ) {
      for (int i = 0; i < 6; ++i) {
        Jtr[i] = 0.0;
        for (int j = 0; j < 6; ++j) JtJ[i][j] = 0.0;
      }
    }
    for (int p = 0; p < n; ++p) {
      const double X = pw[p][0], Y = pw[p][1], Z = pw[p][2];
      const double rx = Rc[0][0] * X + Rc[0][1] * Y + Rc[0][2] * Z, ry = Rc[1][0] * X + Rc[1][1] * Y + Rc[1][2] * Z,
                   rz = Rc[2][0] * X + Rc[2][1] * Y + Rc[2][2] * Z;
      const double xc = rx + tc[0], yc = ry + tc[1], zc = rz + tc[2];
      const double iz = 1.0 / zc, x = xc * iz, y = yc * iz;
      const double r2 = x * x + y * y;
      const double cd = 1.0 + ((cam.k3 * r2 + cam.k2) * r2 + cam.k1) * r2;
      const double xd = x * cd + 2.0 * cam.p1 * x * y + cam.p2 * (r2 + 2.0 * x * x);
      const double yd = y * cd + cam.p1 * (r2 + 2.0 * y * y) + 2.0 * cam.p2 * x * y;
      const double eu = cam.fx * xd + cam.cx - img[p][0], ev = cam.fy * yd + cam.cy - img[p][1];
      S += eu * eu + ev * ev;
      if (JtJ) {
        const double cp = (3.0 * cam.k3 * r2 + 2.0 * cam.k2) * r2 + cam.k1;  // d cd / d r2
        const double dxdx = cd + 2.0 * x * x * cp + 2.0 * cam.p1 * y + 6.0 * cam.p2 * x;
        const double dxdy = 2.0 * x * y * cp + 2.0 * cam.p1 * x + 2.0 * cam.p2 * y;
        const double dydx = dxdy;
        const double dydy = cd + 2.0 * y * y * cp + 6.0 * cam.p1 * y + 2.0 * cam.p2 * x;
        // d(x, y) / d Xc
        const double a00 = iz, a02 = -x * iz, a11 = iz, a12 = -y * iz;
        // d(u, v) / d Xc  (2 x 3)
        const double g[2][3] = {{cam.fx * dxdx * a00, cam.fx * dxdy * a11, cam.fx * (dxdx * a02 + dxdy * a12)},
                                {cam.fy * dydx * a00, cam.fy * dydy * a11, cam.fy * (dydx * a02 + dydy * a12)}};
        // d Xc / d(omega, t): omega rotates R X, i.e. d Xc = omega x (R X) + dt
        double Jr[2][6];
        for (int e = 0; e < 2; ++e) {
          Jr[e][0] = g[e][2] * ry - g[e][1] * rz;
          Jr[e][1] = g[e][0] * rz - g[e][2] * rx;
          Jr[e][2] = g[e][1] * rx - g[e][0] * ry;
          Jr[e][3] = g[e][0], Jr[e][4] = g[e][1], Jr[e][5] = g[e][2];
        }
        for (int i = 0; i < 6; ++i) {
          Jtr[i] += Jr[0][i] * eu + Jr[1][i] * ev;
          for (int j = i; j < 6; ++j) JtJ[i][j] += Jr[0][i] * Jr[0][j] + Jr[1][i] * Jr[1][j];
        }
      }
    }
    return S;
  };
  double JtJ[6][6], Jtr[6];
  double S = cost_and_normal(R, t, JtJ, Jtr);
  double lambda = 1e-3;
  for (int iter = 0; iter < 50; ++iter) {
    // (JtJ + lambda diag) d = -Jtr, Cholesky on the upper triangle
    double U[6][6], d[6];
    bool spd = true;
    for (int i = 0; i < 6 && spd; ++i) {
      double dg = JtJ[i][i] * (1.0 + lambda);
      for (int k = 0; k < i; ++k) dg -= U[k][i] * U[k][i];
      if (!(dg > 0.0)) {
        spd = false;
        break;
      }
      const double inv = 1.0 / sqrt(dg);
      U[i][i] = inv;  // stores 1 / U_ii
      for (int j = i + 1; j < 6; ++j) {
        double acc = JtJ[i][j];
        for (int k = 0; k < i; ++k) acc -= U[k][i] * U[k][j];
        U[i][j] = acc * inv;
      }
    }
    if (!spd) {
      lambda = fmax(lambda * 10.0, 1e-6);
      if (lambda > 1e12) break;
      continue;
    }
    for (int i = 0; i < 6; ++i) {
      double acc = -Jtr[i];
      for (int k = 0; k < i; ++k) acc -= U[k][i] * d[k];
      d[i] = acc * U[i][i];
    }
    for (int i = 5; i >= 0; --i) {
      double acc = d[i];
      for (int j = i + 1; j < 6; ++j) acc -= U[i][j] * d[j];
      d[i] = acc * U[i][i];
    }
    // candidate: R' = exp([w]x) R, t' = t + dt
    const double wx = d[0], wy = d[1], wz = d[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double A_, B_;  // sin(th)/th, (1 - cos th)/th^2
    if (th < 1e-8) {
      A_ = 1.0 - th2 / 6.0;
      B_ = 0.5 - th2 / 24.0;
    } else {
      A_ = sin(th) / th;
      B_ = (1.0 - cos(th)) / th2;
    }
    const double K1[3][3] = {{0, -wz, wy}, {wz, 0, -wx}, {-wy, wx, 0}};
    double E[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double k2 = 0.0;
        for (int k = 0; k < 3; ++k) k2 += K1[i][k] * K1[k][j];
        E[i][j] = (i == j ? 1.0 : 0.0) + A_ * K1[i][j] + B_ * k2;
      }
    double Rn[3][3], tn[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Rn[i][j] = E[i][0] * R[0][j] + E[i][1] * R[1][j] + E[i][2] * R[2][j];
      tn[i] = t[i] + d[3 + i];
    }
    const double Sn = cost_and_normal(Rn, tn, nullptr, nullptr);
    double dmax = 0.0;
    for (int i = 0; i < 6; ++i) dmax = fmax(dmax, fabs(d[i]));
    if (Sn <= S) {
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i][j] = Rn[i][j];
        t[i] = tn[i];
      }
      lambda = fmax(lambda * 0.1, 1e-12);
      S = cost_and_normal(R, t, JtJ, Jtr);
      if (dmax < 1e-12) break;
    } else {
      lambda *= 10.0;
      if (lambda > 1e12 || dmax < 1e-14) break;
    }
  }
}

// n == 4: cv2's P3P branch (p3p_f64.cuh), kept out of line so that the rare path costs the common one no registers
__device__ __noinline__ bool p3p_four_points(const Camera& cam, const double (*pw)[3], const double (*und)[2], double (&R)[3][3], double (&t)[3]) {
  double X[4][3], us[4][2];
  for (int k = 0; k < 4; ++k) {
    for (int c = 0; c < 3; ++c) X[k][c] = pw[k][c];
    us[k][0] = und[k][0] * cam.fx + cam.cx;
    us[k][1] = und[k][1] * cam.fy + cam.cy;
  }
  return solve_p3p_f64(cam, X, us, R, t);
}

// The kernel is serial-latency bound: its duration is the time ONE frame's dependent chain of
// float64 instructions takes, whatever the batch size (spreading whole frames over more warps was
// measured to buy nothing).  What shortens it is splitting a frame over lanes: every frame gets a
// group of 4 lanes that run the sequential parts redundantly and share MtM accumulation, the four
// eigenvectors and the three beta variants (see epnp_f64).  The 12x12 working matrix exists once per
// frame in shared memory (FrameMat12, 9.3 KB per warp).
__global__ void __maxnreg__(SPE_REFIT_REGS) select_refit_kernel(DevModel m, RansacArgs a, RansacWorkspace ws) {
  const int lane = threadIdx.x & 31, sub = lane & 3;
  const unsigned gmask = 0xFu << (lane & ~3);
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  if (b >= a.B) return;  // whole groups leave together (blockDim is a multiple of 4)
  const int n = ws.n[b];
  const unsigned vis = ws.vis[b];
  int status = SPE_FRAME_OK, winner = -1, budget = 0;
  unsigned inl = 0;
  if (n < 4) {
    status = SPE_FRAME_TOO_FEW_POINTS;
  } else if (n == 4) {
    inl = vis;  // cv2: four points -> solvePnP(SOLVEPNP_P3P) on them, no RANSAC, every point an inlier
    winner = 0;
  } else if (n == kModelPoints) {
    inl = vis;  // cv2: model_points == npoints -> plain solvePnP, every point an inlier
    winner = 0;
  } else if (a.exact) {
    // the float64 replay (ransac_exact.cu) has already run cv2's loop for this frame
    winner = ws.x_winner[b];
    inl = ws.x_mask[b];
    budget = ws.x_visited[b];
    if (winner < 0) status = SPE_FRAME_NO_MODEL;
  } else {
    // sequential acceptance over the FP32 inlier counts (App. B.6); hypothesis h is scored at slot[h]
    const uint8_t* counts = ws.counts + (size_t)b * a.H;
    const uint16_t* slot = m.slot + (size_t)(n - 6) * m.max_hyp;
    // cv2's budget starts at iterationsCount; only the first H hypotheses exist here.  `budget` is what cv2's loop
    // would still want: a value above H means the search was cut short (the caller can re-run with SPE_FLAG_EXACT).
    int niters = max(a.iterations, a.H), max_good = 0;
    for (int h = 0; h < min(niters, a.H); ++h) {
      const int g = counts[slot[h]];
      if (g > max(max_good, kModelPoints - 1)) {
        winner = h;
        max_good = g;
        niters = update_num_iters(a.confidence, (double)(n - g) / n, niters);
      }
    }
    budget = niters;
    if (winner < 0) status = SPE_FRAME_NO_MODEL;
    else inl = ws.masks[(size_t)b * a.H + slot[winner]];
  }
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, t[3] = {0, 0, 0};
  if (status == SPE_FRAME_OK) {
    double pw[kMaxLandmarks][3], und[kMaxLandmarks][2], img[kMaxLandmarks][2];
    int k = 0;
    for (int j = 0; j < m.J; ++j)
      if ((inl >> j) & 1u) {
        for (int c = 0; c < 3; ++c) pw[k][c] = (double)m.landmarks[3 * j + c];
        const float2 px = ws.img[(size_t)b * m.J + j];
        img[k][0] = (double)px.x, img[k][1] = (double)px.y;
        const double2 q = ws.und[(size_t)b * m.J + j];
        // RANSAC's final solve converts the image points to float64 before undistorting; the
        // n == 5 shortcut hands cv2.solvePnP the float32 points, whose undistortion stays float32
        und[k][0] = n <= kModelPoints ? (double)(float)q.x : q.x;
        und[k][1] = n <= kModelPoints ? (double)(float)q.y : q.y;
        ++k;
      }
    extern __shared__ double s_mat[];  // [blockDim.x / 4][kFrameMatStride]
    if (n == 4) {
      if (!p3p_four_points(m.cam, pw, und, R, t)) status = SPE_FRAME_NO_MODEL;  // no real solution: cv2 returns garbage
    } else {
      epnp_f64(k, pw, und, m.cam, R, t, sub, gmask, FrameMat12{s_mat + (threadIdx.x >> 2) * kFrameMatStride});
    }
    if (a.refine_lm && status == SPE_FRAME_OK) refine_lm_f64(k, pw, img, m.cam, R, t);
  }
  if (sub != 0) return;
  double q[4] = {1, 0, 0, 0};
  if (status == SPE_FRAME_OK) rotation_to_quat(R, q);
  float* o = a.pose7 + (size_t)b * 7;
  const bool ok = status == SPE_FRAME_OK;
  for (int i = 0; i < 4; ++i) o[i] = ok ? (float)q[i] : 0.f;
  for (int i = 0; i < 3; ++i) o[4 + i] = ok ? (float)t[i] : 0.f;
  a.inlier_mask[b] = ok ? inl : 0u;
  a.status[b] = status;
  if (a.winner) a.winner[b] = winner;
  if (a.budget) a.budget[b] = budget;
  if (a.rt) {
    double* r = a.rt + (size_t)b * 12;
    for (int i = 0; i < 9; ++i) r[i] = ok ? R[i / 3][i % 3] : 0.0;
    for (int i = 0; i < 3; ++i) r[9 + i] = ok ? t[i] : 0.0;
  }
}

}  // namespace

cudaError_t launch_ransac_select_refit(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0) return cudaSuccess;
  // 4 lanes per frame.  Background tail: the warps are packed 8 to a CTA (8 x 255 registers = a whole SM) so that
  // they hide each other's latency on FEW SMs instead of each blocking a 21 K-register CTA slot of the next
  // batch's hypothesis kernel on EVERY SM (profiles/step_r1.md, tools/overlap_probe.py).  Alone: one warp per CTA.
  constexpr int kMaxTailWarps = 65536 / (32 * ((SPE_REFIT_REGS + 7) / 8 * 8));  // one CTA = the register file of one SM
  // Without a common carveout the kernel flips idle SMs to an all-L1 split and the next batch's
  // decode CTAs must wait for it to finish: measured 0.12 -> 0.32 ms decode when overlapped.
  static PerDeviceOnce once;
  const cudaError_t ce = once.run(m.device, [] {
    cudaError_t r = cudaFuncSetAttribute(select_refit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess)
      r = cudaFuncSetAttribute(select_refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxTailWarps * 8 * kFrameMatStride * (int)sizeof(double));
    return r;
  });
  if (ce != cudaSuccess) return ce;
#ifndef SPE_TAIL_WARPS
#define SPE_TAIL_WARPS kMaxTailWarps
#endif
  int warps = a.refit_background ? SPE_TAIL_WARPS : 1;
  if (a.tail_warps >= 1 && a.tail_warps <= kMaxTailWarps) warps = a.tail_warps;
  const int threads = 32 * warps;
  const int ctas = (a.B * 4 + threads - 1) / threads;
  select_refit_kernel<<<ctas, threads, sizeof(double) * kFrameMatStride * 8 * warps, stream>>>(dev_model(m), a, ws);
  return cudaGetLastError();
}

}  // namespace spe
