// float64 pieces shared by the refit (ransac_refit.cu) and the exact replay (ransac_exact.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "epnp_math.cuh"

namespace spe {

// OpenCV's JacobiSVD on the rows of a symmetric n x n matrix (App. B.4): returns the rotated rows
// normalised (= rows of U^T) sorted by descending singular value.  Sign-defining for the PCA axes.
SPE_HD inline void cv_jacobi_rows(double* A, double* w, int n) {
  const double eps = 2.220446049250313e-16 * 10;
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < n; ++k) sd += A[i * n + k] * A[i * n + k];
    w[i] = sd;
  }
  const int max_iter = n > 30 ? n : 30;
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double* Ai = A + i * n;
        double* Aj = A + j * n;
        const double a = w[i], b = w[j];
        double p = 0;
        for (int k = 0; k < n; ++k) p += Ai[k] * Aj[k];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        double na = 0, nb = 0;
        for (int k = 0; k < n; ++k) {
          const double t0 = c * Ai[k] + s * Aj[k];
          const double t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          na += t0 * t0;
          nb += t1 * t1;
        }
        w[i] = na;
        w[j] = nb;
        changed = true;
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < n; ++k) sd += A[i * n + k] * A[i * n + k];
    w[i] = sqrt(sd);
  }
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (w[j] < w[k]) j = k;
    if (i != j) {
      const double tw = w[i];
      w[i] = w[j];
      w[j] = tw;
      for (int k = 0; k < n; ++k) {
        const double ta = A[i * n + k];
        A[i * n + k] = A[j * n + k];
        A[j * n + k] = ta;
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    const double s = w[i] > 2.2250738585072014e-308 ? 1.0 / w[i] : 0.0;
    for (int k = 0; k < n; ++k) A[i * n + k] *= s;
  }
}

}  // namespace spe
