// cv2.solvePnPRansac's n == 4 branch: with exactly four points OpenCV does not run RANSAC but calls
// solvePnP(..., SOLVEPNP_P3P): the perspective-three-point problem on the first three points, the fourth picks among the
// (up to four) solutions by its reprojection error in ideal pixels.  Any exact P3P solver returns the same solution set
// (OpenCV's own P3P and AP3P agree to 1e-9 on the poses, measured); this one follows the classical reduction
// (Fischler & Bolles / Grunert): with d2 = u d1, d3 = v d1 the three distance equations give u as a rational function of
// v and a quartic in v, whose coefficients are built by polynomial arithmetic and whose real roots come from Ferrari's
// closed form polished by Newton steps.  float64, __host__ __device__ (tests/host/exact_eval_host.cu checks it against
// cv2 without a GPU).  Rare path: one thread per frame, performance is irrelevant.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "epnp_math.cuh"
#include "ransac.cuh"

namespace spe {

namespace p3p {

SPE_HD inline double cbrt_signed(double x) { return x < 0 ? -pow(-x, 1.0 / 3.0) : pow(x, 1.0 / 3.0); }

// one real root of t^3 + a t^2 + b t + c (the largest one when there are three)
SPE_HD inline double cubic_real_root(double a, double b, double c) {
  const double q = (a * a - 3.0 * b) / 9.0, r = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
  double t;
  if (r * r < q * q * q) {
    const double th = acos(fmin(fmax(r / sqrt(q * q * q), -1.0), 1.0));
    t = -2.0 * sqrt(q) * cos((th + 2.0 * 3.14159265358979323846) / 3.0) - a / 3.0;  // the largest of the three
    const double t2 = -2.0 * sqrt(q) * cos(th / 3.0) - a / 3.0, t3 = -2.0 * sqrt(q) * cos((th - 2.0 * 3.14159265358979323846) / 3.0) - a / 3.0;
    t = fmax(t, fmax(t2, t3));
  } else {
    const double A = -copysign(cbrt_signed(fabs(r) + sqrt(r * r - q * q * q)), r);
    const double B = A != 0.0 ? q / A : 0.0;
    t = A + B - a / 3.0;
  }
  return t;
}

// real roots of c4 x^4 + c3 x^3 + c2 x^2 + c1 x + c0 (Ferrari), polished by Newton; returns their number
SPE_HD inline int quartic_real_roots(const double (&c)[5], double (&x)[4]) {
  int n = 0;
  if (fabs(c[4]) < 1e-300) return 0;
  const double a = c[3] / c[4], b = c[2] / c[4], cc = c[1] / c[4], d = c[0] / c[4];
  // depressed quartic y^4 + p y^2 + q y + r, x = y - a/4
  const double a2 = a * a;
  const double p = b - 3.0 * a2 / 8.0, q = cc - a * b / 2.0 + a2 * a / 8.0, r = d - a * cc / 4.0 + a2 * b / 16.0 - 3.0 * a2 * a2 / 256.0;
  double y[4];
  if (fabs(q) < 1e-14 * (1.0 + fabs(p) * sqrt(fabs(p)) + fabs(r))) {  // biquadratic
    const double disc = p * p - 4.0 * r;
    if (disc >= 0) {
      const double s = sqrt(disc);
      const double z[2] = {(-p + s) / 2.0, (-p - s) / 2.0};
      for (int k = 0; k < 2; ++k)
        if (z[k] >= 0) y[n++] = sqrt(z[k]), y[n++] = -sqrt(z[k]);
    }
  } else {
    // resolvent cubic m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, m > 0
    const double m = cubic_real_root(p, p * p / 4.0 - r, -q * q / 8.0);
    if (m > 0) {
      const double s = sqrt(2.0 * m);
      double t1 = -(2.0 * p + 2.0 * m + 2.0 * q / s), t2 = -(2.0 * p + 2.0 * m - 2.0 * q / s);
      // a (nearly) double root shows up as a discriminant that rounding pushes just below zero: keep it
      const double tol = 1e-9 * (fabs(p) + fabs(m) + fabs(q / s));
      if (t1 < 0 && t1 > -tol) t1 = 0;
      if (t2 < 0 && t2 > -tol) t2 = 0;
      if (t1 >= 0) y[n++] = (s + sqrt(t1)) / 2.0, y[n++] = (s - sqrt(t1)) / 2.0;
      if (t2 >= 0) y[n++] = (-s + sqrt(t2)) / 2.0, y[n++] = (-s - sqrt(t2)) / 2.0;
    }
  }
  for (int k = 0; k < n; ++k) {
    double xr = y[k] - a / 4.0;
    for (int it = 0; it < 4; ++it) {  // Newton on the original polynomial
      const double f = (((c[4] * xr + c[3]) * xr + c[2]) * xr + c[1]) * xr + c[0];
      const double fp = ((4.0 * c[4] * xr + 3.0 * c[3]) * xr + 2.0 * c[2]) * xr + c[1];
      if (fabs(fp) < 1e-300) break;
      xr -= f / fp;
    }
    x[k] = xr;
  }
  return n;
}

SPE_HD inline void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1], o[1] = a[2] * b[0] - a[0] * b[2], o[2] = a[0] * b[1] - a[1] * b[0];
}
SPE_HD inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
SPE_HD inline void normalise3(double* a) {
  const double n = sqrt(dot3(a, a));
  a[0] /= n, a[1] /= n, a[2] /= n;
}
// orthonormal frame (columns e0, e1, e2) spanned by two edge vectors
SPE_HD inline void frame_of(const double* p, const double* q, double (&F)[3][3]) {
  double e0[3] = {p[0], p[1], p[2]}, e2[3], e1[3];
  normalise3(e0);
  cross3(e0, q, e2);
  normalise3(e2);
  cross3(e2, e0, e1);
  for (int r = 0; r < 3; ++r) F[r][0] = e0[r], F[r][1] = e1[r], F[r][2] = e2[r];
}

}  // namespace p3p

// X [4][3] object points, us [4][2] ideal pixel coordinates (undistorted, K applied).  Returns false when the first three
// points admit no real solution (cv2: solvePnP returns false).
SPE_HD inline bool solve_p3p_f64(const Camera& cam, const double (&X)[4][3], const double (&us)[4][2], double (&R)[3][3], double (&t)[3]) {
  using namespace p3p;
  double f[3][3];
  for (int i = 0; i < 3; ++i) {
    f[i][0] = (us[i][0] - cam.cx) / cam.fx, f[i][1] = (us[i][1] - cam.cy) / cam.fy, f[i][2] = 1.0;
    normalise3(f[i]);
  }
  const double c12 = dot3(f[0], f[1]), c13 = dot3(f[0], f[2]), c23 = dot3(f[1], f[2]);
  double e12[3], e13[3], e23[3];
  for (int c = 0; c < 3; ++c) e12[c] = X[1][c] - X[0][c], e13[c] = X[2][c] - X[0][c], e23[c] = X[2][c] - X[1][c];
  const double A12 = dot3(e12, e12), A13 = dot3(e13, e13), A23 = dot3(e23, e23);
  if (!(A12 > 0 && A13 > 0 && A23 > 0)) return false;
  // u = N(v) / D(v):  D = 2 A13 (c23 v - c12),  N = (A12 - A23)(1 - 2 c13 v + v^2) - A13 (1 - v^2)
  const double K = A12 - A23;
  const double N[3] = {K - A13, -2.0 * K * c13, K + A13};      // ascending powers of v
  const double D[2] = {-2.0 * A13 * c12, 2.0 * A13 * c23};
  // A13 (D^2 + N^2 - 2 c12 N D) - A12 (1 - 2 c13 v + v^2) D^2 = 0
  double D2[3] = {D[0] * D[0], 2.0 * D[0] * D[1], D[1] * D[1]};
  double N2[5] = {N[0] * N[0], 2.0 * N[0] * N[1], N[1] * N[1] + 2.0 * N[0] * N[2], 2.0 * N[1] * N[2], N[2] * N[2]};
  double ND[4] = {N[0] * D[0], N[0] * D[1] + N[1] * D[0], N[1] * D[1] + N[2] * D[0], N[2] * D[1]};
  const double S[3] = {1.0, -2.0 * c13, 1.0};
  double SD2[5] = {S[0] * D2[0], S[0] * D2[1] + S[1] * D2[0], S[0] * D2[2] + S[1] * D2[1] + S[2] * D2[0], S[1] * D2[2] + S[2] * D2[1], S[2] * D2[2]};
  double q[5];
  for (int k = 0; k < 5; ++k) q[k] = A13 * ((k < 3 ? D2[k] : 0.0) + N2[k] - 2.0 * c12 * (k < 4 ? ND[k] : 0.0)) - A12 * SD2[k];
  double scale = 0.0;
  for (int k = 0; k < 5; ++k) scale = fmax(scale, fabs(q[k]));
  if (!(scale > 0)) return false;
  for (int k = 0; k < 5; ++k) q[k] /= scale;
  double v4[4];
  const int nroots = quartic_real_roots(q, v4);
  bool found = false;
  double best = 0.0;
  double Fw[3][3];
  frame_of(e12, e13, Fw);
  for (int k = 0; k < nroots; ++k) {
    const double v = v4[k];
    if (!(v > 0)) continue;
    const double Dv = D[0] + D[1] * v;
    if (fabs(Dv) < 1e-14 * A13) continue;
    const double u = (N[0] + (N[1] + N[2] * v) * v) / Dv;
    if (!(u > 0)) continue;
    const double den = 1.0 + u * u - 2.0 * u * c12;
    if (!(den > 0)) continue;
    const double d1 = sqrt(A12 / den), d2 = u * d1, d3 = v * d1;
    double P[3][3];
    for (int c = 0; c < 3; ++c) P[0][c] = d1 * f[0][c], P[1][c] = d2 * f[1][c], P[2][c] = d3 * f[2][c];
    double g12[3], g13[3], Fc[3][3];
    for (int c = 0; c < 3; ++c) g12[c] = P[1][c] - P[0][c], g13[c] = P[2][c] - P[0][c];
    frame_of(g12, g13, Fc);
    double Rk[3][3], tk[3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Rk[r][c] = Fc[r][0] * Fw[c][0] + Fc[r][1] * Fw[c][1] + Fc[r][2] * Fw[c][2];
    for (int r = 0; r < 3; ++r) tk[r] = P[0][r] - (Rk[r][0] * X[0][0] + Rk[r][1] * X[0][1] + Rk[r][2] * X[0][2]);
    // the fourth point decides (squared reprojection error in ideal pixels)
    const double xc = Rk[0][0] * X[3][0] + Rk[0][1] * X[3][1] + Rk[0][2] * X[3][2] + tk[0];
    const double yc = Rk[1][0] * X[3][0] + Rk[1][1] * X[3][1] + Rk[1][2] * X[3][2] + tk[1];
    const double zc = Rk[2][0] * X[3][0] + Rk[2][1] * X[3][1] + Rk[2][2] * X[3][2] + tk[2];
    const double du = cam.cx + cam.fx * xc / zc - us[3][0], dv = cam.cy + cam.fy * yc / zc - us[3][1];
    const double err = du * du + dv * dv;
    if (!found || err < best) {
      found = true;
      best = err;
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) R[r][c] = Rk[r][c];
        t[r] = tk[r];
      }
    }
  }
  return found;
}

}  // namespace spe
