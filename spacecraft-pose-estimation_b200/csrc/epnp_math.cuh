// Small dense linear algebra shared by the RANSAC-EPnP kernels (ransac_epnp.cu).
//
// Everything is templated on the scalar type: float for the per-hypothesis kernel (FP32 CUDA
// cores, everything in registers, loops fully unrolled), double for the one final refit per frame.
// The algorithm is OpenCV's EPnP (calib3d, un-vendored dependency of the reference; restated in
// SURVEY.md App. B.3 and oracle/epnp_ref.py): control points -> M -> null-space basis -> three
// linearised beta initialisations -> 5 Gauss-Newton steps each -> Procrustes -> best of three.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace spe {

// SFU approximations (rcp/sqrt/rsqrt.approx.ftz, ~1 ulp, no slow-path branches).  The FP32 path
// only scores hypotheses against a 15 px threshold; the one result that is returned to the caller
// is refit in float64.
// (SPE_HD: the float64 instantiations are also compiled for the host by tests/host/exact_eval_host.cu, which checks the
// replay's arithmetic against cv2 without a GPU; the host versions of the approximations below exist only to compile.)
#define SPE_HD __host__ __device__
SPE_HD __forceinline__ float rcp_approx(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
SPE_HD __forceinline__ float sqrt_approx(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return sqrtf(x);
#endif
}
SPE_HD __forceinline__ float rsqrt_approx(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / sqrtf(x);
#endif
}
template <typename T>
struct Real;
template <>
struct Real<float> {
  static SPE_HD __forceinline__ float sqrt(float x) { return sqrt_approx(x); }
  static SPE_HD __forceinline__ float rsqrt(float x) { return rsqrt_approx(x); }
  static SPE_HD __forceinline__ float abs(float x) { return fabsf(x); }
  static SPE_HD __forceinline__ float rcp(float x) { return rcp_approx(x); }
  static SPE_HD __forceinline__ float div(float a, float b) { return a * rcp_approx(b); }
  static SPE_HD __forceinline__ float copysign(float m, float s) { return copysignf(m, s); }
  static constexpr float eps = 1.1920929e-7f;
  static constexpr float tiny = 1e-30f;
  static constexpr float pivot_floor = 1e-7f;  // relative size below which a QR pivot counts as zero
  static constexpr int svd3_sweeps = 4;  // converged to rounding in 4 (200k random FP32 cases, worst |dR| 2.6e-6)
};
template <>
struct Real<double> {
  static SPE_HD __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static SPE_HD __forceinline__ double rsqrt(double x) {
#ifdef __CUDA_ARCH__
    return ::rsqrt(x);
#else
    return 1.0 / ::sqrt(x);
#endif
  }
  static SPE_HD __forceinline__ double abs(double x) { return fabs(x); }
  static SPE_HD __forceinline__ double rcp(double x) { return 1.0 / x; }
  static SPE_HD __forceinline__ double div(double a, double b) { return a / b; }
  static SPE_HD __forceinline__ double copysign(double m, double s) { return ::copysign(m, s); }
  static constexpr double eps = 2.220446049250313e-16;
  static constexpr double tiny = 1e-280;
  static constexpr double pivot_floor = 1e-15;
  static constexpr int svd3_sweeps = 8;
};

// Jacobi rotation that orthogonalises two columns with squared norms (a, b) and inner product p:
// returns (c, s, t = s/c) of  x' = c x - s y,  y' = s x + c y;  new norms a - t p, b + t p.
template <typename T>
SPE_HD __forceinline__ void jacobi_angle(T a, T b, T p, T& c, T& s, T& t) {
  const T zeta = (b - a) / (T(2) * p);
  t = Real<T>::copysign(T(1), zeta) / (Real<T>::abs(zeta) + Real<T>::sqrt(T(1) + zeta * zeta));
  c = Real<T>::rsqrt(T(1) + t * t);
  s = c * t;
}

// FP32: 3 MUFU + ~8 FP32 ops.  A rotation only has to be orthogonal to rounding accuracy, which
// c = rsqrt(1 + t^2), s = c t guarantees irrespective of how exact t is.
SPE_HD __forceinline__ void jacobi_angle_fast(float a, float b, float p, float& c, float& s, float& t) {
  // t = 2p / (h + sign(h) sqrt(h^2 + 4p^2)), h = b - a  ==  sign(zeta) / (|zeta| + sqrt(zeta^2 + 1)) with zeta = h / 2p,
  // in a 3-MUFU dependent chain (sqrt, rcp, rsqrt) instead of 4
  const float h = b - a, gg = p + p;
  const float q = sqrt_approx(fmaf(h, h, fmaf(gg, gg, 1e-37f)));  // h = p = 0 -> t = 0
  t = gg * rcp_approx(h + copysignf(q, h));
  c = rsqrt_approx(fmaf(t, t, 1.0f));
  s = c * t;
}
template <>
SPE_HD __forceinline__ void jacobi_angle<float>(float a, float b, float p, float& c, float& s, float& t) {
  jacobi_angle_fast(a, b, p, c, s, t);
}

// Householder least squares min |A x - b| for a tiny R x C system held in registers.
// Zero (masked) columns are skipped and get x = 0.  A and b are overwritten.
template <typename T, int R, int C>
SPE_HD __forceinline__ void lsq_householder(T (&A)[R][C], T (&b)[R], T (&x)[C]) {
  T diag[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    T s2 = T(0);
#pragma unroll
    for (int i = k; i < R; ++i) s2 += A[i][k] * A[i][k];
    const T norm = Real<T>::sqrt(s2);
    const T akk = A[k][k];
    const T alpha = akk > T(0) ? -norm : norm;  // R_kk
    const T vk = akk - alpha;
    const T denom = -alpha * vk;  // = v^T v / 2  (>= norm^2)
    const T inv = denom > Real<T>::tiny ? Real<T>::rcp(denom) : T(0);
    A[k][k] = vk;
#pragma unroll
    for (int j = k + 1; j < C; ++j) {
      T dot = T(0);
#pragma unroll
      for (int i = k; i < R; ++i) dot += A[i][k] * A[i][j];
      const T tau = dot * inv;
#pragma unroll
      for (int i = k; i < R; ++i) A[i][j] -= tau * A[i][k];
    }
    T dot = T(0);
#pragma unroll
    for (int i = k; i < R; ++i) dot += A[i][k] * b[i];
    const T tau = dot * inv;
#pragma unroll
    for (int i = k; i < R; ++i) b[i] -= tau * A[i][k];
    diag[k] = alpha;
  }
#pragma unroll
  for (int k = C - 1; k >= 0; --k) {
    T acc = b[k];
#pragma unroll
    for (int j = k + 1; j < C; ++j) acc -= A[k][j] * x[j];
    x[k] = Real<T>::abs(diag[k]) > Real<T>::tiny ? Real<T>::div(acc, diag[k]) : T(0);
  }
}

// Least squares through the normal equations + Cholesky for the FP32 Gauss-Newton step: the
// step is a small correction of an iteration that re-linearises five times, so the squared
// conditioning is harmless there (agreement with cv2 unchanged, tests/test_pnp_gpu.py) and it costs
// less than half of the Householder version.
template <typename T, int R, int C>
SPE_HD __forceinline__ void lsq_normal(const T (&A)[R][C], const T (&b)[R], T (&x)[C]) {
  T G[C][C], y[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
#pragma unroll
    for (int j = i; j < C; ++j) {
      T acc = T(0);
#pragma unroll
      for (int r = 0; r < R; ++r) acc += A[r][i] * A[r][j];
      G[i][j] = acc;
    }
    T acc = T(0);
#pragma unroll
    for (int r = 0; r < R; ++r) acc += A[r][i] * b[r];
    y[i] = acc;
  }
  // Cholesky G = U^T U (upper), in place; inv[i] = 1 / U_ii
  T inv[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
    T dgn = G[i][i];
#pragma unroll
    for (int k = 0; k < i; ++k) dgn -= G[k][i] * G[k][i];
    inv[i] = dgn > Real<T>::tiny ? Real<T>::rsqrt(dgn) : T(0);
#pragma unroll
    for (int j = i + 1; j < C; ++j) {
      T acc = G[i][j];
#pragma unroll
      for (int k = 0; k < i; ++k) acc -= G[k][i] * G[k][j];
      G[i][j] = acc * inv[i];
    }
  }
#pragma unroll
  for (int i = 0; i < C; ++i) {  // U^T z = y
    T acc = y[i];
#pragma unroll
    for (int k = 0; k < i; ++k) acc -= G[k][i] * y[k];
    y[i] = acc * inv[i];
  }
#pragma unroll
  for (int i = C - 1; i >= 0; --i) {  // U x = z
    T acc = y[i];
#pragma unroll
    for (int j = i + 1; j < C; ++j) acc -= G[i][j] * x[j];
    x[i] = acc * inv[i];
  }
}

// L (6x10) from the four null-space candidates v[i][0..11] (four 3-vectors each); App. B.3g.
// kDoubledDiag (FP32 hypothesis path): the squared terms are stored doubled too, L[k][{0,2,5,9}] = 2 |d_i|^2,
// which turns a Gauss-Newton Jacobian row into four plain dot products (gauss_newton_doubled below).
template <typename T, bool kDoubledDiag = false>
SPE_HD __forceinline__ void build_L(const T (&v)[4][12], T (&L)[6][10]) {
  constexpr int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    T d[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) d[i][c] = v[i][3 * pa[k] + c] - v[i][3 * pb[k] + c];
    auto dot = [&](int i, int j) { return d[i][0] * d[j][0] + d[i][1] * d[j][1] + d[i][2] * d[j][2]; };
    constexpr T dg = kDoubledDiag ? T(2) : T(1);
    L[k][0] = dg * dot(0, 0);
    L[k][1] = T(2) * dot(0, 1);
    L[k][2] = dg * dot(1, 1);
    L[k][3] = T(2) * dot(0, 2);
    L[k][4] = T(2) * dot(1, 2);
    L[k][5] = dg * dot(2, 2);
    L[k][6] = T(2) * dot(0, 3);
    L[k][7] = T(2) * dot(1, 3);
    L[k][8] = T(2) * dot(2, 3);
    L[k][9] = dg * dot(3, 3);
  }
}

// rho_k = |c_a - c_b|^2 over the six control-point pairs.
template <typename T>
SPE_HD __forceinline__ void build_rho(const T (&cws)[4][3], T (&rho)[6]) {
  constexpr int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const T dx = cws[pa[k]][0] - cws[pb[k]][0], dy = cws[pa[k]][1] - cws[pb[k]][1], dz = cws[pa[k]][2] - cws[pb[k]][2];
    rho[k] = dx * dx + dy * dy + dz * dz;
  }
}

// The three linearised initialisations of EPnP (App. B.3h), variant = 1, 2 or 3, as one piece
// of straight-line code so that lanes running different variants do not diverge.
// kDoubledDiag: L comes from build_L<T, true>; a doubled column halves its unknown (exactly, a power
// of two), which is undone after the solve.
template <typename T, bool kNormalEq = false, bool kDoubledDiag = false>
SPE_HD __forceinline__ void approx_betas(const T (&L)[6][10], const T (&rho)[6], int variant, T (&betas)[4]) {
  T A[6][5], b[6], x[5];
  const bool v1 = variant == 1, v3 = variant == 3;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    A[k][0] = L[k][0];
    A[k][1] = L[k][1];
    A[k][2] = v1 ? L[k][3] : L[k][2];
    A[k][3] = v1 ? L[k][6] : (v3 ? L[k][3] : T(0));
    A[k][4] = v3 ? L[k][4] : T(0);
    b[k] = rho[k];
  }
  if constexpr (kNormalEq) lsq_normal<T, 6, 5>(A, b, x);  // masked (all-zero) columns come out as x = 0 in both
  else lsq_householder<T, 6, 5>(A, b, x);
  if constexpr (kDoubledDiag) {
    x[0] *= T(2);             // column L[.][0]
    if (!v1) x[2] *= T(2);    // column L[.][2] (variants 2, 3)
  }
  const bool neg = x[0] < T(0);
  const T b0mag = Real<T>::sqrt(Real<T>::abs(x[0]));
  if (v1) {
    const T sg = neg ? T(-1) : T(1);
    const T inv0 = sg * Real<T>::rcp(b0mag);
    betas[0] = b0mag;
    betas[1] = x[1] * inv0;
    betas[2] = x[2] * inv0;
    betas[3] = x[3] * inv0;
  } else {
    const bool same_sign = neg ? (x[2] < T(0)) : (x[2] > T(0));
    T b0 = b0mag;
    if (x[1] < T(0)) b0 = -b0;
    betas[0] = b0;
    betas[1] = same_sign ? Real<T>::sqrt(Real<T>::abs(x[2])) : T(0);
    betas[2] = v3 ? Real<T>::div(x[3], b0) : T(0);
    betas[3] = T(0);
  }
}

// Exactly five Gauss-Newton steps on the six distance constraints (App. B.3i).
// kNormalEq selects the normal-equation solve (FP32 hypotheses) over Householder QR (float64 refit).
template <typename T, bool kNormalEq = false>
SPE_HD __forceinline__ void gauss_newton(const T (&L)[6][10], const T (&rho)[6], T (&be)[4]) {
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    T A[6][4], r[6], x[4];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const T* l = L[k];
      A[k][0] = T(2) * l[0] * be[0] + l[1] * be[1] + l[3] * be[2] + l[6] * be[3];
      A[k][1] = l[1] * be[0] + T(2) * l[2] * be[1] + l[4] * be[2] + l[7] * be[3];
      A[k][2] = l[3] * be[0] + l[4] * be[1] + T(2) * l[5] * be[2] + l[8] * be[3];
      A[k][3] = l[6] * be[0] + l[7] * be[1] + l[8] * be[2] + T(2) * l[9] * be[3];
      r[k] = rho[k] - (l[0] * be[0] * be[0] + l[1] * be[0] * be[1] + l[2] * be[1] * be[1] + l[3] * be[0] * be[2] +
                       l[4] * be[1] * be[2] + l[5] * be[2] * be[2] + l[6] * be[0] * be[3] + l[7] * be[1] * be[3] +
                       l[8] * be[2] * be[3] + l[9] * be[3] * be[3]);
    }
    if constexpr (kNormalEq) lsq_normal<T, 6, 4>(A, r, x);
    else lsq_householder<T, 6, 4>(A, r, x);
#pragma unroll
    for (int i = 0; i < 4; ++i) be[i] += x[i];
  }
}

// FP32 hypothesis path: L comes with doubled squared terms (build_L<T, true>), so a Jacobian row is four plain
// 4-term dot products and beta^T Q_k beta = (J_k . beta) / 2 gives the residual from it (126 instead of 204
// instructions per step).
// The five steps for NB beta vectors at once (the three variants of one hypothesis).  The 6x4 system of a
// step is never stored: each Jacobian row goes straight into the packed normal equations G (10) and y (4),
// and the NB Cholesky factorisations / substitutions — short loops dominated by dependent rsqrt and FMA
// chains — run interleaved, which is the point: NB independent chains per thread for the scheduler.
template <int NB>
SPE_HD __forceinline__ void gauss_newton_doubled_batch(const float (&L)[6][10], const float (&rho)[6], float (&be)[NB][4]) {
#pragma unroll 1
  for (int it = 0; it < 5; ++it) {
    float G[NB][4][4], y[NB][4];  // upper triangles only
#pragma unroll
    for (int m = 0; m < NB; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        y[m][i] = 0.f;
#pragma unroll
        for (int j = i; j < 4; ++j) G[m][i][j] = 0.f;
      }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float* l = L[k];
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        const float* b = be[m];
        float a[4];
        a[0] = fmaf(l[6], b[3], fmaf(l[3], b[2], fmaf(l[1], b[1], l[0] * b[0])));
        a[1] = fmaf(l[7], b[3], fmaf(l[4], b[2], fmaf(l[2], b[1], l[1] * b[0])));
        a[2] = fmaf(l[8], b[3], fmaf(l[5], b[2], fmaf(l[4], b[1], l[3] * b[0])));
        a[3] = fmaf(l[9], b[3], fmaf(l[8], b[2], fmaf(l[7], b[1], l[6] * b[0])));
        const float q2 = fmaf(a[3], b[3], fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0])));
        const float r = fmaf(-0.5f, q2, rho[k]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          y[m][i] = fmaf(a[i], r, y[m][i]);
#pragma unroll
          for (int j = i; j < 4; ++j) G[m][i][j] = fmaf(a[i], a[j], G[m][i][j]);
        }
      }
    }
    // Cholesky G = U^T U in place (inv = 1 / U_ii), U^T z = y, U x = z; m innermost
    float inv[NB][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        float dgn = G[m][i][i];
#pragma unroll
        for (int k = 0; k < i; ++k) dgn = fmaf(-G[m][k][i], G[m][k][i], dgn);
        inv[m][i] = dgn > Real<float>::tiny ? rsqrt_approx(dgn) : 0.f;
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
          float acc = G[m][i][j];
#pragma unroll
          for (int k = 0; k < i; ++k) acc = fmaf(-G[m][k][i], G[m][k][j], acc);
          G[m][i][j] = acc * inv[m][i];
        }
        float z = y[m][i];
#pragma unroll
        for (int k = 0; k < i; ++k) z = fmaf(-G[m][k][i], y[m][k], z);
        y[m][i] = z * inv[m][i];
      }
    }
#pragma unroll
    for (int i = 3; i >= 0; --i) {
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        float acc = y[m][i];
#pragma unroll
        for (int j = i + 1; j < 4; ++j) acc = fmaf(-G[m][i][j], y[m][j], acc);
        y[m][i] = acc * inv[m][i];  // y now holds x
      }
    }
#pragma unroll
    for (int m = 0; m < NB; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) be[m][i] += y[m][i];
  }
}

// R = U V^T of the 3x3 matrix A = U S V^T (orthogonal Procrustes factor), by one-sided Jacobi.
// The left vector of the smallest singular value is rebuilt as a cross product so that nearly
// planar configurations stay orthonormal; its sign follows the rotated column, which preserves
// det(U V^T) exactly as a full SVD would give it.
template <typename T>
SPE_HD __forceinline__ void procrustes_uvt(const T (&A)[3][3], T (&R)[3][3]) {
  T B[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      B[i][j] = A[i][j];
      V[i][j] = i == j ? T(1) : T(0);
    }
#pragma unroll 1
  for (int sweep = 0; sweep < Real<T>::svd3_sweeps; ++sweep) {
    bool any = false;  // a sweep without a rotation: converged (every later sweep would be the identity too)
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const T a = B[0][p] * B[0][p] + B[1][p] * B[1][p] + B[2][p] * B[2][p];
      const T b = B[0][q] * B[0][q] + B[1][q] * B[1][q] + B[2][q] * B[2][q];
      const T g = B[0][p] * B[0][q] + B[1][p] * B[1][q] + B[2][p] * B[2][q];
      const bool rot = g * g > (Real<T>::eps * Real<T>::eps) * a * b;
      any = any || rot;
      T c, s, t;
      jacobi_angle<T>(a, b, rot ? g : T(1), c, s, t);
      c = rot ? c : T(1);
      s = rot ? s : T(0);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const T x = B[r][p], y = B[r][q];
        B[r][p] = c * x - s * y;
        B[r][q] = s * x + c * y;
        const T vx = V[r][p], vy = V[r][q];
        V[r][p] = c * vx - s * vy;
        V[r][q] = s * vx + c * vy;
      }
    }
    if (!any) break;
  }
  T n2[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) n2[j] = B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j];
  const int jmin = (n2[0] <= n2[1] && n2[0] <= n2[2]) ? 0 : (n2[1] <= n2[2] ? 1 : 2);
  T U[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const T inv = Real<T>::rsqrt(n2[j] > Real<T>::tiny ? n2[j] : T(1));
#pragma unroll
    for (int r = 0; r < 3; ++r) U[r][j] = B[r][j] * inv;
  }
  // rebuild column jmin from the other two (cyclic order keeps the orientation bookkeeping simple)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j == jmin) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      T cx = U[1][j1] * U[2][j2] - U[2][j1] * U[1][j2];
      T cy = U[2][j1] * U[0][j2] - U[0][j1] * U[2][j2];
      T cz = U[0][j1] * U[1][j2] - U[1][j1] * U[0][j2];
      const T along = cx * B[0][j] + cy * B[1][j] + cz * B[2][j];
      const T sg = along < T(0) ? T(-1) : T(1);
      U[0][j] = sg * cx;
      U[1][j] = sg * cy;
      U[2][j] = sg * cz;
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) R[r][c] = U[r][0] * V[c][0] + U[r][1] * V[c][1] + U[r][2] * V[c][2];
  // OpenCV's handling of a reflection: negate the third ROW of R (App. B.3j)
  const T det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
  if (det < T(0)) {
    R[2][0] = -R[2][0];
    R[2][1] = -R[2][1];
    R[2][2] = -R[2][2];
  }
}

// The same for NB independent matrices at once (FP32 hypothesis path: the three beta variants of one
// hypothesis).  A single Procrustes is one long dependent chain (dot products -> 3 MUFU -> rotation,
// 12 times); interleaving NB of them gives the scheduler NB independent chains per thread.
template <int NB>
SPE_HD __forceinline__ void procrustes_uvt_batch(const float (&A)[NB][3][3], float (&R)[NB][3][3]) {
  // One-sided Jacobi on B = A V without accumulating V: the rotated columns are b_j = sigma_j u_j, and the right
  // vectors follow afterwards as v_j = A^T b_j / sigma_j^2 for the two largest columns; the third is their cross
  // product (V is a product of rotations, det +1), exactly as the third u is rebuilt from the other two.
  float B[NB][3][3];
#pragma unroll
  for (int m = 0; m < NB; ++m)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) B[m][i][j] = A[m][i][j];
#pragma unroll 1
  for (int sweep = 0; sweep < Real<float>::svd3_sweeps; ++sweep) {
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      float c[NB], s[NB];
#pragma unroll
      for (int m = 0; m < NB; ++m) {
        const float a = B[m][0][p] * B[m][0][p] + B[m][1][p] * B[m][1][p] + B[m][2][p] * B[m][2][p];
        const float b = B[m][0][q] * B[m][0][q] + B[m][1][q] * B[m][1][q] + B[m][2][q] * B[m][2][q];
        const float g = B[m][0][p] * B[m][0][q] + B[m][1][p] * B[m][1][q] + B[m][2][p] * B[m][2][q];
        // no "already orthogonal" test: a negligible g gives a negligible t (jacobi_angle_fast keeps 0/0 away)
        float t;
        jacobi_angle_fast(a, b, g, c[m], s[m], t);
      }
#pragma unroll
      for (int m = 0; m < NB; ++m)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float x = B[m][r][p], y = B[m][r][q];
          B[m][r][p] = c[m] * x - s[m] * y;
          B[m][r][q] = s[m] * x + c[m] * y;
        }
    }
  }
#pragma unroll
  for (int m = 0; m < NB; ++m) {
    float n2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) n2[j] = B[m][0][j] * B[m][0][j] + B[m][1][j] * B[m][1][j] + B[m][2][j] * B[m][2][j];
    const int jmin = (n2[0] <= n2[1] && n2[0] <= n2[2]) ? 0 : (n2[1] <= n2[2] ? 1 : 2);
    float U[3][3], V[3][3];  // V[c][j] = component c of v_j
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const bool ok = n2[j] > Real<float>::tiny;
      const float inv = rsqrt_approx(ok ? n2[j] : 1.0f);
#pragma unroll
      for (int r = 0; r < 3; ++r) U[r][j] = B[m][r][j] * inv;
      // v_j = A^T u_j / sigma_j = A^T b_j / sigma_j^2;  inv2 * sigma_j = 1 / sigma_j^2
      const float w = ok ? inv * inv : 0.0f;
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
        V[cc][j] = (A[m][0][cc] * B[m][0][j] + A[m][1][cc] * B[m][1][j] + A[m][2][cc] * B[m][2][j]) * w;
    }
    // rebuild column jmin of U and of V from the other two (cyclic order keeps the orientation bookkeeping simple)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (j == jmin) {
        const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        const float cx = U[1][j1] * U[2][j2] - U[2][j1] * U[1][j2];
        const float cy = U[2][j1] * U[0][j2] - U[0][j1] * U[2][j2];
        const float cz = U[0][j1] * U[1][j2] - U[1][j1] * U[0][j2];
        const float along = cx * B[m][0][j] + cy * B[m][1][j] + cz * B[m][2][j];
        const float sg = along < 0.0f ? -1.0f : 1.0f;
        U[0][j] = sg * cx;
        U[1][j] = sg * cy;
        U[2][j] = sg * cz;
        const float vx = V[1][j1] * V[2][j2] - V[2][j1] * V[1][j2];
        const float vy = V[2][j1] * V[0][j2] - V[0][j1] * V[2][j2];
        const float vz = V[0][j1] * V[1][j2] - V[1][j1] * V[0][j2];
        V[0][j] = vx;
        V[1][j] = vy;
        V[2][j] = vz;
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) R[m][r][cc] = U[r][0] * V[cc][0] + U[r][1] * V[cc][1] + U[r][2] * V[cc][2];
    const float det = R[m][0][0] * (R[m][1][1] * R[m][2][2] - R[m][1][2] * R[m][2][1]) -
                      R[m][0][1] * (R[m][1][0] * R[m][2][2] - R[m][1][2] * R[m][2][0]) +
                      R[m][0][2] * (R[m][1][0] * R[m][2][1] - R[m][1][1] * R[m][2][0]);
    if (det < 0.0f) {  // OpenCV's handling of a reflection: negate the third ROW of R (App. B.3j)
      R[m][2][0] = -R[m][2][0];
      R[m][2][1] = -R[m][2][1];
      R[m][2][2] = -R[m][2][2];
    }
  }
}

// ------------------------------------------------------------------------------------------
// EPnP's four vectors without a full SVD (a one-sided Jacobi SVD of M^T is kept as a development variant,
// dev_variants.cuh).  EPnP needs the four smallest right singular directions of
// the 10 x 12 matrix M only:
//   * v0, v1 span the exact null space.  Householder QR of A = M^T (12 x 10, in place: R in the upper
//     triangle, the reflectors below it) gives it for free: the last two columns of Q;
//   * v2, v3 belong to the two smallest singular values of R (M^T M = Q R R^T Q^T): block inverse
//     iteration x <- R^-T R^-1 x on two vectors (two triangular solves each, Gram-Schmidt every
//     step), a 2 x 2 Rayleigh-Ritz rotation to separate them, then v = Q [w; 0; 0].
// The bottom of M's spectrum is strongly graded for a perspective camera metres away from a
// sub-metre target (sigma_2 / sigma_3 ~ 0.15 median on the benchmark data), so the iteration reaches
// FP32 noise in 4-6 steps; where it has not (close range, sigma_3 ~ sigma_4) v3 is a mixture inside an
// almost degenerate pair, which is as arbitrary in OpenCV's own SVD.  Checked against cv2 with the
// NumPy model tools/proto_eig.py before it was written and on the GPU afterwards: per-hypothesis
// inlier-count agreement and winner-mask agreement are the same as with the Jacobi SVD.
// ~3.7 k instead of ~18 k instructions per hypothesis for this stage.
// Row/column order of A = M^T in this routine (eig_row / the fill in the kernel): columns 0..4 are the
// x-equations of the five points, 5..9 their y-equations; rows 0..7 are the (x, z) components of the
// four control points, rows 8..11 the y components.  An x-equation has no y component, so the first
// five reflectors and the columns they come from live in rows 0..7 only: every inner loop of steps
// 0..4 (and of their later applications) stops at row 8 instead of 12 (kQrRowEnd).
SPE_HD __forceinline__ constexpr int kQrRowEnd(int k) { return k < 5 ? 8 : 12; }
// position in the 12-vector (control point j, component c) <- row of A
SPE_HD __forceinline__ constexpr int eig_row_to_coord(int r) { return r < 8 ? 3 * (r / 2) + ((r & 1) ? 2 : 0) : 3 * (r - 8) + 1; }

// T = float: the FP32 hypothesis kernel (SFU approximations, `work` in shared memory); T = double: the float64
// replay (ransac_exact.cu).
template <typename T>
SPE_HD __forceinline__ void eig_qr_inverse_iteration(T (&A)[12][10], T* __restrict__ work, int iters) {
  using R_ = Real<T>;
  constexpr T kTiny = T(1e-30), kTinier = T(1e-37);
  // ---- Householder QR, H_k = I - tau_k v_k v_k^T with v_k = (1, A[k+1..][k]) --------------------
  // tau_k is parked in work[24 + k] (the v2 slot is free until the very end).
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    T ss = T(0);
#pragma unroll
    for (int i = k + 1; i < kQrRowEnd(k); ++i) ss = fma(A[i][k], A[i][k], ss);
    const T x0 = A[k][k];
    const T nn = fma(x0, x0, ss);
    const bool ok = nn > kTiny;
    const T nrm = R_::sqrt(nn);
    const T v0 = x0 + R_::copysign(nrm, x0);
    const T iv0 = ok ? R_::rcp(v0) : T(0);
    const T tau = ok ? (R_::abs(x0) + nrm) * R_::rcp(nrm) : T(0);
    work[24 + k] = tau;
    A[k][k] = -R_::copysign(nrm, x0);
#pragma unroll
    for (int i = k + 1; i < kQrRowEnd(k); ++i) A[i][k] *= iv0;
#pragma unroll
    for (int j = k + 1; j < 10; ++j) {
      T s = A[k][j];
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) s = fma(A[i][k], A[i][j], s);
      s *= tau;
      A[k][j] -= s;
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) A[i][j] = fma(-s, A[i][k], A[i][j]);
    }
  }
  // y <- Q y = H_0 ( ... (H_9 y))
  auto apply_q = [&](T (&y)[12]) {
#pragma unroll
    for (int k = 9; k >= 0; --k) {
      T s = y[k];
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) s = fma(A[i][k], y[i], s);
      s *= work[24 + k];
      y[k] -= s;
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) y[i] = fma(-s, A[i][k], y[i]);
    }
  };
  // ---- null space: the last two columns of Q ---------------------------------------------------
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    T y[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) y[r] = r == 10 + c ? T(1) : T(0);
    apply_q(y);
#pragma unroll
    for (int r = 0; r < 12; ++r) work[12 * c + eig_row_to_coord(r)] = y[r];
  }
  // ---- block inverse iteration on R R^T ---------------------------------------------------------
  T rinv[10];
  {
    T rmax = T(0);
#pragma unroll
    for (int i = 0; i < 10; ++i) rmax = fmax(rmax, R_::abs(A[i][i]));
    const T floor_ = fmax(rmax * R_::pivot_floor, kTiny);  // a numerically zero pivot: keep the solves finite
#pragma unroll
    for (int i = 0; i < 10; ++i) rinv[i] = R_::rcp(R_::copysign(fmax(R_::abs(A[i][i]), floor_), A[i][i]));
  }
  T w0[10] = {T(1.0), T(-0.7), T(0.5), T(0.9), T(-0.4), T(0.8), T(-0.6), T(0.3), T(-0.95), T(0.65)};
  T w1[10] = {T(0.6), T(0.85), T(-0.45), T(0.35), T(0.75), T(-0.9), T(-0.5), T(0.55), T(0.4), T(-0.8)};
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    // Both substitutions in their column-oriented (axpy) form: once an unknown is final it is eliminated from
    // all the remaining equations with independent FMAs, so the dependent chain is 10 x (mul, fma) instead of
    // the 55 serial FMAs of the row-oriented loops.
    // R a = w (back substitution), in place
#pragma unroll
    for (int j = 9; j >= 0; --j) {
      w0[j] *= rinv[j];
      w1[j] *= rinv[j];
#pragma unroll
      for (int i = 0; i < j; ++i) {
        w0[i] = fma(-A[i][j], w0[j], w0[i]);
        w1[i] = fma(-A[i][j], w1[j], w1[i]);
      }
    }
    // R^T y = a (forward substitution), in place
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      w0[j] *= rinv[j];
      w1[j] *= rinv[j];
#pragma unroll
      for (int i = j + 1; i < 10; ++i) {
        w0[i] = fma(-A[j][i], w0[j], w0[i]);
        w1[i] = fma(-A[j][i], w1[j], w1[i]);
      }
    }
    // Gram-Schmidt
    T n0 = T(0), d01 = T(0);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      n0 = fma(w0[i], w0[i], n0);
      d01 = fma(w0[i], w1[i], d01);
    }
    const T i0 = R_::rsqrt(fmax(n0, kTiny));
    const T proj = d01 * i0 * i0;
    T n1 = T(0);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      w1[i] = fma(-proj, w0[i], w1[i]);
      w0[i] *= i0;
      n1 = fma(w1[i], w1[i], n1);
    }
    const T i1 = R_::rsqrt(fmax(n1, kTiny));
#pragma unroll
    for (int i = 0; i < 10; ++i) w1[i] *= i1;
  }
  // ---- Rayleigh-Ritz inside the pair: S = W^T R R^T W, rotate so that S is diagonal ---------------
  {
    T s00 = T(0), s01 = T(0), s11 = T(0);
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      T g0 = T(0), g1 = T(0);
#pragma unroll
      for (int k = 0; k <= j; ++k) {
        g0 = fma(A[k][j], w0[k], g0);
        g1 = fma(A[k][j], w1[k], g1);
      }
      s00 = fma(g0, g0, s00);
      s01 = fma(g0, g1, s01);
      s11 = fma(g1, g1, s11);
    }
    const T h = s11 - s00, gg = s01 + s01;
    const T q = R_::sqrt(fma(h, h, fma(gg, gg, kTinier)));
    const T t = gg * R_::rcp(h + R_::copysign(q, h));
    const T c = R_::rsqrt(fma(t, t, T(1))), sn = c * t;
    const bool swap = fma(-t, s01, s00) > fma(t, s01, s11);  // v2 = the smaller Ritz value
    T y2[12], y3[12];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const T a = c * w0[i] - sn * w1[i], b = sn * w0[i] + c * w1[i];
      y2[i] = swap ? b : a;
      y3[i] = swap ? a : b;
    }
    y2[10] = y2[11] = y3[10] = y3[11] = T(0);
    // both back-transforms read tau from work[24..33] before v2 is written over it
    apply_q(y2);
    apply_q(y3);
#pragma unroll
    for (int r = 0; r < 12; ++r) work[24 + eig_row_to_coord(r)] = y2[r], work[36 + eig_row_to_coord(r)] = y3[r];
  }
}

// The float64 variant of the stage above for the replay of cv2's loop (ransac_exact.cu): same Householder QR and null
// space, but the two smallest singular directions of R come from a block of FOUR vectors (two guard vectors) and a 4 x 4
// Rayleigh-Ritz step.  With a block of two, v3 converges like (sigma_3 / sigma_4)^2 per step, which is ~0.93 when the
// third and fourth smallest singular values of M nearly coincide (seen on close-range frames: v3 was still a 80-degree
// mixture after 12 steps and the hypothesis lost all its inliers); with two guard vectors the rate is
// (sigma_3 / sigma_6)^2.  The FP32 kernel keeps the two-vector block: it only screens.
template <typename T>
SPE_HD __forceinline__ void eig_qr_subspace4(T (&A)[12][10], T* __restrict__ work, int iters) {
  using R_ = Real<T>;
  constexpr T kTiny = T(1e-30);
#pragma unroll
  for (int k = 0; k < 10; ++k) {  // Householder QR, as in eig_qr_inverse_iteration
    T ss = T(0);
#pragma unroll
    for (int i = k + 1; i < kQrRowEnd(k); ++i) ss = fma(A[i][k], A[i][k], ss);
    const T x0 = A[k][k];
    const T nn = fma(x0, x0, ss);
    const bool ok = nn > kTiny;
    const T nrm = R_::sqrt(nn);
    const T v0 = x0 + R_::copysign(nrm, x0);
    const T iv0 = ok ? R_::rcp(v0) : T(0);
    const T tau = ok ? (R_::abs(x0) + nrm) * R_::rcp(nrm) : T(0);
    work[24 + k] = tau;
    A[k][k] = -R_::copysign(nrm, x0);
#pragma unroll
    for (int i = k + 1; i < kQrRowEnd(k); ++i) A[i][k] *= iv0;
#pragma unroll
    for (int j = k + 1; j < 10; ++j) {
      T s = A[k][j];
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) s = fma(A[i][k], A[i][j], s);
      s *= tau;
      A[k][j] -= s;
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) A[i][j] = fma(-s, A[i][k], A[i][j]);
    }
  }
  auto apply_q = [&](T (&y)[12]) {
#pragma unroll
    for (int k = 9; k >= 0; --k) {
      T s = y[k];
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) s = fma(A[i][k], y[i], s);
      s *= work[24 + k];
      y[k] -= s;
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) y[i] = fma(-s, A[i][k], y[i]);
    }
  };
#pragma unroll
  for (int c = 0; c < 2; ++c) {  // null space: the last two columns of Q
    T y[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) y[r] = r == 10 + c ? T(1) : T(0);
    apply_q(y);
#pragma unroll
    for (int r = 0; r < 12; ++r) work[12 * c + eig_row_to_coord(r)] = y[r];
  }
  T rinv[10];
  {
    T rmax = T(0);
#pragma unroll
    for (int i = 0; i < 10; ++i) rmax = fmax(rmax, R_::abs(A[i][i]));
    const T floor_ = fmax(rmax * R_::pivot_floor, kTiny);
#pragma unroll
    for (int i = 0; i < 10; ++i) rinv[i] = R_::rcp(R_::copysign(fmax(R_::abs(A[i][i]), floor_), A[i][i]));
  }
  T w[4][10] = {{T(1.0), T(-0.7), T(0.5), T(0.9), T(-0.4), T(0.8), T(-0.6), T(0.3), T(-0.95), T(0.65)},
                {T(0.6), T(0.85), T(-0.45), T(0.35), T(0.75), T(-0.9), T(-0.5), T(0.55), T(0.4), T(-0.8)},
                {T(-0.3), T(0.45), T(0.95), T(-0.65), T(0.2), T(0.5), T(0.85), T(-0.75), T(0.6), T(0.1)},
                {T(0.8), T(-0.2), T(-0.6), T(-0.5), T(0.9), T(0.15), T(0.7), T(0.95), T(-0.35), T(0.55)}};
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 9; j >= 0; --j) {  // R a = w
#pragma unroll
      for (int m = 0; m < 4; ++m) w[m][j] *= rinv[j];
#pragma unroll
      for (int i = 0; i < j; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) w[m][i] = fma(-A[i][j], w[m][j], w[m][i]);
    }
#pragma unroll
    for (int j = 0; j < 10; ++j) {  // R^T y = a
#pragma unroll
      for (int m = 0; m < 4; ++m) w[m][j] *= rinv[j];
#pragma unroll
      for (int i = j + 1; i < 10; ++i)
#pragma unroll
        for (int m = 0; m < 4; ++m) w[m][i] = fma(-A[j][i], w[m][j], w[m][i]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {  // modified Gram-Schmidt
#pragma unroll
      for (int p = 0; p < m; ++p) {
        T d = T(0);
#pragma unroll
        for (int i = 0; i < 10; ++i) d = fma(w[p][i], w[m][i], d);
#pragma unroll
        for (int i = 0; i < 10; ++i) w[m][i] = fma(-d, w[p][i], w[m][i]);
      }
      T n2 = T(0);
#pragma unroll
      for (int i = 0; i < 10; ++i) n2 = fma(w[m][i], w[m][i], n2);
      const T inv = R_::rsqrt(fmax(n2, kTiny));
#pragma unroll
      for (int i = 0; i < 10; ++i) w[m][i] *= inv;
    }
  }
  // Rayleigh-Ritz on the block: S = G^T G with G = R^T W (10 x 4); cyclic Jacobi on the 4 x 4 matrix, E = eigenvectors
  T S[4][4], E[4][4];
  {
    T G[4][10];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        T g = T(0);
#pragma unroll
        for (int k = 0; k <= j; ++k) g = fma(A[k][j], w[m][k], g);
        G[m][j] = g;
      }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        T s = T(0);
#pragma unroll
        for (int j = 0; j < 10; ++j) s = fma(G[a][j], G[b][j], s);
        S[a][b] = s;
        E[a][b] = a == b ? T(1) : T(0);
      }
  }
#pragma unroll 1
  for (int sweep = 0; sweep < 6; ++sweep) {
    bool any = false;  // the block has (nearly) converged to eigenvectors: S is close to diagonal, one or two sweeps do
#pragma unroll
    for (int p = 0; p < 3; ++p)
#pragma unroll
      for (int q = p + 1; q < 4; ++q) {
        const T apq = S[p][q];
        if (!(apq * apq > (R_::eps * R_::eps) * R_::abs(S[p][p] * S[q][q]))) continue;
        any = true;
        const T h = S[q][q] - S[p][p], gg = apq + apq;
        const T den = h + R_::copysign(R_::sqrt(fma(h, h, gg * gg)), h);
        const T t = R_::abs(den) > T(0) ? gg / den : T(0);
        const T c = R_::rsqrt(fma(t, t, T(1))), s = c * t;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // S <- S J
          const T x = S[k][p], y = S[k][q];
          S[k][p] = c * x - s * y, S[k][q] = s * x + c * y;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // S <- J^T S,  E <- E J
          const T x = S[p][k], y = S[q][k];
          S[p][k] = c * x - s * y, S[q][k] = s * x + c * y;
          const T ex = E[k][p], ey = E[k][q];
          E[k][p] = c * ex - s * ey, E[k][q] = s * ex + c * ey;
        }
      }
    if (!any) break;
  }
  // the two smallest Ritz values
  int i2 = 0, i3 = 0;
  {
    T b2 = S[0][0];
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (S[k][k] < b2) b2 = S[k][k], i2 = k;
    T b3 = T(0);
    bool have = false;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k != i2 && (!have || S[k][k] < b3)) b3 = S[k][k], i3 = k, have = true;
  }
  T y2[12], y3[12];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    T a = T(0), b = T(0);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      // E[m][k] picked without dynamic register indexing
      T e2 = T(0), e3 = T(0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        e2 = k == i2 ? E[m][k] : e2;
        e3 = k == i3 ? E[m][k] : e3;
      }
      a = fma(e2, w[m][i], a);
      b = fma(e3, w[m][i], b);
    }
    y2[i] = a, y3[i] = b;
  }
  y2[10] = y2[11] = y3[10] = y3[11] = T(0);
  apply_q(y2);
  apply_q(y3);
#pragma unroll
  for (int r = 0; r < 12; ++r) work[24 + eig_row_to_coord(r)] = y2[r], work[36 + eig_row_to_coord(r)] = y3[r];
}

// cv_rotation_matrix_to_quat (pose_estimation/export_predicted_poses_real.py:22-57):
// scalar-first quaternion, branch on the largest of the four candidate magnitudes.
SPE_HD __forceinline__ void rotation_to_quat(const double (&r)[3][3], double (&q)[4]) {
  const double e0 = sqrt(fmax(1.0 + r[0][0] + r[1][1] + r[2][2], 0.0)) * 0.5;
  const double e1 = sqrt(fmax(1.0 + r[0][0] - r[1][1] - r[2][2], 0.0)) * 0.5;
  const double e2 = sqrt(fmax(1.0 - r[0][0] + r[1][1] - r[2][2], 0.0)) * 0.5;
  const double e3 = sqrt(fmax(1.0 - r[0][0] - r[1][1] + r[2][2], 0.0)) * 0.5;
  int m = 0;  // np.argmax: first maximum
  double best = e0;
  if (e1 > best) best = e1, m = 1;
  if (e2 > best) best = e2, m = 2;
  if (e3 > best) best = e3, m = 3;
  if (m == 0) {
    q[0] = e0;
    q[1] = (r[2][1] - r[1][2]) / (4 * e0);
    q[2] = (r[0][2] - r[2][0]) / (4 * e0);
    q[3] = (r[1][0] - r[0][1]) / (4 * e0);
  } else if (m == 1) {
    q[0] = (r[2][1] - r[1][2]) / (4 * e1);
    q[1] = e1;
    q[2] = (r[1][0] + r[0][1]) / (4 * e1);
    q[3] = (r[2][0] + r[0][2]) / (4 * e1);
  } else if (m == 2) {
    q[0] = (r[0][2] - r[2][0]) / (4 * e2);
    q[1] = (r[1][0] + r[0][1]) / (4 * e2);
    q[2] = e2;
    q[3] = (r[2][1] + r[1][2]) / (4 * e2);
  } else {
    q[0] = (r[1][0] - r[0][1]) / (4 * e3);
    q[1] = (r[2][0] + r[0][2]) / (4 * e3);
    q[2] = (r[2][1] + r[1][2]) / (4 * e3);
    q[3] = e3;
  }
}

}  // namespace spe
