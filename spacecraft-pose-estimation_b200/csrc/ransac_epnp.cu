// Batched RANSAC-EPnP for sm_100a.
//
// Replaces the per-frame loop of pose_estimation/export_predicted_poses_real.py:177-204
// (confidence filter, cv2.solvePnPRansac(flags=SOLVEPNP_EPNP, iterationsCount, reprojectionError),
// cv2.Rodrigues, cv_rotation_matrix_to_quat).  OpenCV calib3d is an un-vendored dependency of the
// reference; its algorithm is restated in SURVEY.md App. B and oracle/{ocv_rng,epnp_ref,pnp_ref}.py.
//
// Three kernels per call, all on one stream, no host synchronisation:
//   1. frame_prep_kernel   warp per frame: confidence filter -> visible set, 5-iteration
//                          undistortion of every landmark (float64, cv2.undistortPoints)
//   2. hypothesis_kernel   FP32.  Every minimal set of OpenCV's fixed-seed RNG is scored; there is
//                          no early exit on the GPU.  A group of 4 lanes owns one (frame, hypothesis):
//                          the 10x12 matrix M and the 12x12 accumulator V are distributed by rows
//                          over the 4 lanes (72 registers each) and orthogonalised by one-sided
//                          Jacobi rotations — the Jacobi eigensolve of MtM applied implicitly, without
//                          squaring M's condition number in FP32; dot products cross lanes with
//                          two shuffles.  The three EPnP beta initialisations + Gauss-Newton +
//                          Procrustes then run one variant per lane, and the inlier scoring is
//                          split over the lanes again (shuffle OR of the inlier bits).
//   3. select_refit_kernel replays cv2's sequential acceptance rule (first strictly better count,
//                          adaptively shrinking iteration budget) over the H inlier counts, then
//                          runs the final EPnP on the winner's inliers in float64, following
//                          OpenCV operation by operation where signs depend on it (PCA axes).
//
// Compute-bound on FP32 CUDA cores (no tensor cores: nothing here is a dense contraction);
// HBM traffic is ~100 B per hypothesis.
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"
#include "device_util.cuh"
#include "epnp_math.cuh"
#include "ransac.cuh"

namespace spe {

// ------------------------------------------------------------------------------------------
// OpenCV RNG (multiply-with-carry, seeded with (uint64)-1 on every RANSAC run)
void opencv_minimal_sets(int count, int num, uint8_t* out) {
  uint64_t state = ~0ull;
  auto next = [&]() -> uint32_t {
    state = (uint64_t)(uint32_t)state * 4164903690ull + (state >> 32);
    return (uint32_t)state;
  };
  for (int h = 0; h < num; ++h) {
    uint8_t* s = out + (size_t)h * kModelPoints;
    for (int i = 0; i < kModelPoints; ++i) {
      for (;;) {
        const uint32_t v = next() % (uint32_t)count;
        bool dup = false;
        for (int k = 0; k < i; ++k) dup |= (s[k] == v);
        if (!dup) {
          s[i] = (uint8_t)v;
          break;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Control points of every 5-subset (host, float64, once per model).  EPnP's control points are the
// centroid and the PCA axes of the object points scaled by sqrt(lambda_i / 5) (App. B.3c), the
// barycentric coordinates alpha_ki = (p_k - c0) . v_i / k_i; both depend on the 5 landmarks only, so
// the hypothesis kernel looks them up instead of running a 5x3 Jacobi per hypothesis.  Entry order =
// combinatorial number system: rank(j0<j1<j2<j3<j4) = C(j0,1)+C(j1,2)+C(j2,3)+C(j3,4)+C(j4,5).
static void sym3_jacobi(double S[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
    if (off <= 1e-300 || off <= 1e-18 * (fabs(S[0][0]) + fabs(S[1][1]) + fabs(S[2][2]))) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (S[p][q] == 0.0) continue;
        const double zeta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int k = 0; k < 3; ++k) {  // S <- S J
          const double x = S[k][p], y = S[k][q];
          S[k][p] = c * x - sn * y, S[k][q] = sn * x + c * y;
        }
        for (int k = 0; k < 3; ++k) {  // S <- J^T S
          const double x = S[p][k], y = S[q][k];
          S[p][k] = c * x - sn * y, S[q][k] = sn * x + c * y;
        }
        for (int k = 0; k < 3; ++k) {
          const double x = V[k][p], y = V[k][q];
          V[k][p] = c * x - sn * y, V[k][q] = sn * x + c * y;
        }
      }
  }
}

// one table entry from the float32 landmark coordinates of the five points (ascending landmark order)
void control_table_entry(const float* landmarks_f32, const int (&ids)[5], float* e) {
  double P[5][3], c0[3] = {0, 0, 0};
  for (int k = 0; k < 5; ++k)
    for (int c = 0; c < 3; ++c) {
      P[k][c] = (double)landmarks_f32[3 * ids[k] + c];
      c0[c] += P[k][c] * 0.2;
    }
  double S[3][3] = {}, V[3][3];
  for (int k = 0; k < 5; ++k)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) S[r][c] += (P[k][r] - c0[r]) * (P[k][c] - c0[c]);
  sym3_jacobi(S, V);
  for (int i = 0; i < kCtrlEntryFloats; ++i) e[i] = 0.f;
  for (int i = 0; i < 3; ++i) {
    const double lam = S[i][i] > 0.0 ? S[i][i] : 0.0;
    const double ki = sqrt(lam * 0.2);
    const double inv = ki > 1e-12 ? 1.0 / ki : 0.0;
    for (int k = 0; k < 5; ++k) {
      double proj = 0.0;
      for (int c = 0; c < 3; ++c) proj += (P[k][c] - c0[c]) * V[c][i];
      e[3 * k + i] = (float)(proj * inv);
    }
    e[15 + i] = (float)(ki * ki);
  }
}

// rank of a sorted 5-subset in the table (the device computes the same expression, hypothesis_kernel_t1)
size_t control_table_rank(const int (&j)[5]) {
  auto c = [](size_t n, int k) {
    size_t r = 1;
    for (int i = 0; i < k; ++i) r = r * (n - i) / (i + 1);
    return n >= (size_t)k ? r : 0;
  };
  return c(j[0], 1) + c(j[1], 2) + c(j[2], 3) + c(j[3], 4) + c(j[4], 5);
}

static void build_control_table(const Model& m, std::vector<float>& table) {
  const int J = m.J;
  size_t count = 0;
  for (int a = 4; a < J; ++a) count += (size_t)a * (a - 1) * (a - 2) * (a - 3) / 24;  // C(J,5) = sum C(a,4)
  table.assign(count * kCtrlEntryFloats, 0.f);
  size_t rank = 0;  // j4 outermost ... j0 innermost enumerates the ranks in increasing order
  for (int j4 = 4; j4 < J; ++j4)
    for (int j3 = 3; j3 < j4; ++j3)
      for (int j2 = 2; j2 < j3; ++j2)
        for (int j1 = 1; j1 < j2; ++j1)
          for (int j0 = 0; j0 < j1; ++j0, ++rank) {
            const int ids[5] = {j0, j1, j2, j3, j4};
            control_table_entry(m.landmarks_f32, ids, table.data() + rank * kCtrlEntryFloats);
          }
}

cudaError_t model_upload(Model& m) {
  cudaError_t e = cudaGetDevice(&m.device);
  if (e != cudaSuccess) return e;
  e = cudaMalloc(&m.d_landmarks, sizeof(float) * 3 * m.J);
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(m.d_landmarks, m.landmarks_f32, sizeof(float) * 3 * m.J, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  const int tables = m.J >= 6 ? m.J - 5 : 0;
  m.h_subsets.assign((size_t)tables * m.max_hyp * kModelPoints, 0);
  for (int n = 6; n <= m.J; ++n) opencv_minimal_sets(n, m.max_hyp, m.h_subsets.data() + (size_t)(n - 6) * m.max_hyp * kModelPoints);
  if (tables > 0) {
    e = cudaMalloc(&m.d_subsets, m.h_subsets.size());
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(m.d_subsets, m.h_subsets.data(), m.h_subsets.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    std::vector<float> ctrl;
    build_control_table(m, ctrl);
    m.ctrl_entries = ctrl.size() / kCtrlEntryFloats;
    e = cudaMalloc(&m.d_ctrl, ctrl.size() * sizeof(float));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(m.d_ctrl, ctrl.data(), ctrl.size() * sizeof(float), cudaMemcpyHostToDevice);
  }
  return e;
}

void model_free(Model& m) {
  if (m.d_landmarks) cudaFree(m.d_landmarks);
  if (m.d_subsets) cudaFree(m.d_subsets);
  if (m.d_ctrl) cudaFree(m.d_ctrl);
  m.d_landmarks = nullptr;
  m.d_subsets = nullptr;
  m.d_ctrl = nullptr;
}

// ------------------------------------------------------------------------------------------
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

RansacWorkspace carve_workspace(void* base, int J, int B, int H) {
  RansacWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<unsigned char*>(base) + off : nullptr;
    off += align16(bytes);
    return p;
  };
  w.und = static_cast<double2*>(take(sizeof(double2) * (size_t)B * J));
  w.us_hyp = static_cast<float2*>(take(sizeof(float2) * (size_t)B * J));
  w.img = static_cast<float2*>(take(sizeof(float2) * (size_t)B * J));
  w.n = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
  w.vis = static_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)B));
  w.masks = static_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)B * H));
  w.counts = static_cast<uint8_t*>(take((size_t)B * H));
  w.need = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
  w.frames = B;
  w.bytes = off;
  return w;
}

size_t ransac_workspace_bytes(int J, int B, int H) { return carve_workspace(nullptr, J, B, H).bytes; }

namespace {

constexpr unsigned kFull = 0xffffffffu;

struct DevModel {
  const float* landmarks;  // [J,3]
  const uint8_t* subsets;  // [J-5][max_hyp][5]
  int J, max_hyp;
  Camera cam;
  const float4* ctrl;  // [C(J,5)][kCtrlEntryFloats / 4] control-point table (thread-per-hypothesis kernel)
};

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

// ------------------------------------------------------------------------------------------
// 2. hypothesis kernel
constexpr int kGroup = 4;                   // lanes per hypothesis
constexpr int kHypPerCta = 32;              // 128 threads
constexpr int kVStride = 4 * 12 + 20 + 1;   // v[4][12] + alphas[5][4] + pad (odd: conflict-free across groups)

// round-robin (circle) schedule of the 66 column pairs of a sweep: 11 rounds x 6 disjoint pairs
__host__ __device__ constexpr int rr_p(int r, int k) { return k == 0 ? r : (r + k) % 11; }
__host__ __device__ constexpr int rr_q(int r, int k) { return k == 0 ? 11 : (r - k + 11) % 11; }

__device__ __forceinline__ float group_sum(float x) {
  x += __shfl_xor_sync(kFull, x, 1);
  x += __shfl_xor_sync(kFull, x, 2);
  return x;
}

// One-sided Jacobi on the columns of W = [M (rows spread over the 4 lanes); V].  Mr/Vr hold this
// lane's 3 rows of each.  Equivalent to the Jacobi eigensolve of MtM with V accumulating the
// eigenvectors; after convergence the column norms are the singular values of M.
__device__ __forceinline__ void jacobi_sweeps(float (&Mr)[3][12], float (&Vr)[3][12], float (&d)[12], int sweeps) {
  constexpr float kTol2 = 9e-14f;  // (3e-7)^2: skip pairs that are orthogonal to FP32 accuracy
#pragma unroll 1
  for (int sw = 0; sw < sweeps; ++sw) {
#pragma unroll
    for (int j = 0; j < 12; ++j) d[j] = group_sum(Mr[0][j] * Mr[0][j] + Mr[1][j] * Mr[1][j] + Mr[2][j] * Mr[2][j]);
#pragma unroll
    for (int r = 0; r < 11; ++r) {
      float g[6], c[6], s[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int p = rr_p(r, k), q = rr_q(r, k);
        g[k] = Mr[0][p] * Mr[0][q] + Mr[1][p] * Mr[1][q] + Mr[2][p] * Mr[2][q];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) g[k] += __shfl_xor_sync(kFull, g[k], 1);
#pragma unroll
      for (int k = 0; k < 6; ++k) g[k] += __shfl_xor_sync(kFull, g[k], 2);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int p = rr_p(r, k), q = rr_q(r, k);
        const bool rot = g[k] * g[k] > kTol2 * d[p] * d[q];
        float t;
        jacobi_angle_fast(d[p], d[q], rot ? g[k] : 1.0f, c[k], s[k], t);
        c[k] = rot ? c[k] : 1.0f;
        s[k] = rot ? s[k] : 0.0f;
        t = rot ? t : 0.0f;
        d[p] = fmaxf(d[p] - t * g[k], 0.0f);
        d[q] = fmaxf(d[q] + t * g[k], 0.0f);
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int p = rr_p(r, k), q = rr_q(r, k);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float x = Mr[i][p], y = Mr[i][q];
          Mr[i][p] = c[k] * x - s[k] * y;
          Mr[i][q] = s[k] * x + c[k] * y;
          const float vx = Vr[i][p], vy = Vr[i][q];
          Vr[i][p] = c[k] * vx - s[k] * vy;
          Vr[i][q] = s[k] * vx + c[k] * vy;
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) d[j] = group_sum(Mr[0][j] * Mr[0][j] + Mr[1][j] * Mr[1][j] + Mr[2][j] * Mr[2][j]);
}

// Control points and barycentric coordinates of a 5-point set (App. B.3c-d) without forming the
// covariance: one-sided Jacobi on the centred 5x3 point matrix gives the PCA axes (columns of V)
// and P0 V, whose column norms are sqrt(lambda).  Axis order/sign is free for a hypothesis.
__device__ __forceinline__ void control_points5(const float (&pw)[5][3], float (&cws)[4][3], float (&al)[5][4]) {
  float c0[3], Bm[5][3], V[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) c0[c] = (pw[0][c] + pw[1][c] + pw[2][c] + pw[3][c] + pw[4][c]) * 0.2f;
#pragma unroll
  for (int k = 0; k < 5; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) Bm[k][c] = pw[k][c] - c0[c];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.f : 0.f;
#pragma unroll 1
  for (int sweep = 0; sweep < Real<float>::svd3_sweeps; ++sweep) {
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      float a = 0.f, b = 0.f, g = 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        a += Bm[k][p] * Bm[k][p];
        b += Bm[k][q] * Bm[k][q];
        g += Bm[k][p] * Bm[k][q];
      }
      const bool rot = g * g > 1.4e-14f * a * b;
      float c, s, t;
      jacobi_angle_fast(a, b, rot ? g : 1.f, c, s, t);
      c = rot ? c : 1.f;
      s = rot ? s : 0.f;
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float x = Bm[k][p], y = Bm[k][q];
        Bm[k][p] = c * x - s * y;
        Bm[k][q] = s * x + c * y;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float x = V[k][p], y = V[k][q];
        V[k][p] = c * x - s * y;
        V[k][q] = s * x + c * y;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) cws[0][c] = c0[c];
  float inv_k[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) s2 += Bm[k][i] * Bm[k][i];
    const float ki = sqrt_approx(s2 * 0.2f);  // sqrt(lambda_i / 5)
    inv_k[i] = ki > 1e-12f ? rcp_approx(ki) : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) cws[i + 1][c] = c0[c] + ki * V[c][i];
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    al[k][1] = Bm[k][0] * inv_k[0];
    al[k][2] = Bm[k][1] * inv_k[1];
    al[k][3] = Bm[k][2] * inv_k[2];
    al[k][0] = 1.0f - al[k][1] - al[k][2] - al[k][3];
  }
}

__global__ void __launch_bounds__(kHypPerCta * kGroup, 3)
hypothesis_kernel(DevModel m, const float* __restrict__ kpts, int H, int hblocks, float thr2, int sweeps, RansacWorkspace ws) {
  __shared__ float s_pw[kMaxLandmarks][3];
  __shared__ float2 s_us[kMaxLandmarks];
  __shared__ float2 s_img[kMaxLandmarks];
  __shared__ float s_work[kHypPerCta][kVStride];

  const int b = blockIdx.x / hblocks, hb = blockIdx.x - b * hblocks;
  const int n = ws.n[b];
  if (n <= kModelPoints) return;  // n < 6: no RANSAC (handled by the refit kernel); uniform per CTA
  const unsigned vis = ws.vis[b];
  const int tid = threadIdx.x;
  if (tid < n) {  // compact the visible landmarks: position k <- k-th set bit of vis
    const int j = __fns(vis, 0, tid + 1);
    s_pw[tid][0] = m.landmarks[3 * j], s_pw[tid][1] = m.landmarks[3 * j + 1], s_pw[tid][2] = m.landmarks[3 * j + 2];
    s_us[tid] = ws.us_hyp[(size_t)b * m.J + j];
    const float* k = kpts + ((size_t)b * m.J + j) * 3;
    s_img[tid] = make_float2(k[0], k[1]);
  }
  __syncthreads();

  const int grp = tid >> 2, l = tid & 3;
  const int h = hb * kHypPerCta + grp;
  const bool live = h < H;
  const uint8_t* sub = m.subsets + ((size_t)(n - 6) * m.max_hyp + (live ? h : 0)) * kModelPoints;
  float* work = s_work[grp];
  const float fu = (float)m.cam.fx, fv = (float)m.cam.fy, uc = (float)m.cam.cx, vc = (float)m.cam.cy;

  // ---- control points, alphas, this lane's rows of M ---------------------------------------
  float rho[6];
  float Mr[3][12], Vr[3][12], d[12];
  {
    float pw[5][3], al[5][4], us[5][2], cws[4][3];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int si = sub[k];
      pw[k][0] = s_pw[si][0], pw[k][1] = s_pw[si][1], pw[k][2] = s_pw[si][2];
      us[k][0] = s_us[si].x, us[k][1] = s_us[si].y;
    }
    control_points5(pw, cws, al);
    build_rho<float>(cws, rho);
    if (l == 0) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) work[48 + 4 * k + j] = al[k][j];
    }
    // rows l, l+4, l+8 of M: row r belongs to point r>>1, odd rows are the v-equations (App. B.3e)
    const bool isv = l & 1, hi = l >> 1;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const bool valid = (s < 2) || (l < 2);
      const int p0 = 2 * s, p1 = (2 * s + 1 < 5) ? 2 * s + 1 : 4;
      const float uu = hi ? us[p1][0] : us[p0][0], vv = hi ? us[p1][1] : us[p0][1];
      const float w = isv ? (vc - vv) : (uc - uu);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a = hi ? al[p1][j] : al[p0][j];
        a = valid ? a : 0.f;
        Mr[s][3 * j] = isv ? 0.f : a * fu;
        Mr[s][3 * j + 1] = isv ? a * fv : 0.f;
        Mr[s][3 * j + 2] = a * w;
      }
#pragma unroll
      for (int j = 0; j < 12; ++j) Vr[s][j] = (3 * l + s == j) ? 1.f : 0.f;
    }
  }

  // ---- implicit Jacobi eigensolve of MtM -----------------------------------------------------
  jacobi_sweeps(Mr, Vr, d, sweeps);

  // the four smallest singular directions, ascending: v0 = smallest (OpenCV's ut[11]) ... v3
  {
    float dd[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) dd[j] = d[j];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float best = dd[0];
      int jb = 0;
#pragma unroll
      for (int j = 1; j < 12; ++j) {
        const bool lt = dd[j] < best;
        best = lt ? dd[j] : best;
        jb = lt ? j : jb;
      }
      float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
      for (int j = 0; j < 12; ++j) {
        const bool sel = j == jb;
        v0 = sel ? Vr[0][j] : v0;
        v1 = sel ? Vr[1][j] : v1;
        v2 = sel ? Vr[2][j] : v2;
        dd[j] = sel ? INFINITY : dd[j];
      }
      work[12 * i + 3 * l + 0] = v0;
      work[12 * i + 3 * l + 1] = v1;
      work[12 * i + 3 * l + 2] = v2;
    }
  }
  __syncwarp();

  // ---- betas: one EPnP variant per lane (lane 3 repeats variant 1) ----------------------------
  const int variant = l < 3 ? l + 1 : 1;
  float betas[4];
  {
    float v[4][12], L[6][10];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 12; ++j) v[i][j] = work[12 * i + j];
    build_L<float>(v, L);
    approx_betas<float>(L, rho, variant, betas);
    gauss_newton<float>(L, rho, betas);
  }

  // ---- camera-frame control points -> Procrustes -> reprojection error on the 5 points --------
  float R[3][3], t[3], err;
  {
    float ccs[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        ccs[j][c] = betas[0] * work[3 * j + c] + betas[1] * work[12 + 3 * j + c] + betas[2] * work[24 + 3 * j + c] +
                    betas[3] * work[36 + 3 * j + c];
    float pcs[5][3], pw[5][3];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int si = sub[k];
      pw[k][0] = s_pw[si][0], pw[k][1] = s_pw[si][1], pw[k][2] = s_pw[si][2];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        pcs[k][c] = work[48 + 4 * k] * ccs[0][c] + work[48 + 4 * k + 1] * ccs[1][c] + work[48 + 4 * k + 2] * ccs[2][c] +
                    work[48 + 4 * k + 3] * ccs[3][c];
    }
    const float sgn = pcs[0][2] < 0.f ? -1.f : 1.f;  // solve_for_sign
    float pc0[3] = {0.f, 0.f, 0.f}, pw0[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        pcs[k][c] *= sgn;
        pc0[c] += pcs[k][c];
        pw0[c] += pw[k][c];
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) pc0[c] *= 0.2f, pw0[c] *= 0.2f;
    float abt[3][3] = {};
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) abt[r][c] += (pcs[k][r] - pc0[r]) * (pw[k][c] - pw0[c]);
    procrustes_uvt<float>(abt, R);
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = pc0[r] - (R[r][0] * pw0[0] + R[r][1] * pw0[1] + R[r][2] * pw0[2]);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int si = sub[k];
      const float Xc = R[0][0] * pw[k][0] + R[0][1] * pw[k][1] + R[0][2] * pw[k][2] + t[0];
      const float Yc = R[1][0] * pw[k][0] + R[1][1] * pw[k][1] + R[1][2] * pw[k][2] + t[1];
      const float iz = rcp_approx(R[2][0] * pw[k][0] + R[2][1] * pw[k][1] + R[2][2] * pw[k][2] + t[2]);
      const float du = s_us[si].x - (uc + fu * Xc * iz), dv = s_us[si].y - (vc + fv * Yc * iz);
      sum += sqrt_approx(du * du + dv * dv);
    }
    err = sum * 0.2f;
  }

  // ---- best of the three variants (App. B.3k), broadcast to the group -------------------------
  {
    const int base = (threadIdx.x & 31) & ~3;
    const float e1 = __shfl_sync(kFull, err, base), e2 = __shfl_sync(kFull, err, base + 1), e3 = __shfl_sync(kFull, err, base + 2);
    int N = 0;
    if (e2 < e1) N = 1;
    if (e3 < (N == 1 ? e2 : e1)) N = 2;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) R[r][c] = __shfl_sync(kFull, R[r][c], base + N);
      t[r] = __shfl_sync(kFull, t[r], base + N);
    }
  }

  // ---- score all n points: cv2.projectPoints + squared error <= reproj^2 (App. B.5) -----------
  unsigned bits = 0;
  {
    const float k1 = (float)m.cam.k1, k2 = (float)m.cam.k2, p1 = (float)m.cam.p1, p2 = (float)m.cam.p2, k3 = (float)m.cam.k3;
    for (int k = l; k < n; k += kGroup) {
      const float X = s_pw[k][0], Y = s_pw[k][1], Z = s_pw[k][2];
      const float xc = R[0][0] * X + R[0][1] * Y + R[0][2] * Z + t[0];
      const float yc = R[1][0] * X + R[1][1] * Y + R[1][2] * Z + t[1];
      const float zc = R[2][0] * X + R[2][1] * Y + R[2][2] * Z + t[2];
      const float iz = rcp_approx(zc);
      const float x = xc * iz, y = yc * iz;
      const float r2 = x * x + y * y;
      const float cd = 1.0f + ((k3 * r2 + k2) * r2 + k1) * r2;
      const float xd = x * cd + 2.0f * p1 * x * y + p2 * (r2 + 2.0f * x * x);
      const float yd = y * cd + p1 * (r2 + 2.0f * y * y) + 2.0f * p2 * x * y;
      const float du = s_img[k].x - (fu * xd + uc), dv = s_img[k].y - (fv * yd + vc);
      const float e = du * du + dv * dv;
      if (e <= thr2) bits |= 1u << __fns(vis, 0, k + 1);  // back to landmark numbering
    }
    bits |= __shfl_xor_sync(kFull, bits, 1);
    bits |= __shfl_xor_sync(kFull, bits, 2);
  }
  if (live && l == 0) {
    ws.masks[(size_t)b * H + h] = bits;
    ws.counts[(size_t)b * H + h] = (uint8_t)__popc(bits);
  }
}

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
#ifndef SPE_REFIT_REGS
#define SPE_REFIT_REGS 255
#endif
__device__ __forceinline__ int ws_frames(const RansacWorkspace& ws) { return ws.frames; }
constexpr int kT1Stride = 4 * 12 + 20 + 1;  // per-thread scratch: v[4][12], alphas[5][4] (+1: odd stride, conflict-free)
// shared memory of one warp: landmarks [32][3], ideal + raw pixels [32] float2 each, per-thread scratch
constexpr int kT1WarpBytes = (int)(sizeof(float) * 3 * kMaxLandmarks + 2 * sizeof(float2) * kMaxLandmarks + sizeof(float) * 32 * kT1Stride +
                                   kMaxLandmarks /* landmark id of every compacted point */);
constexpr int kT1MaxWarps = 12;  // 12 x 168 registers x 32 = one SM's register file
static_assert(kT1WarpBytes % 16 == 0, "per-warp shared-memory slice must stay 16-byte aligned");

// One Jacobi rotation between the columns at register positions P and Q of A, with the two
// columns SWAPPED on output.  With the swap built in, the odd-even ordering below brings every
// pair of columns together exactly once per sweep while the pairs always sit at the same register
// positions, so a sweep is a short loop (no 45-pair unrolled body that overflows the instruction
// cache, no register moves).
//
// Columns are stored scaled ("fast Givens"): true column j = w[j] * A[:, j].  A rotation then
// costs two FMAs per row instead of four multiply-adds,
//     new Q = c wP (x - t wQ/wP y),   new P = c wQ (y + t wP/wQ x),
// with the factors c wP, c wQ absorbed into w.  d[] holds the TRUE squared norms.
template <int P, int Q>
__device__ __forceinline__ void rotate_swap(float (&A)[12][10], float (&w)[10], float (&d)[10], float g_scaled) {
  const float wp = w[P], wq = w[Q];
  const float g = g_scaled * wp * wq;
  // t = tan(theta) = 2g / (h + sign(h) sqrt(h^2 + 4 g^2)), h = dQ - dP; c = rsqrt(1 + t^2): 3 MUFU.
  // No "already orthogonal" test: a negligible g gives a negligible t (the 1e-30 keeps 0/0 away).
  const float h = d[Q] - d[P], gg = g + g;
  const float q = sqrt_approx(fmaf(h, h, fmaf(gg, gg, 1e-30f)));
  const float t = gg * rcp_approx(h + copysignf(q, h));
  const float c = rsqrt_approx(fmaf(t, t, 1.0f));
  const float dp = fmaf(-t, g, d[P]), dq = fmaf(t, g, d[Q]);
  d[P] = dq;
  d[Q] = dp;
  const float tau1 = t * wq * rcp_approx(wp), tau2 = t * wp * rcp_approx(wq);
  w[Q] = c * wp;
  w[P] = c * wq;
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    const float x = A[i][P], y = A[i][Q];
    A[i][Q] = fmaf(-tau1, y, x);
    A[i][P] = fmaf(tau2, x, y);
  }
}

template <int P, int Q>
__device__ __forceinline__ float col_dot(const float (&A)[12][10]) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 12; i += 3) {
    s0 = fmaf(A[i][P], A[i][Q], s0);
    s1 = fmaf(A[i + 1][P], A[i + 1][Q], s1);
    s2 = fmaf(A[i + 2][P], A[i + 2][Q], s2);
  }
  return s0 + s1 + s2;
}

// exact TRUE squared norms of the scaled columns (once per sweep; the update formula drifts)
__device__ __forceinline__ void true_norms(const float (&A)[12][10], const float (&w)[10], float (&d)[10]) {
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r = 0; r < 12; r += 2) {
      s0 = fmaf(A[r][j], A[r][j], s0);
      s1 = fmaf(A[r + 1][j], A[r + 1][j], s1);
    }
    d[j] = (s0 + s1) * (w[j] * w[j]);
  }
}

// fold the scale factors back into the columns and recompute the exact squared norms
__device__ __forceinline__ void fold_and_norms(float (&A)[12][10], float (&w)[10], float (&d)[10]) {
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int r = 0; r < 12; r += 2) {
      A[r][j] *= w[j];
      A[r + 1][j] *= w[j];
      s0 = fmaf(A[r][j], A[r][j], s0);
      s1 = fmaf(A[r + 1][j], A[r + 1][j], s1);
    }
    w[j] = 1.0f;
    d[j] = s0 + s1;
  }
}

// One-sided Jacobi on the 10 columns of A = M^T in odd-even (transposition) order: a sweep is
// 5 x { pairs (0,1)(2,3)(4,5)(6,7)(8,9) ; pairs (1,2)(3,4)(5,6)(7,8) } = 45 rotations.
__device__ __forceinline__ void jacobi_mt(float (&A)[12][10], float (&d)[10], int sweeps) {
  float w[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) w[j] = 1.0f;
#pragma unroll 1
  for (int it = 0; it < sweeps * 5; ++it) {
    // once per sweep; in between the norms follow the update formula.  The scale factors only
    // shrink by c >= 0.707 per rotation (54 rotations per column in 6 sweeps), far from underflow.
    if (it % 5 == 0) true_norms(A, w, d);
    {
      const float g0 = col_dot<0, 1>(A), g1 = col_dot<2, 3>(A), g2 = col_dot<4, 5>(A), g3 = col_dot<6, 7>(A), g4 = col_dot<8, 9>(A);
      rotate_swap<0, 1>(A, w, d, g0);
      rotate_swap<2, 3>(A, w, d, g1);
      rotate_swap<4, 5>(A, w, d, g2);
      rotate_swap<6, 7>(A, w, d, g3);
      rotate_swap<8, 9>(A, w, d, g4);
    }
    {
      const float g0 = col_dot<1, 2>(A), g1 = col_dot<3, 4>(A), g2 = col_dot<5, 6>(A), g3 = col_dot<7, 8>(A);
      rotate_swap<1, 2>(A, w, d, g0);
      rotate_swap<3, 4>(A, w, d, g1);
      rotate_swap<5, 6>(A, w, d, g2);
      rotate_swap<7, 8>(A, w, d, g3);
    }
  }
  fold_and_norms(A, w, d);
}


// ------------------------------------------------------------------------------------------
// The same four vectors without a full SVD (default; the Jacobi above stays selectable with
// SPE_HYP_KERNEL=jacobi for A/B runs).  EPnP needs the four smallest right singular directions of
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
__device__ __forceinline__ constexpr int kQrRowEnd(int k) { return k < 5 ? 8 : 12; }
// position in the 12-vector (control point j, component c) <- row of A
__device__ __forceinline__ constexpr int eig_row_to_coord(int r) { return r < 8 ? 3 * (r / 2) + ((r & 1) ? 2 : 0) : 3 * (r - 8) + 1; }

__device__ __forceinline__ void eig_qr_inverse_iteration(float (&A)[12][10], float* __restrict__ work, int iters) {
  // ---- Householder QR, H_k = I - tau_k v_k v_k^T with v_k = (1, A[k+1..][k]) --------------------
  // tau_k is parked in work[24 + k] (the v2 slot is free until the very end).
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    float ss = 0.f;
#pragma unroll
    for (int i = k + 1; i < kQrRowEnd(k); ++i) ss = fmaf(A[i][k], A[i][k], ss);
    const float x0 = A[k][k];
    const float nn = fmaf(x0, x0, ss);
    const bool ok = nn > 1e-30f;
    const float nrm = sqrt_approx(nn);
    const float v0 = x0 + copysignf(nrm, x0);
    const float iv0 = ok ? rcp_approx(v0) : 0.f;
    const float tau = ok ? (fabsf(x0) + nrm) * rcp_approx(nrm) : 0.f;
    work[24 + k] = tau;
    A[k][k] = -copysignf(nrm, x0);
#pragma unroll
    for (int i = k + 1; i < kQrRowEnd(k); ++i) A[i][k] *= iv0;
#pragma unroll
    for (int j = k + 1; j < 10; ++j) {
      float s = A[k][j];
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) s = fmaf(A[i][k], A[i][j], s);
      s *= tau;
      A[k][j] -= s;
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) A[i][j] = fmaf(-s, A[i][k], A[i][j]);
    }
  }
  // y <- Q y = H_0 ( ... (H_9 y))
  auto apply_q = [&](float (&y)[12]) {
#pragma unroll
    for (int k = 9; k >= 0; --k) {
      float s = y[k];
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) s = fmaf(A[i][k], y[i], s);
      s *= work[24 + k];
      y[k] -= s;
#pragma unroll
      for (int i = k + 1; i < kQrRowEnd(k); ++i) y[i] = fmaf(-s, A[i][k], y[i]);
    }
  };
  // ---- null space: the last two columns of Q ---------------------------------------------------
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float y[12];
#pragma unroll
    for (int r = 0; r < 12; ++r) y[r] = r == 10 + c ? 1.0f : 0.0f;
    apply_q(y);
#pragma unroll
    for (int r = 0; r < 12; ++r) work[12 * c + eig_row_to_coord(r)] = y[r];
  }
  // ---- block inverse iteration on R R^T ---------------------------------------------------------
  float rinv[10];
  {
    float rmax = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) rmax = fmaxf(rmax, fabsf(A[i][i]));
    const float floor_ = fmaxf(rmax * 1e-7f, 1e-30f);  // a numerically zero pivot: keep the solves finite
#pragma unroll
    for (int i = 0; i < 10; ++i) rinv[i] = rcp_approx(copysignf(fmaxf(fabsf(A[i][i]), floor_), A[i][i]));
  }
  float w0[10] = {1.0f, -0.7f, 0.5f, 0.9f, -0.4f, 0.8f, -0.6f, 0.3f, -0.95f, 0.65f};
  float w1[10] = {0.6f, 0.85f, -0.45f, 0.35f, 0.75f, -0.9f, -0.5f, 0.55f, 0.4f, -0.8f};
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
        w0[i] = fmaf(-A[i][j], w0[j], w0[i]);
        w1[i] = fmaf(-A[i][j], w1[j], w1[i]);
      }
    }
    // R^T y = a (forward substitution), in place
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      w0[j] *= rinv[j];
      w1[j] *= rinv[j];
#pragma unroll
      for (int i = j + 1; i < 10; ++i) {
        w0[i] = fmaf(-A[j][i], w0[j], w0[i]);
        w1[i] = fmaf(-A[j][i], w1[j], w1[i]);
      }
    }
    // Gram-Schmidt
    float n0 = 0.f, d01 = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      n0 = fmaf(w0[i], w0[i], n0);
      d01 = fmaf(w0[i], w1[i], d01);
    }
    const float i0 = rsqrt_approx(fmaxf(n0, 1e-30f));
    const float proj = d01 * i0 * i0;
    float n1 = 0.f;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      w1[i] = fmaf(-proj, w0[i], w1[i]);
      w0[i] *= i0;
      n1 = fmaf(w1[i], w1[i], n1);
    }
    const float i1 = rsqrt_approx(fmaxf(n1, 1e-30f));
#pragma unroll
    for (int i = 0; i < 10; ++i) w1[i] *= i1;
  }
  // ---- Rayleigh-Ritz inside the pair: S = W^T R R^T W, rotate so that S is diagonal ---------------
  {
    float s00 = 0.f, s01 = 0.f, s11 = 0.f;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
      float g0 = 0.f, g1 = 0.f;
#pragma unroll
      for (int k = 0; k <= j; ++k) {
        g0 = fmaf(A[k][j], w0[k], g0);
        g1 = fmaf(A[k][j], w1[k], g1);
      }
      s00 = fmaf(g0, g0, s00);
      s01 = fmaf(g0, g1, s01);
      s11 = fmaf(g1, g1, s11);
    }
    const float h = s11 - s00, gg = s01 + s01;
    const float q = sqrt_approx(fmaf(h, h, fmaf(gg, gg, 1e-37f)));
    const float t = gg * rcp_approx(h + copysignf(q, h));
    const float c = rsqrt_approx(fmaf(t, t, 1.0f)), sn = c * t;
    const bool swap = fmaf(-t, s01, s00) > fmaf(t, s01, s11);  // v2 = the smaller Ritz value
    float y2[12], y3[12];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const float a = c * w0[i] - sn * w1[i], b = sn * w0[i] + c * w1[i];
      y2[i] = swap ? b : a;
      y3[i] = swap ? a : b;
    }
    y2[10] = y2[11] = y3[10] = y3[11] = 0.f;
    // both back-transforms read tau from work[24..33] before v2 is written over it
    apply_q(y2);
    apply_q(y3);
#pragma unroll
    for (int r = 0; r < 12; ++r) work[24 + eig_row_to_coord(r)] = y2[r], work[36 + eig_row_to_coord(r)] = y3[r];
  }
}

template <int kEig>  // 0: Householder QR + inverse iteration (default), 1: one-sided Jacobi SVD of M^T
__global__ void __maxnreg__(SPE_T1_REGS)
hypothesis_kernel_t1(DevModel m, const float* __restrict__ kpts, int H, int h_begin, int hblocks, const int32_t* __restrict__ need,
                     float thr2, int sweeps, RansacWorkspace ws) {
  // One warp = one work item: 32 consecutive hypotheses [h_begin + 32*hb, +32) of frame b.  Warps of
  // a CTA are independent (different frames in general), each with its own copy of the frame data.
  // Warps are independent, each with its own slice of (dynamic) shared memory, so the CTA size is a
  // launch-time choice (launch_ransac_score): small CTAs pack around the previous batch's refit CTAs.
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
  const int h = h_begin + hb * 32 + lane;
  // adaptive second pass: only hypotheses below the frame's remaining budget are scored
  const int limit = need ? min(need[b], H) : H;
  if (h_begin + hb * 32 >= limit) return;
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
  if (h >= H) return;
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
  float A[12][10], d[10];
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
  ws.masks[(size_t)b * H + h] = bits;
  ws.counts[(size_t)b * H + h] = (uint8_t)__popc(bits);
}

// ------------------------------------------------------------------------------------------
// 2c. adaptive budget (optional): after the first kFirstPass hypotheses of every frame have been
// scored, replay cv2's acceptance loop over them.  need[b] = how many hypotheses cv2 could still
// look at (its shrinking iteration budget, capped at H), or 0 when its loop has already ended.
// The second pass scores only those; select_refit then reads exactly the entries cv2 would read,
// so the result is identical to scoring all H.
constexpr int kFirstPass = 32;
__device__ int update_num_iters(double p, double ep, int max_iters);

__global__ void __launch_bounds__(128) budget_kernel(RansacWorkspace ws, int B, int H, double confidence, int32_t* need) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int n = ws.n[b];
  int out = 0;
  if (n > kModelPoints) {
    const uint8_t* counts = ws.counts + (size_t)b * H;
    int niters = H, max_good = 0;
    const int first = min(kFirstPass, H);
    for (int h = 0; h < min(niters, first); ++h) {
      const int g = counts[h];
      if (g > max(max_good, kModelPoints - 1)) {
        max_good = g;
        niters = update_num_iters(confidence, (double)(n - g) / n, niters);
      }
    }
    out = niters > first ? min(niters, H) : 0;
  }
  need[b] = out;
}

// ------------------------------------------------------------------------------------------
// 3. selection + final refit (float64, one thread per frame)

// OpenCV's JacobiSVD on the rows of a symmetric n x n matrix (App. B.4): returns the rotated rows
// normalised (= rows of U^T) sorted by descending singular value.  Sign-defining for the PCA axes.
__device__ void cv_jacobi_rows(double* A, double* w, int n) {
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

// Eigenvectors of the four SMALLEST eigenvalues of a symmetric N x N matrix (float64), for MtM in
// the final refit: Householder tridiagonalisation (the classic tred2 reduction, reflectors kept,
// Q never formed), eigenvalues by implicit QL without vectors (tql1), the four wanted
// eigenvectors of the tridiagonal matrix by inverse iteration (pivoted tridiagonal LU, three
// iterations, orthogonalised against the vectors already found) and back-transformation through
// the stored reflectors.  ~5x fewer instructions than accumulating all 12 eigenvectors through QL,
// ~20x fewer than the one-sided Jacobi OpenCV runs here; residuals |A v - w v| / |A| ~ 4e-16 on EPnP
// matrices (tools prototype).  Eigenvector SIGNS are irrelevant for MtM (they matter for the 3x3
// PCA, which keeps cv_jacobi_rows).  V is destroyed.  out[k] <-> k-th smallest eigenvalue.
// A per-thread 12x12 float64 matrix kept in SHARED memory, element-major so that the 32 threads of
// a warp touch 32 consecutive doubles: the refit kernel's working set per thread (~6 KB of local
// arrays) does not fit the L1 next to the 132 KB shared-memory carveout, and the tridiagonalisation
// sweeps this matrix many times (long_scoreboard was its top stall with the matrix in local memory).
struct SmemMat12 {
  double* base;  // &storage[0][thread]
  int stride;    // threads per CTA
  __device__ __forceinline__ double& operator()(int r, int c) const { return base[(r * 12 + c) * stride]; }
};

struct LocalMat12 {  // the same matrix in the thread's local memory (small footprint, see select_refit_kernel)
  double (*a)[12];
  __device__ __forceinline__ double& operator()(int r, int c) const { return a[r][c]; }
};
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

// RANSACUpdateNumIters (App. B.6)
__device__ int update_num_iters(double p, double ep, int max_iters) {
  p = fmin(fmax(p, 0.0), 1.0);
  ep = fmin(fmax(ep, 0.0), 1.0);
  double num = fmax(1.0 - p, 2.2250738585072014e-308);
  double denom = 1.0 - pow(1.0 - ep, (double)kModelPoints);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
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
  if constexpr (MatTraits<Mat>::kSharedPerFrame) {
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
  } else {
  // MtM (12x12): upper triangle over this lane's points, summed over the group, mirrored
  for (int r = 0; r < 12; ++r)
    for (int c = 0; c < 12; ++c) mtm(r, c) = 0.0;
  for (int i = sub; i < n; i += 4) {
    double r1[12], r2[12];
    for (int j = 0; j < 4; ++j) {
      r1[3 * j] = al[i][j] * fu, r1[3 * j + 1] = 0.0, r1[3 * j + 2] = al[i][j] * (uc - us[i][0]);
      r2[3 * j] = 0.0, r2[3 * j + 1] = al[i][j] * fv, r2[3 * j + 2] = al[i][j] * (vc - us[i][1]);
    }
    for (int r = 0; r < 12; ++r)
      for (int c = r; c < 12; ++c) mtm(r, c) += r1[r] * r1[c] + r2[r] * r2[c];
  }
  __syncwarp(gmask);
  for (int r = 0; r < 12; ++r)
    for (int c = r; c < 12; ++c) {
      double x = mtm(r, c);
      x += __shfl_xor_sync(gmask, x, 1);
      x += __shfl_xor_sync(gmask, x, 2);
      mtm(r, c) = x;
      mtm(c, r) = x;
    }
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

// The kernel is serial-latency bound: its duration is the time ONE frame's dependent chain of
// float64 instructions takes, whatever the batch size (spreading whole frames over more warps was
// measured to buy nothing).  What shortens it is splitting a frame over lanes: every frame gets a
// group of 4 lanes that run the sequential parts redundantly and share MtM accumulation, the four
// eigenvectors and the three beta variants (see epnp_f64).
// kSmemMat: keep the 12x12 working matrix in shared memory (36 KB per 32-thread CTA: lowest latency
// when the kernel has the GPU to itself) or in local memory (no shared-memory footprint: the
// variant used when it runs as a background tail under the next batch's hypothesis kernel, whose
// three CTAs per SM leave no room for 36 KB more).
// kMatMode 2: one shared-memory copy per frame (FrameMat12, 9.3 KB per warp): what both launch shapes use now;
// 0 and 1 stay selectable with SPE_REFIT_MAT for A/B measurements.
template <int kMatMode>
__global__ void __maxnreg__(SPE_REFIT_REGS) select_refit_kernel(DevModel m, RansacArgs a, RansacWorkspace ws) {
  const int lane = threadIdx.x & 31, sub = lane & 3;
  const unsigned gmask = 0xFu << (lane & ~3);
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  if (b >= a.B) return;  // whole groups leave together (blockDim is a multiple of 4)
  const int n = ws.n[b];
  const unsigned vis = ws.vis[b];
  int status = SPE_FRAME_OK, winner = -1;
  unsigned inl = 0;
  if (n < 4) {
    status = SPE_FRAME_TOO_FEW_POINTS;
  } else if (n == 4) {
    status = SPE_FRAME_P3P_UNSUPPORTED;
  } else if (n == kModelPoints) {
    inl = vis;  // cv2: model_points == npoints -> plain solvePnP, every point an inlier
    winner = 0;
  } else {
    // sequential acceptance over the inlier counts (App. B.6)
    const uint8_t* counts = ws.counts + (size_t)b * a.H;
    int niters = a.H, max_good = 0;
    for (int h = 0; h < niters; ++h) {
      const int g = counts[h];
      if (g > max(max_good, kModelPoints - 1)) {
        winner = h;
        max_good = g;
        niters = update_num_iters(a.confidence, (double)(n - g) / n, niters);
      }
    }
    if (winner < 0) status = SPE_FRAME_NO_MODEL;
    else inl = ws.masks[(size_t)b * a.H + winner];
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
        und[k][0] = n == kModelPoints ? (double)(float)q.x : q.x;
        und[k][1] = n == kModelPoints ? (double)(float)q.y : q.y;
        ++k;
      }
    if constexpr (kMatMode == 2) {
      extern __shared__ double s_mat[];  // [blockDim.x / 4][kFrameMatStride]
      epnp_f64(k, pw, und, m.cam, R, t, sub, gmask, FrameMat12{s_mat + (threadIdx.x >> 2) * kFrameMatStride});
    } else if constexpr (kMatMode == 1) {
      extern __shared__ double s_mat[];  // [144][blockDim.x]
      epnp_f64(k, pw, und, m.cam, R, t, sub, gmask, SmemMat12{s_mat + threadIdx.x, (int)blockDim.x});
    } else {
      double mtm[12][12];
      epnp_f64(k, pw, und, m.cam, R, t, sub, gmask, LocalMat12{mtm});
    }
    if (a.refine_lm) refine_lm_f64(k, pw, img, m.cam, R, t);
  }
  if (sub != 0) return;
  double q[4] = {1, 0, 0, 0};
  if (status == SPE_FRAME_OK) rotation_to_quat(R, q);
  float* o = a.pose7 + (size_t)b * 7;
  const bool ok = status == SPE_FRAME_OK;
  for (int i = 0; i < 4; ++i) o[i] = ok ? (float)q[i] : 0.f;
  for (int i = 0; i < 3; ++i) o[4 + i] = ok ? (float)t[i] : 0.f;
  a.inlier_mask[b] = inl;
  a.status[b] = status;
  if (a.winner) a.winner[b] = winner;
  if (a.rt) {
    double* r = a.rt + (size_t)b * 12;
    for (int i = 0; i < 9; ++i) r[i] = ok ? R[i / 3][i % 3] : 0.0;
    for (int i = 0; i < 3; ++i) r[9 + i] = ok ? t[i] : 0.0;
  }
}

__global__ void debug_scores_kernel(RansacWorkspace ws, long long total, int32_t* counts, uint32_t* masks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (counts) counts[i] = ws.counts[i];
  if (masks) masks[i] = ws.masks[i];
}

}  // namespace

cudaError_t launch_ransac_score(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0) return cudaSuccess;
  DevModel dm{m.d_landmarks, m.d_subsets, m.J, m.max_hyp, m.cam, reinterpret_cast<const float4*>(m.d_ctrl)};
  const int wpb = 4;
  frame_prep_kernel<<<(a.B + wpb - 1) / wpb, wpb * 32, 0, stream>>>(dm, a.kpts, a.B, a.conf_floor, ws);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (m.J > kModelPoints) {
    const float thr2 = a.reproj_err * a.reproj_err;
    if (a.kernel_variant == 1) {  // 4 lanes per hypothesis (kept for A/B measurements)
      const int hblocks = (a.H + kHypPerCta - 1) / kHypPerCta;
      const long long ctas = (long long)a.B * hblocks;
      if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
      hypothesis_kernel<<<(unsigned)ctas, kHypPerCta * kGroup, 0, stream>>>(dm, a.kpts, a.H, hblocks, thr2, a.jacobi_sweeps, ws);
    } else {  // one thread per hypothesis, one warp per (frame, 32 hypotheses)
      static PerDeviceOnce once;  // same shared-memory/L1 split as the decode kernel (decode.cuh)
      e = once.run(m.device, [] {
        cudaError_t r = cudaFuncSetAttribute(hypothesis_kernel_t1<0>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
        if (r == cudaSuccess) r = cudaFuncSetAttribute(hypothesis_kernel_t1<1>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
        if (r == cudaSuccess) r = cudaFuncSetAttribute(hypothesis_kernel_t1<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT1MaxWarps * kT1WarpBytes);
        if (r == cudaSuccess) r = cudaFuncSetAttribute(hypothesis_kernel_t1<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT1MaxWarps * kT1WarpBytes);
        return r;
      });
      if (e != cudaSuccess) return e;
      const bool jacobi = a.kernel_variant == 2;  // full SVD of M^T (kept for A/B measurements)
      const int max_warps = a.t1_warps >= 1 && a.t1_warps <= kT1MaxWarps ? a.t1_warps : kT1MaxWarps;
      int dev = 0, num_sms = 0;
      e = current_device(dev, num_sms);
      if (e != cudaSuccess) return e;
      auto launch = [&](int h_begin, int h_count, const int32_t* need) -> cudaError_t {
        const int hblocks = (h_count + 31) / 32;
        // small batches: spread the warps over the SMs instead of packing 12 of them into one CTA
        const long long items = (long long)a.B * hblocks;
        const int kWarps = (int)std::min<long long>(max_warps, std::max<long long>(1, (items + num_sms - 1) / num_sms));
        const size_t smem = (size_t)kWarps * kT1WarpBytes;
        const long long ctas = (items + kWarps - 1) / kWarps;
        if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
        if (ctas == 0) return cudaSuccess;
        if (jacobi)
          hypothesis_kernel_t1<1><<<(unsigned)ctas, kWarps * 32, smem, stream>>>(dm, a.kpts, a.H, h_begin, hblocks, need, thr2, a.jacobi_sweeps, ws);
        else
          hypothesis_kernel_t1<0><<<(unsigned)ctas, kWarps * 32, smem, stream>>>(dm, a.kpts, a.H, h_begin, hblocks, need, thr2, a.jacobi_sweeps, ws);
        return cudaGetLastError();
      };
      if (a.adaptive && a.H > kFirstPass) {
        e = launch(0, kFirstPass, nullptr);
        if (e != cudaSuccess) return e;
        budget_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(ws, a.B, a.H, a.confidence, ws.need);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        e = launch(kFirstPass, a.H - kFirstPass, ws.need);
      } else {
        e = launch(0, a.H, nullptr);
      }
      if (e != cudaSuccess) return e;
    }
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_ransac_select_refit(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0) return cudaSuccess;
  DevModel dm{m.d_landmarks, m.d_subsets, m.J, m.max_hyp, m.cam, reinterpret_cast<const float4*>(m.d_ctrl)};
  // Without a common carveout the kernel flips idle SMs to an all-L1 split and the next batch's
  // decode CTAs must wait for it to finish: measured 0.12 -> 0.32 ms decode when overlapped.
  static PerDeviceOnce once;
  const cudaError_t ce = once.run(m.device, [] {
    cudaError_t r = cudaFuncSetAttribute(select_refit_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(select_refit_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(select_refit_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(select_refit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8 * kFrameMatStride * (int)sizeof(double));
    return r;
  });
  if (ce != cudaSuccess) return ce;
  // 4 lanes per frame.  Background tail: the warps are packed 8 to a CTA (8 x 255 registers = a whole SM) so that
  // they hide each other's latency on FEW SMs instead of each blocking a 21 K-register CTA slot of the next
  // batch's hypothesis kernel on EVERY SM (profiles/step_r1.md, tools/overlap_probe.py).  Alone: one warp per CTA.
  constexpr int kMaxTailWarps = 65536 / (32 * ((SPE_REFIT_REGS + 7) / 8 * 8));  // one CTA = the register file of one SM
  int warps = a.refit_background ? kMaxTailWarps : 1;
  if (a.tail_warps >= 1 && a.tail_warps <= kMaxTailWarps) warps = a.tail_warps;
  const int threads = 32 * warps;
  const int ctas = (a.B * 4 + threads - 1) / threads;
  static const int mat_mode = [] {
    const char* v = getenv("SPE_REFIT_MAT");  // dev knob: 0 local memory, 1 per-thread shared copies, 2 per-frame shared copy
    return v ? atoi(v) : 2;
  }();
  if (mat_mode == 0 || (mat_mode == 1 && warps > 1)) {
    select_refit_kernel<0><<<ctas, threads, 0, stream>>>(dm, a, ws);
  } else if (mat_mode == 1) {
    select_refit_kernel<1><<<ctas, 32, sizeof(double) * 144 * 32, stream>>>(dm, a, ws);
  } else {
    select_refit_kernel<2><<<ctas, threads, sizeof(double) * kFrameMatStride * 8 * warps, stream>>>(dm, a, ws);
  }
  return cudaGetLastError();
}

cudaError_t launch_ransac_epnp(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  const cudaError_t e = launch_ransac_score(m, a, ws, stream);
  return e != cudaSuccess ? e : launch_ransac_select_refit(m, a, ws, stream);
}

cudaError_t launch_debug_scores(const RansacWorkspace& ws, int B, int H, int32_t* counts, uint32_t* masks, cudaStream_t stream) {
  const long long total = (long long)B * H;
  if (total == 0) return cudaSuccess;
  debug_scores_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(ws, total, counts, masks);
  return cudaGetLastError();
}

}  // namespace spe
