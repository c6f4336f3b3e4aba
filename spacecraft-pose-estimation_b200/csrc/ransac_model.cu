// Host side of the RANSAC-EPnP model: OpenCV's fixed-seed minimal sets, the per-model control-point table, the
// duplicate-free hypothesis lists, device upload, and the workspace carve-up (include/spe_b200.h, ransac.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/spe_b200.h"
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

// Duplicate-free lists for one point count (see Model): sets are compared as sets (sorted ids packed 6 bits each).
int build_unique(const uint8_t* subsets, int num, uint16_t* uniq, uint16_t* slot) {
  std::vector<std::pair<uint32_t, uint16_t>> seen;  // (packed sorted set, slot); sorted by key for the lookup
  seen.reserve(num);
  int count = 0;
  for (int h = 0; h < num; ++h) {
    uint8_t s[kModelPoints];
    memcpy(s, subsets + (size_t)h * kModelPoints, kModelPoints);
    std::sort(s, s + kModelPoints);
    uint32_t key = 0;
    for (int k = 0; k < kModelPoints; ++k) key = key << 6 | s[k];
    auto it = std::lower_bound(seen.begin(), seen.end(), std::make_pair(key, (uint16_t)0));
    if (it != seen.end() && it->first == key) {
      slot[h] = it->second;
    } else {
      slot[h] = (uint16_t)count;
      uniq[count] = (uint16_t)h;
      seen.insert(it, std::make_pair(key, (uint16_t)count));
      ++count;
    }
  }
  for (int u = count; u < num; ++u) uniq[u] = 0xffff;  // never below any H: keeps the array sorted for the binary search
  return count;
}

int unique_sets(const Model& m, int n, int H) {
  if (n <= kModelPoints || n > m.J) return 0;
  const uint16_t* u = m.h_uniq.data() + (size_t)(n - 6) * m.max_hyp;
  return (int)(std::lower_bound(u, u + m.h_uniq_count[n - 6], (uint16_t)std::min(H, 0xffff)) - u);
}

template <typename T>
static cudaError_t upload(T*& dst, const std::vector<T>& src) {
  cudaError_t e = cudaMalloc(&dst, src.size() * sizeof(T));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
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
  m.h_uniq.assign((size_t)tables * m.max_hyp, 0);
  m.h_slot.assign((size_t)tables * m.max_hyp, 0);
  m.h_uniq_count.assign(tables, 0);
  for (int n = 6; n <= m.J; ++n) {
    uint8_t* sub = m.h_subsets.data() + (size_t)(n - 6) * m.max_hyp * kModelPoints;
    opencv_minimal_sets(n, m.max_hyp, sub);
    m.h_uniq_count[n - 6] = build_unique(sub, m.max_hyp, m.h_uniq.data() + (size_t)(n - 6) * m.max_hyp, m.h_slot.data() + (size_t)(n - 6) * m.max_hyp);
  }
  if (tables > 0) {
    e = upload(m.d_subsets, m.h_subsets);
    if (e == cudaSuccess) e = upload(m.d_uniq, m.h_uniq);
    if (e == cudaSuccess) e = upload(m.d_slot, m.h_slot);
    if (e != cudaSuccess) return e;
    std::vector<float> ctrl;
    build_control_table(m, ctrl);
    m.ctrl_entries = ctrl.size() / kCtrlEntryFloats;
    e = upload(m.d_ctrl, ctrl);
  }
  return e;
}

void model_free(Model& m) {
  if (m.d_landmarks) cudaFree(m.d_landmarks);
  if (m.d_subsets) cudaFree(m.d_subsets);
  if (m.d_uniq) cudaFree(m.d_uniq);
  if (m.d_slot) cudaFree(m.d_slot);
  if (m.d_ctrl) cudaFree(m.d_ctrl);
  m.d_landmarks = nullptr;
  m.d_subsets = nullptr;
  m.d_uniq = m.d_slot = nullptr;
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
  w.x_winner = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
  w.x_mask = static_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)B));
  w.x_visited = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
  w.claim = static_cast<uint32_t*>(take(sizeof(uint32_t) * kClaimWords));
  w.x_state = take((size_t)kReplayStateBytes * B);
  w.x_frames = take((size_t)kReplayFrameBytes * B);
  w.x_width = static_cast<int32_t*>(take(sizeof(int32_t) * (size_t)B));
  w.x_items = static_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)B * kReplayPlanMax));
  w.x_done = static_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)B));
  w.x_active = static_cast<int32_t*>(take(sizeof(int32_t) * 2 * (size_t)B));
  w.x_masks = static_cast<uint32_t*>(take(sizeof(uint32_t) * (size_t)B * kReplayMaxWidth));
  w.frames = B;
  w.bytes = off;
  return w;
}

size_t ransac_workspace_bytes(int J, int B, int H) { return carve_workspace(nullptr, J, B, H).bytes; }

}  // namespace spe
