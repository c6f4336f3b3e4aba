// Development-only kernel variants (compiled with -DSPE_DEV, never in the shipped library): the 4-lanes-per-hypothesis
// kernel of the first design and the full one-sided Jacobi SVD of M^T.  They are the measured alternatives that
// profiles/solver_r1.md tabulates; the product path is ransac_score.cu.  Included by ransac_score.cu inside its namespaces.
#pragma once
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

