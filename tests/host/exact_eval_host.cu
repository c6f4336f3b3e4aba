// Host build of the float64 replay's arithmetic (csrc/ransac_exact_eval.cuh + csrc/ransac_common.cuh), for the CPU test
// suite: the very functions replay_kernel runs, compiled for the host by nvcc (no device code is launched), driven by
// the same sequential loop.  Test infrastructure only — the product never loads this file.
//
//   nvcc -O2 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off -shared -o exact_eval_host.so exact_eval_host.cu
#include <stdint.h>
#include <string.h>

#include "../../spacecraft-pose-estimation_b200/csrc/ransac_common.cuh"
#include "../../spacecraft-pose-estimation_b200/csrc/p3p_f64.cuh"
#include "../../spacecraft-pose-estimation_b200/csrc/ransac_exact_eval.cuh"

extern "C" {

// One frame: obj [n,3] float32-rounded landmarks (as doubles), und [n,2] float64 undistorted normalised points,
// img [n,2] float32 pixels, subsets [iterations,5] minimal sets for n points.  Returns the winner (-1 none) and fills
// mask (bit i = point i of the n), visited, and optionally the per-hypothesis masks of the visited prefix.
int spe_host_replay_frame(const double* obj, const double* und, const float* img, int n, const double* cam9, const uint8_t* subsets, int iterations,
                          float reproj_err, double confidence, uint32_t* mask_out, int32_t* visited_out, uint32_t* hyp_masks, int hyp_masks_len) {
  spe::Camera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8]};
  spe::FramePoints f;
  memset(&f, 0, sizeof(f));
  for (int k = 0; k < n; ++k) {
    for (int c = 0; c < 3; ++c) f.pw[k][c] = obj[3 * k + c];
    f.us[k][0] = (double)(float)und[2 * k] * cam.fx + cam.cx;
    f.us[k][1] = (double)(float)und[2 * k + 1] * cam.fy + cam.cy;
    f.img[k][0] = img[2 * k], f.img[k][1] = img[2 * k + 1];
    f.id[k] = (uint8_t)k;
  }
  const float thr2 = reproj_err * reproj_err;
  int niters = iterations, max_good = 0, winner = -1;
  uint32_t best = 0;
  for (int h = 0; h < niters; ++h) {
    const unsigned bits = spe::hypothesis_f64(cam, f, n, subsets + (size_t)h * 5, thr2);
    if (hyp_masks && h < hyp_masks_len) hyp_masks[h] = bits;
    const int g = __builtin_popcount(bits);
    if (g > (max_good > 4 ? max_good : 4)) {
      winner = h, max_good = g, best = bits;
      niters = spe::update_num_iters(confidence, (double)(n - g) / n, niters);
    }
  }
  *mask_out = best;
  *visited_out = niters;
  return winner;
}

// One hypothesis with its intermediates (dbg[75]: v[48], (err, betas[4]) x 3, R[9], t[3]); returns the inlier mask.
unsigned spe_host_hypothesis(const double* obj, const double* und, const float* img, int n, const double* cam9, const uint8_t* subset, float reproj_err,
                             double* dbg) {
  spe::Camera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8]};
  spe::FramePoints f;
  memset(&f, 0, sizeof(f));
  for (int k = 0; k < n; ++k) {
    for (int c = 0; c < 3; ++c) f.pw[k][c] = obj[3 * k + c];
    f.us[k][0] = (double)(float)und[2 * k] * cam.fx + cam.cx;
    f.us[k][1] = (double)(float)und[2 * k + 1] * cam.fy + cam.cy;
    f.img[k][0] = img[2 * k], f.img[k][1] = img[2 * k + 1];
    f.id[k] = (uint8_t)k;
  }
  return spe::hypothesis_f64(cam, f, n, subset, reproj_err * reproj_err, dbg);
}

// cv2's n == 4 branch: obj [4,3] (float32-rounded, as doubles), und [4,2] float64 undistorted normalised points (rounded to
// float32 inside, as cv2 keeps the input dtype); Rt_out [12] = R row-major, then t.
int spe_host_p3p(const double* obj, const double* und, const double* cam9, double* Rt_out) {
  spe::Camera cam{cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], cam9[8]};
  double X[4][3], us[4][2], R[3][3], t[3];
  for (int k = 0; k < 4; ++k) {
    for (int c = 0; c < 3; ++c) X[k][c] = obj[3 * k + c];
    us[k][0] = (double)(float)und[2 * k] * cam.fx + cam.cx;
    us[k][1] = (double)(float)und[2 * k + 1] * cam.fy + cam.cy;
  }
  if (!spe::solve_p3p_f64(cam, X, us, R, t)) return 0;
  for (int i = 0; i < 9; ++i) Rt_out[i] = R[i / 3][i % 3];
  for (int i = 0; i < 3; ++i) Rt_out[9 + i] = t[i];
  return 1;
}

int spe_host_quartic(const double* c5, double* roots4) {
  double c[5], x[4];
  for (int i = 0; i < 5; ++i) c[i] = c5[i];
  const int n = spe::p3p::quartic_real_roots(c, x);
  for (int i = 0; i < n; ++i) roots4[i] = x[i];
  return n;
}

}  // extern "C"
