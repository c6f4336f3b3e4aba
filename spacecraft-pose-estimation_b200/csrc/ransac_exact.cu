// Batched RANSAC-EPnP for sm_100a, part 2: float64 replay of cv2's own RANSAC loop (SPE_FLAG_EXACT).
//
// cv2.solvePnPRansac (pose_estimation/export_predicted_poses_real.py:199-201) is sequential and adaptive: hypothesis h
// is looked at only while h < niters, and every accepted model (first strictly better inlier count) shrinks niters
// (SURVEY App. B.6).  On the benchmark data it looks at 6.5 hypotheses per frame on average (90 % of the frames: <= 12)
// and decides on float64 arithmetic.  This kernel does exactly that, frame by frame, in float64:
//
//   * one thread evaluates one hypothesis: EPnP on the five points of the minimal set in float64 (control points from
//     OpenCV's sign-defining Jacobi, M^T -> Householder QR + four-vector inverse iteration for the four vectors, three
//     beta initialisations, five Gauss-Newton steps each with Householder QR, Procrustes, best of three by mean
//     reprojection error — OpenCV's sequence, oracle/epnp_ref.py), then cv2.projectPoints of all n visible points in
//     float64, rounded to float32, squared error in float32 <= thr^2 (ransac_exact_eval.cuh);
//   * the loop is sequential per frame but its length is only known at run time (8 hypotheses finish 75 % of the
//     benchmark frames, a 5-inlier model of 11 points needs 235, a frame without a model all 10000), so it runs in
//     PHASES of growing width — hypotheses [0,8), [8,40), [40,168), [168,680), then 2048 at a time — one kernel launch
//     each, no host synchronisation.  Inside a phase every (frame, hypothesis) still wanted by the frame's budget AT THE
//     START of the phase is evaluated in parallel (the budget only shrinks, so nothing cv2 looks at is missed and at
//     most 4x too much is evaluated); the warp that finishes a frame's last block then walks the phase's results in
//     order with cv2's acceptance rule and RANSACUpdateNumIters, and either finishes the frame or appends it to the next
//     phase's work list.  A phase whose list is empty is a kernel that exits at once.
//   * work is drawn from counters in the caller's workspace (zeroed by the launcher on the call's stream): a persistent
//     grid of warps claims (frame, block) items until the phase is done.
//
// The FP32 scores of ransac_score.cu are not read here: which hypotheses decide is cv2's rule, and those are
// re-evaluated in float64 whatever FP32 said about them.
// Why a replay and not per-hypothesis bit-parity: a 5-point EPnP has an exact 2-D null space, so ~8 % of cv2's own
// hypothesis results are decided by rounding noise (measured: a float64 NumPy restatement reproduces cv2's
// per-hypothesis inlier count on 92 % of minimal sets whether it uses a port of OpenCV's Jacobi SVD or LAPACK's
// eigensolver).  What is reproducible is the outcome of the loop — the winner's inlier set — and that needs float64 on
// the hypotheses cv2 looks at: 99.0 % of close-range frames against 77.6 % with FP32 scores (profiles/parity_r2.md).
#include <cuda_runtime.h>

#include <algorithm>
#include <math.h>
#include <stdint.h>

#include "../../include/spe_b200.h"
#include "decode.cuh"
#include "device_util.cuh"
#include "epnp_f64.cuh"
#include "epnp_math.cuh"
#include "ransac.cuh"
#include "ransac_common.cuh"
#include "ransac_exact_eval.cuh"

namespace spe {

namespace {

constexpr int kReplayWarps = 4;  // 128 threads x 255 registers: two CTAs per SM
constexpr int kFirstWidth = 8;   // hypotheses of phase 0: four frames per warp

struct ReplayState {  // per frame, between phases
  int32_t niters, max_good, winner;
  uint32_t best_mask;
};
static_assert(sizeof(ReplayState) == 16, "ReplayState is carved as 16 bytes per frame");

__device__ __forceinline__ void load_frame(const DevModel& m, const RansacWorkspace& ws, int b, int n, FramePoints& f, int sub, int lanes) {
  const unsigned vis = ws.vis[b];
  for (int k = sub; k < n; k += lanes) {
    const int j = __fns(vis, 0, k + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) f.pw[k][c] = (double)m.landmarks[3 * j + c];
    const double2 q = ws.und[(size_t)b * m.J + j];
    // hypotheses see the float32-rounded undistorted point (cv2 keeps the input dtype), App. B.3a
    f.us[k][0] = (double)(float)q.x * m.cam.fx + m.cam.cx;
    f.us[k][1] = (double)(float)q.y * m.cam.fy + m.cam.cy;
    const float2 px = ws.img[(size_t)b * m.J + j];
    f.img[k][0] = px.x, f.img[k][1] = px.y;
    f.id[k] = (uint8_t)j;
  }
}

// cv2's acceptance loop over one segment of consecutive hypotheses held one per lane (mk = inlier mask of hypothesis h;
// lanes of the segment = seg_bits).  Every lane of the warp must call this (different segments may belong to
// different frames); `live` = the lane's segment belongs to a frame.
__device__ __forceinline__ void accept_segment(unsigned mk, int h, bool live, unsigned seg_bits, int lane, int n, double confidence, ReplayState& st) {
  const int g = __popc(mk);
  int consumed = -1;  // lanes up to here have been walked
  for (;;) {
    const bool c = live && lane > consumed && h < st.niters && g > max(st.max_good, kModelPoints - 1);
    const unsigned cand = __ballot_sync(kFullMask, c) & seg_bits;
    if (__all_sync(kFullMask, cand == 0u)) break;
    const int first = cand ? __ffs(cand) - 1 : lane;
    const int g1 = __shfl_sync(kFullMask, g, first);
    const unsigned m1 = __shfl_sync(kFullMask, mk, first);
    const int h1 = __shfl_sync(kFullMask, h, first);
    if (cand) {
      st.winner = h1;
      st.max_good = g1;
      st.best_mask = m1;
      st.niters = update_num_iters(confidence, (double)(n - g1) / n, st.niters);
      consumed = first;
    }
  }
}

// the frame has been walked up to hypothesis `end`: finished, or on to the next phase
__device__ __forceinline__ void finish_or_continue(const RansacWorkspace& ws, int b, const ReplayState& st, int end, int iterations, int next_phase) {
  if (end >= st.niters || end >= iterations) {
    ws.x_winner[b] = st.winner;
    ws.x_mask[b] = st.best_mask;
    ws.x_visited[b] = st.niters;
  } else {
    reinterpret_cast<ReplayState*>(ws.x_state)[b] = st;
    const unsigned at = atomicAdd(ws.claim + kClaimActive + next_phase, 1u);
    ws.x_active[(size_t)(next_phase & 1) * ws.frames + at] = b;
  }
}

// Phase p: hypotheses [lo, lo + W) of every frame in the phase's work list (phase 0: of every frame).
__global__ void __launch_bounds__(kReplayWarps * 32) replay_phase_kernel(DevModel m, RansacArgs a, RansacWorkspace ws, int p, int lo, int W) {
  __shared__ FramePoints s_frames[kReplayWarps * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float thr2 = a.reproj_err * a.reproj_err;
  const int L = W < 32 ? W : 32;             // lanes (hypotheses) per frame and block
  const int F = 32 / L;                      // frames per warp item
  const int nb = W / L;                      // blocks per frame
  const int sub = lane % L, grp = lane / L;  // lane within the frame's segment, segment within the warp
  const unsigned seg_bits = (L == 32 ? kFullMask : ((1u << L) - 1u)) << (grp * L);
  FramePoints& f = s_frames[warp * 4 + grp];
  const int n_frames = p == 0 ? a.B : (int)ws.claim[kClaimActive + p];
  const long long items = p == 0 ? (n_frames + F - 1) / F : (long long)n_frames * nb;
  const int32_t* active = ws.x_active + (size_t)(p & 1) * ws.frames;
  for (;;) {
    long long item = 0;
    if (lane == 0) item = (long long)atomicAdd(ws.claim + kClaimItem + p, 1u);
    item = __shfl_sync(kFullMask, item, 0);
    if (item >= items) break;
    // which frame, which block
    int b = -1, blk = 0, row = 0;
    if (p == 0) {
      row = (int)item * F + grp;
      b = row < n_frames ? row : -1;
    } else {
      row = (int)(item / nb);
      blk = (int)(item - (long long)row * nb);
      b = active[row];
    }
    int n = b >= 0 ? ws.n[b] : 0;
    ReplayState st{a.iterations, 0, -1, 0u};
    if (b >= 0 && n <= kModelPoints) {  // no RANSAC for this frame: select_refit_kernel settles it
      if (sub == 0) ws.x_winner[b] = -1, ws.x_mask[b] = 0u, ws.x_visited[b] = 0;
      b = -1;
    }
    if (b >= 0 && p > 0) st = reinterpret_cast<const ReplayState*>(ws.x_state)[b];
    __syncwarp();  // the previous item's readers are done with the warp's shared slots
    if (b >= 0) load_frame(m, ws, b, n, f, sub, L);
    __syncwarp();
    const int h = lo + blk * L + sub;
    unsigned bits = 0;
    if (b >= 0 && h < st.niters && h < a.iterations) {
      const uint8_t* subset = m.subsets + ((size_t)(n - 6) * m.max_hyp + h) * kModelPoints;
      bits = hypothesis_f64(m.cam, f, n, subset, thr2);
    }
    if (nb == 1) {
      // the whole phase of this frame sits in the warp: walk it here
      accept_segment(bits, h, b >= 0, seg_bits, lane, n, a.confidence, st);
      if (b >= 0 && sub == 0) finish_or_continue(ws, b, st, lo + W, a.iterations, p + 1);
    } else {
      uint32_t* masks = ws.x_masks + (size_t)row * kReplayMaxWidth;
      masks[blk * 32 + lane] = bits;
      __threadfence();
      unsigned done = 0;
      if (lane == 0) done = atomicAdd(ws.x_done + b, 1u);
      done = __shfl_sync(kFullMask, done, 0);
      if (done == (unsigned)nb - 1u) {  // this warp completed the frame's last block: walk the whole phase in order
        __threadfence();
        for (int c = 0; c < nb; ++c) {
          const int hc = lo + c * 32 + lane;
          const unsigned mk = __ldcg(masks + c * 32 + lane);
          accept_segment(mk, hc, true, kFullMask, lane, n, a.confidence, st);
          if (hc - lane + 32 >= st.niters) break;  // (uniform: st is replicated on the lanes)
        }
        if (lane == 0) {
          ws.x_done[b] = 0u;
          finish_or_continue(ws, b, st, lo + W, a.iterations, p + 1);
        }
      }
    }
  }
}

}  // namespace

cudaError_t launch_ransac_replay(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0 || m.J <= kModelPoints) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(ws.claim, 0, sizeof(uint32_t) * kClaimWords, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(ws.x_done, 0, sizeof(uint32_t) * (size_t)a.B, stream);
  if (e != cudaSuccess) return e;
  static PerDeviceOnce once;
  e = once.run(m.device, [] { return cudaFuncSetAttribute(replay_phase_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct); });
  if (e != cudaSuccess) return e;
  int dev = 0, num_sms = 0;
  e = current_device(dev, num_sms);
  if (e != cudaSuccess) return e;
  const DevModel dm = dev_model(m);
  int lo = 0, W = kFirstWidth;
  for (int p = 0; lo < a.iterations && p < kReplayMaxPhases; ++p) {
    // persistent grid: two CTAs per SM at most; phase 0 never needs more warps than it has items
    long long warps = p == 0 ? ((long long)a.B + 3) / 4 : (long long)2 * num_sms * kReplayWarps;
    const int ctas = (int)std::max<long long>(1, std::min<long long>(2 * num_sms, (warps + kReplayWarps - 1) / kReplayWarps));
    replay_phase_kernel<<<ctas, kReplayWarps * 32, 0, stream>>>(dm, a, ws, p, lo, W);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    lo += W;
    W = std::min(W * 4, kReplayMaxWidth);
  }
  return cudaSuccess;
}

}  // namespace spe
