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
//   * the loop is sequential per frame but its length is only known at run time (one hypothesis finishes 30 % of the
//     benchmark frames, five finish 69 %, a 5-inlier model of 11 points needs 235, a frame without a model all 10000),
//     so it runs in PHASES, one kernel launch each, no host synchronisation:
//       phase 0   hypotheses [0, w_b) of frame b, where w_b <= 32 is the length cv2's loop WOULD have if the FP32
//                 scores were cv2's own (replay_plan_kernel walks them with cv2's rule; 8 when nothing was scored).
//                 The (frame, hypothesis) pairs of all frames form one dense list, one thread each — no idle lanes;
//       phase p   hypotheses [0, 32) (whatever phase 0 left of them), [32, 160), [160, 672), then 2048 at a time, of the
//                 frames still in the loop.
//     Inside a phase every (frame, hypothesis) still wanted by the frame's budget AT THE START of the phase is
//     evaluated in parallel (the budget only shrinks, so nothing cv2 looks at is missed); the phase's results are then
//     walked in order with cv2's acceptance rule and RANSACUpdateNumIters, and the frame either finishes or joins the
//     next phase's work list.  The FP32 prediction is only a schedule: every hypothesis cv2 looks at is decided in
//     float64, and a frame whose prediction was too short simply continues in the next phase.  A phase whose list is
//     empty is a kernel that exits at once.
//   * work lists and counters live in the caller's workspace and are reset by the first kernel of the call.
// The FP32 scores of ransac_score.cu never decide anything here: which hypotheses count is cv2's rule, and those are
// evaluated in float64 whatever FP32 said about them.
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

#ifndef SPE_REPLAY_CTAS_PER_SM
#define SPE_REPLAY_CTAS_PER_SM 2  // 128 threads x 255 registers: two CTAs per SM (A/B: tools/ab_builds.py)
#endif
#ifndef SPE_REPLAY_WARPS
#define SPE_REPLAY_WARPS 4
#endif
constexpr int kReplayWarps = SPE_REPLAY_WARPS;
constexpr int kReplayCtasPerSm = SPE_REPLAY_CTAS_PER_SM;
constexpr int kPlanWidthNoScores = 8;

struct ReplayState {  // per frame, between phases
  int32_t niters, max_good, winner;
  uint32_t best_mask;
  int32_t next;  // first hypothesis cv2's loop has not walked yet
  int32_t pad[3];
};
static_assert(sizeof(ReplayState) == kReplayStateBytes, "ReplayState is carved as kReplayStateBytes per frame");

__device__ __forceinline__ ReplayState* state_of(const RansacWorkspace& ws, int b) { return reinterpret_cast<ReplayState*>(ws.x_state) + b; }
__device__ __forceinline__ const FramePoints& frame_of(const RansacWorkspace& ws, int b) { return reinterpret_cast<const FramePoints*>(ws.x_frames)[b]; }

// One acceptance of cv2's loop.
__device__ __forceinline__ void accept(ReplayState& st, int h, unsigned mask, int n, double confidence) {
  const int g = __popc(mask);
  if (g > max(st.max_good, kModelPoints - 1)) {
    st.winner = h;
    st.max_good = g;
    st.best_mask = mask;
    st.niters = update_num_iters(confidence, (double)(n - g) / n, st.niters);
  }
}

// the frame has been walked up to hypothesis `end`: finished, or on to phase `next_phase`
__device__ __forceinline__ void finish_or_continue(const RansacWorkspace& ws, int b, ReplayState& st, int end, int iterations, int next_phase) {
  if (end >= st.niters || end >= iterations) {
    ws.x_winner[b] = st.winner;
    ws.x_mask[b] = st.best_mask;
    ws.x_visited[b] = st.niters;
  } else {
    st.next = end;
    *state_of(ws, b) = st;
    const unsigned at = atomicAdd(ws.claim + kClaimActive + next_phase, 1u);
    ws.x_active[(size_t)(next_phase & 1) * ws.frames + at] = b;
  }
}

// ---- phase 0, step 1: per frame (one warp) — compact the visible landmarks once for all phases, walk the FP32 scores
// with cv2's rule to predict how many hypotheses cv2's loop looks at, and append that many (frame, hypothesis) items.
__global__ void __launch_bounds__(128) replay_plan_kernel(DevModel m, RansacArgs a, RansacWorkspace ws) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= a.B) return;
  const int n = ws.n[b];
  if (n <= kModelPoints) {  // no RANSAC for this frame: select_refit_kernel settles it
    if (lane == 0) ws.x_winner[b] = -1, ws.x_mask[b] = 0u, ws.x_visited[b] = 0, ws.x_width[b] = 0;
    return;
  }
  FramePoints& f = reinterpret_cast<FramePoints*>(ws.x_frames)[b];
  const unsigned vis = ws.vis[b];
  if (lane < n) {
    const int j = __fns(vis, 0, lane + 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) f.pw[lane][c] = (double)m.landmarks[3 * j + c];
    const double2 q = ws.und[(size_t)b * m.J + j];
    // hypotheses see the float32-rounded undistorted point (cv2 keeps the input dtype), App. B.3a
    f.us[lane][0] = (double)(float)q.x * m.cam.fx + m.cam.cx;
    f.us[lane][1] = (double)(float)q.y * m.cam.fy + m.cam.cy;
    const float2 px = ws.img[(size_t)b * m.J + j];
    f.img[lane][0] = px.x, f.img[lane][1] = px.y;
    f.id[lane] = (uint8_t)j;
  }
  if (lane != 0) return;
  int width = min(kPlanWidthNoScores, a.iterations);
  if (a.H > 0) {
    // cv2's loop over the FP32 counts of the first min(H, 32) draws: where would it stop?
    const uint8_t* counts = ws.counts + (size_t)b * a.H;
    const uint16_t* slot = m.slot + (size_t)(n - 6) * m.max_hyp;
    int niters = a.iterations, max_good = 0, h = 0;
    const int cap = min(min(a.H, kReplayPlanMax), a.iterations);
    for (; h < cap && h < niters; ++h) {
      const int g = counts[slot[h]];
      if (g > max(max_good, kModelPoints - 1)) {
        max_good = g;
        niters = update_num_iters(a.confidence, (double)(n - g) / n, niters);
      }
    }
    width = max(1, min(niters, cap));
  }
  ws.x_width[b] = width;
  const unsigned at = atomicAdd(ws.claim + kClaimPlanItems, (unsigned)width);
  for (int h = 0; h < width; ++h) ws.x_items[at + h] = ((uint32_t)b << 5) | (uint32_t)h;
}

// ---- phase 0, step 2: one thread per (frame, hypothesis) item
__global__ void __launch_bounds__(kReplayWarps * 32, kReplayCtasPerSm) replay_eval_items_kernel(DevModel m, RansacArgs a, RansacWorkspace ws) {
  const unsigned total = ws.claim[kClaimPlanItems];
  const float thr2 = a.reproj_err * a.reproj_err;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t item = ws.x_items[i];
    const int b = (int)(item >> 5), h = (int)(item & 31u);
    const int n = ws.n[b];
    const uint8_t* subset = m.subsets + ((size_t)(n - 6) * m.max_hyp + h) * kModelPoints;
    ws.x_masks[(size_t)b * kReplayMaxWidth + h] = hypothesis_f64(m.cam, frame_of(ws, b), n, subset, thr2);
  }
}

// ---- phase 0, step 3: per frame, cv2's loop over the evaluated prefix
__global__ void __launch_bounds__(128) replay_scan_items_kernel(RansacArgs a, RansacWorkspace ws) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  const int n = ws.n[b];
  if (n <= kModelPoints) return;
  const int width = ws.x_width[b];
  ReplayState st{a.iterations, 0, -1, 0u, 0, {0, 0, 0}};
  const uint32_t* masks = ws.x_masks + (size_t)b * kReplayMaxWidth;
  for (int h = 0; h < width && h < st.niters; ++h) accept(st, h, masks[h], n, a.confidence);
  finish_or_continue(ws, b, st, width, a.iterations, 1);
}

// cv2's acceptance loop over 32 consecutive hypotheses held one per lane (mk = inlier mask of hypothesis h); the state is
// replicated on the lanes.  Every lane of the warp must call this.
__device__ __forceinline__ void accept_block(unsigned mk, int h, int lane, int n, double confidence, ReplayState& st) {
  const int g = __popc(mk);
  int consumed = -1;  // lanes up to here have been walked
  for (;;) {
    const bool c = lane > consumed && h >= st.next && h < st.niters && g > max(st.max_good, kModelPoints - 1);
    const unsigned cand = __ballot_sync(kFullMask, c);
    if (cand == 0u) break;
    const int first = __ffs(cand) - 1;
    accept(st, __shfl_sync(kFullMask, h, first), __shfl_sync(kFullMask, mk, first), n, confidence);
    consumed = first;
  }
}

// U(n, lo) and U(n, lo + W) for every point count: the distinct minimal sets FIRST drawn in [lo, lo + W) are the slots
// [begin[n], end[n]) of Model::d_uniq (host tables, ransac_model.cu)
struct PhaseSets {
  uint16_t begin[kMaxLandmarks + 1], end[kMaxLandmarks + 1];
};

// ---- phase p >= 1: hypotheses [lo, lo + W) of every frame in the phase's work list, 32 per warp item.
// Phases wider than one warp (p >= 2) evaluate every DISTINCT minimal set once, at its first draw: a later draw of the
// same five points (in whatever order) gives the same pose up to rounding, so its inlier count cannot be strictly
// larger than the first one's and cv2's loop never accepts it.  OpenCV's RNG repeats itself heavily (n = 11: 462 sets
// exist, the 10000 draws of a frame without a model contain each of them ~22 times), so a frame that stays in the loop
// costs U(n, 10000) evaluations instead of 10000.  The first 32 draws (phase 0 / 1) are evaluated as drawn.
__global__ void __launch_bounds__(kReplayWarps * 32, kReplayCtasPerSm) replay_phase_kernel(DevModel m, RansacArgs a, RansacWorkspace ws, int p, int lo, int W,
                                                                                           PhaseSets sets) {
  const int lane = threadIdx.x & 31;
  const float thr2 = a.reproj_err * a.reproj_err;
  const int nb = W / 32;  // blocks per frame (upper bound: fewer when draws repeat earlier sets)
  const int n_frames = (int)ws.claim[kClaimActive + p];
  const long long items = (long long)n_frames * nb;
  const int32_t* active = ws.x_active + (size_t)(p & 1) * ws.frames;
  const long long n_warps = (long long)gridDim.x * kReplayWarps;
  constexpr int kNever = 0x7fffffff;  // a lane without a hypothesis: beyond every budget
  for (long long item = (long long)blockIdx.x * kReplayWarps + (threadIdx.x >> 5); item < items; item += n_warps) {
    const int row = (int)(item / nb), blk = (int)(item - (long long)row * nb);
    const int b = active[row];
    const int n = ws.n[b];
    if (nb == 1) {
      // the whole phase of this frame sits in the warp: evaluate the draws themselves and walk them here
      ReplayState st = *state_of(ws, b);
      const int h = lo + lane;
      unsigned bits = 0;
      if (h >= st.next && h < st.niters && h < a.iterations) {
        const uint8_t* subset = m.subsets + ((size_t)(n - 6) * m.max_hyp + h) * kModelPoints;
        bits = hypothesis_f64(m.cam, frame_of(ws, b), n, subset, thr2);
      }
      accept_block(bits, h, lane, n, a.confidence, st);
      if (lane == 0) finish_or_continue(ws, b, st, lo + W, a.iterations, p + 1);
      continue;
    }
#ifdef SPE_REPLAY_ALL_DRAWS  // A/B builds (tools/ab_builds.py): every draw evaluated, as before the distinct-set phases
    const int u0 = lo, u1 = min(lo + W, a.iterations);
#else
    const int u0 = sets.begin[n], u1 = sets.end[n];
#endif
    const int nbf = max(1, (u1 - u0 + 31) >> 5);  // blocks this frame really has (>= 1: somebody has to walk it on)
    if (blk >= nbf) continue;
    const uint16_t* uniq = m.uniq + (size_t)(n - 6) * m.max_hyp;
    ReplayState st = *state_of(ws, b);
    const int u = u0 + blk * 32 + lane;
#ifdef SPE_REPLAY_ALL_DRAWS
    const int h = u < u1 ? u : kNever;
#else
    const int h = u < u1 ? (int)uniq[u] : kNever;  // the draw that introduces this set (ascending in u)
#endif
    unsigned bits = 0;
    if (h < st.niters) {  // (h >= lo >= st.next and h < iterations by construction)
      const uint8_t* subset = m.subsets + ((size_t)(n - 6) * m.max_hyp + h) * kModelPoints;
      bits = hypothesis_f64(m.cam, frame_of(ws, b), n, subset, thr2);
    }
    uint32_t* masks = ws.x_masks + (size_t)row * kReplayMaxWidth;
    masks[blk * 32 + lane] = bits;
    __threadfence();
    unsigned done = 0;
    if (lane == 0) done = atomicAdd(ws.x_done + b, 1u);
    done = __shfl_sync(kFullMask, done, 0);
    if (done == (unsigned)nbf - 1u) {  // this warp completed the frame's last block: walk the whole phase in order
      __threadfence();
      for (int c = 0; c < nbf; ++c) {
        const int uc = u0 + c * 32 + lane;
#ifdef SPE_REPLAY_ALL_DRAWS
        const int hc = uc < u1 ? uc : kNever;
#else
        const int hc = uc < u1 ? (int)uniq[uc] : kNever;
#endif
        if (__shfl_sync(kFullMask, hc, 0) >= st.niters) break;  // (uniform: st is replicated on the lanes, uniq ascends)
        accept_block(__ldcg(masks + c * 32 + lane), hc, lane, n, a.confidence, st);
      }
      if (lane == 0) {
        ws.x_done[b] = 0u;
        finish_or_continue(ws, b, st, lo + W, a.iterations, p + 1);
      }
    }
  }
}

__global__ void replay_reset_kernel(RansacWorkspace ws, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kClaimWords) ws.claim[i] = 0u;
  if (i < B) ws.x_done[i] = 0u;
}

}  // namespace

cudaError_t launch_ransac_replay(const Model& m, const RansacArgs& a, const RansacWorkspace& ws, cudaStream_t stream) {
  if (a.B == 0 || m.J <= kModelPoints) return cudaSuccess;
  static PerDeviceOnce once;
  cudaError_t e = once.run(m.device, [] {
    cudaError_t r = cudaFuncSetAttribute(replay_phase_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    if (r == cudaSuccess) r = cudaFuncSetAttribute(replay_eval_items_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, kSmemCarveoutPct);
    return r;
  });
  if (e != cudaSuccess) return e;
  int dev = 0, num_sms = 0;
  e = current_device(dev, num_sms);
  if (e != cudaSuccess) return e;
  const DevModel dm = dev_model(m);
  // counters and per-frame completion counts are reset by a kernel of the call itself (a memset node costs an engine
  // switch of ~40 us between two kernels of a stream)
  replay_reset_kernel<<<(std::max(a.B, kClaimWords) + 255) / 256, 256, 0, stream>>>(ws, a.B);
  replay_plan_kernel<<<(a.B + 3) / 4, 128, 0, stream>>>(dm, a, ws);
  // the item count lives on the device: a persistent grid sized for the most there can be
  const long long max_items = (long long)a.B * (a.H > 0 ? kReplayPlanMax : kPlanWidthNoScores);
  const int eval_ctas = (int)std::max<long long>(1, std::min<long long>(kReplayCtasPerSm * num_sms, (max_items + kReplayWarps * 32 - 1) / (kReplayWarps * 32)));
  replay_eval_items_kernel<<<eval_ctas, kReplayWarps * 32, 0, stream>>>(dm, a, ws);
  replay_scan_items_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a, ws);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // phase 1 starts at 0 again: a frame whose predicted prefix was shorter than cv2's real loop resumes at its own `next`
  int lo = 0, W = 32;
  for (int p = 1; lo < a.iterations && p < kReplayMaxPhases; ++p) {
    PhaseSets sets{};
    for (int n = kModelPoints + 1; n <= m.J; ++n) {
      sets.begin[n] = (uint16_t)unique_sets(m, n, lo);
      sets.end[n] = (uint16_t)unique_sets(m, n, std::min(lo + W, a.iterations));
    }
    replay_phase_kernel<<<kReplayCtasPerSm * num_sms, kReplayWarps * 32, 0, stream>>>(dm, a, ws, p, lo, W, sets);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    lo += W;
    W = std::min(W * 4, kReplayMaxWidth);
  }
  return cudaSuccess;
}

}  // namespace spe
