/* spe_b200 — C ABI of the B200-native heatmap -> 6-DoF pose stage.
 *
 * Drop-in boundary for ONE path of mohsij/spacecraft-pose-estimation.  The
 * reference has no FFI for this path; its boundary is three Python call sites.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference root):
 *
 *   spe_max_preds_f32        get_max_preds        landmark_regression/lib/core/inference.py:18-46
 *   spe_decode_f32           get_final_preds      landmark_regression/lib/core/inference.py:49-79
 *                            (+ transform_preds   landmark_regression/lib/utils/transforms.py:49-110)
 *   spe_ransac_epnp_f32      the per-frame loop   pose_estimation/export_predicted_poses_real.py:177-204
 *                            (confidence filter :186-197, cv2.solvePnPRansac :199-201,
 *                             cv2.Rodrigues :203, cv_rotation_matrix_to_quat :22-57)
 *   spe_pick_boxes_f32 / spe_boxes_to_center_scale_f64
 *                            detection box choice + _xywh2cs (object_detection/export_object_detection_bounding_boxes.py
 *                            :313-329, landmark_regression/lib/dataset/PEdataset.py:98-113)
 *   spe_pck_counts_f32       calc_dists + dist_acc behind accuracy()   landmark_regression/lib/core/evaluate.py:16-80
 *                            (the other half of accuracy() is get_max_preds = spe_max_preds_f32)
 *   spe_heatmap_to_pose_f32  the two above back to back with no host hop (the reference goes
 *                            through pred.mat: lib/dataset/PEdataset.py:121-123 ->
 *                            export_predicted_poses_real.py:172-173)
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer, C-contiguous;
 *     `stream` is a cudaStream_t passed as void* (NULL = default stream)
 *   - the caller owns every buffer; nothing is allocated per call (scratch comes from the
 *     caller-provided workspace, size from spe_ransac_workspace_bytes)
 *   - functions enqueue work on `stream` and return without synchronising; they are re-entrant
 *     and read no environment variables.  Work counters of the pose calls live in the caller's
 *     workspace and are reset by the first kernel of the call; the decode calls (which take no
 *     workspace) draw theirs from a per-(device, stream) slot table inside the library that every
 *     launch leaves clean.  A model handle is immutable after creation and may be shared by threads
 *     using the same device.  Calls that share a workspace must be ordered by the caller (same
 *     stream, or events).
 *   - return value: SPE_OK (0) or a negative spe_status code; never throws across the ABI.
 *     Per-frame conditions are reported in status[b] (SPE_FRAME_*), not as errors.
 */
#ifndef SPE_B200_H
#define SPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPE_ABI_VERSION 2
#define SPE_MAX_LANDMARKS 32 /* inlier masks are 32-bit */
#define SPE_MAX_HYPOTHESES 16384 /* >= the reference's iterationsCount = 10000 (export_predicted_poses_real.py:201) */

typedef enum spe_status {
  SPE_OK = 0,
  SPE_ERR_INVALID_ARGUMENT = -1, /* null pointer, non-positive size, J > SPE_MAX_LANDMARKS ... */
  SPE_ERR_CUDA = -2,             /* a CUDA runtime call failed; see spe_last_cuda_error() */
  SPE_ERR_WORKSPACE = -3,        /* workspace too small or misaligned */
  SPE_ERR_UNSUPPORTED = -4,       /* a configuration outside the reference's (camera skew, development-only flags ...) */
  SPE_ERR_OUT_OF_MEMORY = -5      /* host allocation failed while building a model */
} spe_status;

/* status[b] values written by spe_ransac_epnp_f32 (what cv2 would have done) */
typedef enum spe_frame_status {
  SPE_FRAME_OK = 0,
  SPE_FRAME_TOO_FEW_POINTS = 1, /* n < 4: cv2.solvePnPRansac raises cv2.error */
  SPE_FRAME_P3P_UNSUPPORTED = 2, /* (ABI 1; no longer produced: n == 4 is solved like cv2 does, with P3P on the four points) */
  SPE_FRAME_NO_MODEL = 3         /* no hypothesis reached 5 inliers (cv2 returns ret = False), or n == 4 and the P3P has
                                    no real solution (cv2 returns a NaN pose) */
} spe_frame_status;

typedef struct spe_model spe_model_t; /* opaque: landmarks, camera, per-n minimal-set tables */

int spe_abi_version(void);
const char* spe_status_string(int status);
/* cudaGetErrorString of the last CUDA failure seen by this thread inside the library */
const char* spe_last_cuda_error(void);

/* ---- decode -------------------------------------------------------------------------------
 * hm        [B,J,H,W] float32   raw HRNet output
 * preds     [B,J,2]   float32   (x, y): bit-equal to the reference in 99.99 % of the values; the rest differ by one
 *                               float32 ulp (<= 6.1e-5 px below 1024 px, 1.22e-4 px above), the residue of
 *                               cv2.getAffineTransform's LU inside transform_preds
 * maxvals   [B,J]     float32   (the reference's [B,J,1])
 * argmax    [B,J]     int32     flat index of the maximum (first on ties, first NaN wins); may be NULL
 */
int spe_max_preds_f32(const float* hm, int B, int J, int H, int W, float* preds, float* maxvals,
                      int32_t* argmax, void* stream);

/* center, scale [B,2] float32 (scale[:,1] is ignored, as in the reference); post_process =
 * config.TEST.POST_PROCESS.  preds are image pixels after the inverse box affine. */
int spe_decode_f32(const float* hm, int B, int J, int H, int W, const float* center,
                   const float* scale, int post_process, float* preds, float* maxvals,
                   int32_t* argmax, void* stream);

/* Same as spe_decode_f32 but writes the pred.mat row layout directly:
 * kpts [B,J,3] float32 = (x, y, maxval) (lib/core/function.py:392-393). */
int spe_decode_kpts_f32(const float* hm, int B, int J, int H, int W, const float* center,
                        const float* scale, int post_process, float* kpts, int32_t* argmax,
                        void* stream);

/* spe_decode_kpts_f32 with flags.  SPE_DECODE_BACKGROUND: the caller runs this decode on its own stream UNDER a
 * compute-bound kernel of another stream (the hypothesis scoring of the previous batch): it is launched as small CTAs
 * (3 warps, 51 KB of shared memory) that take the place of one hypothesis CTA per SM as those retire, instead of the
 * 7-warp / 113 KB CTAs that want an SM's shared memory to themselves.  Same results. */
#define SPE_DECODE_BACKGROUND 1
int spe_decode_kpts_ex_f32(const float* hm, int B, int J, int H, int W, const float* center,
                           const float* scale, int post_process, float* kpts, int32_t* argmax,
                           int flags, void* stream);

/* Decode of a combination of K heatmap tensors that is never written to memory.
 * srcs: HOST array of K DEVICE pointers, each [B,J,H,W] float32.
 *   SPE_COMBINE_MEAN  ((src0 + src1) + ... ) * (1/K)          model ensemble, validate_cv,
 *                     landmark_regression/lib/core/function.py:525-536 (torch's tensor/scalar on CUDA)
 *   SPE_COMBINE_FLIP  K = 2: (src0 + shift(flip_back(src1))) * 0.5   flip test, function.py:347-366;
 *                     flip_back = lib/utils/transforms.py:15-29.  flip_perm: DEVICE int32 [J], the joint
 *                     whose flipped map lands on joint j (NULL = no matched pairs, as in this dataset);
 *                     shift_heatmap = config.TEST.SHIFT_HEATMAP.
 * Output as spe_decode_kpts_f32. */
#define SPE_COMBINE_MEAN 0
#define SPE_COMBINE_FLIP 1
int spe_decode_combined_kpts_f32(const float* const* srcs, int K, int mode, const int32_t* flip_perm,
                                 int shift_heatmap, int B, int J, int H, int W, const float* center,
                                 const float* scale, int post_process, float* kpts, int32_t* argmax,
                                 void* stream);

/* ---- detection boxes -> (center, scale) -----------------------------------------
 * spe_boxes_to_center_scale_f64   JointsDataset._xywh2cs, landmark_regression/lib/dataset/PEdataset.py:98-113
 *                                 (pixel_std = 200, :41): xywh [B,4] float64 COCO boxes (the annotation's 'bbox') ->
 *                                 center [B,2], scale [B,2] float32 = the meta['center'] / meta['scale'] the decode takes.
 * spe_pick_boxes_f32              the per-image box choice of object_detection/export_object_detection_bounding_boxes.py
 *                                 :313-329 followed by _xywh2cs: boxes [B,K,4] float32 (x1,y1,x2,y2), scores [B,K]
 *                                 float32, counts [B] int32 = detections per image (NULL: K everywhere).  An image
 *                                 with 1 or 2 detections keeps the best-scoring one (first maximum, a NaN wins), any
 *                                 other count takes the whole image [0,0,image_w,image_h] with score 0.
 *                                 Outputs (each may be NULL; center and scale only together): xywh [B,4] float64 (the
 *                                 annotation's 'bbox'), best_score [B] float32, best_index [B] int32 (-1: whole image),
 *                                 center/scale [B,2] float32. */
int spe_boxes_to_center_scale_f64(const double* xywh, int B, float* center, float* scale, void* stream);
int spe_pick_boxes_f32(const float* boxes, const float* scores, const int32_t* counts, int B, int K, double image_w,
                       double image_h, double* xywh, float* best_score, int32_t* best_index, float* center,
                       float* scale, void* stream);

/* ---- accuracy() ------------------------------------------------------------------
 * The counting half of core.evaluate.accuracy (landmark_regression/lib/core/evaluate.py:16-39: calc_dists, dist_acc) for
 * the training / validation loops (lib/core/function.py:61-62, :395-396).  pred, target: [B,J,2] float32 argmax
 * coordinates of the network output and of the target heatmaps (spe_max_preds_f32).  A (frame, joint) is counted when both
 * target coordinates are > 1; its distance is |pred / norm - target / norm| in float64 with norm = (norm_x, norm_y) applied
 * to (x, y) — the reference passes (H / 10, W / 10), in that order.  counts [J,2] int32 DEVICE (overwritten):
 * per joint the number of counted frames and how many of them lie below `thr` (the reference: always 0.5, its own `thr`
 * argument is not passed on, evaluate.py:71). */
int spe_pck_counts_f32(const float* pred, const float* target, int B, int J, double norm_x, double norm_y, double thr,
                       int32_t* counts, void* stream);

/* ---- pose ---------------------------------------------------------------------------------
 * landmarks [J,3] float64 HOST (metres; rounded to float32 internally exactly like cv2 does),
 * K[9] row-major float64 HOST (pinhole without skew; anything else: SPE_ERR_UNSUPPORTED),
 * dist[5] = (k1,k2,p1,p2,k3) float64 HOST (NULL = no distortion).
 * max_hypotheses = cv2's iterationsCount (10000 in the reference, export_predicted_poses_real.py:201): it
 * bounds `hypotheses` of later calls and is where the iteration budget of cv2's loop starts.  Builds, for every
 * point count 6..J on the current device, the minimal sets OpenCV's fixed-seed RNG draws, the list of DISTINCT
 * sets among them (a repeated 5-subset is scored once) and the control-point table of every 5-subset. */
int spe_pnp_model_create(const double* landmarks, int J, const double* K, const double* dist,
                         int max_hypotheses, spe_model_t** out);
int spe_pnp_model_destroy(spe_model_t* model);
int spe_pnp_model_num_landmarks(const spe_model_t* model);
/* Host-only (no CUDA call): one entry of the per-model control-point table that spe_pnp_model_create builds for every
 * 5-subset of the landmarks and the hypothesis kernel looks up (EPnP's control points depend on the object points only,
 * OpenCV epnp.cpp choose_control_points / compute_barycentric_coordinates).
 * ids: five landmark indices in ascending order.  entry [20] float32: alpha[5][3] (barycentric coordinates with respect
 * to the three PCA control points; the one of the centroid is 1 - their sum), then the three squared distances of those
 * control points from the centroid, then two zeros.  rank (optional): position of the subset in the table. */
int spe_pnp_control_entry(const double* landmarks, int J, const int32_t* ids, float* entry, int64_t* rank);

/* Host-only (no CUDA call, no model): the first `count` minimal sets OpenCV's RANSAC draws for n points (sets
 * [count*5], draw order), the distinct-set number of every draw (slot [count]) and the draw that introduces each
 * distinct set (uniq [count], -1 beyond *num_unique).  What spe_pnp_model_create uploads; any output may be NULL. */
int spe_pnp_minimal_sets_host(int n, int count, int32_t* sets, int32_t* slot, int32_t* uniq, int32_t* num_unique);

/* copies the first `count` minimal sets for n points into out[count*5] (HOST), draw order kept */
int spe_pnp_model_minimal_sets(const spe_model_t* model, int n, int count, int32_t* out);

size_t spe_ransac_workspace_bytes(const spe_model_t* model, int B, int hypotheses);

/* kpts        [B,J,3] float32   (x, y, conf) rows as stored in pred.mat
 * hypotheses  number of minimal sets scored per frame by the FP32 hypothesis kernel (every distinct set once);
 *             may be 0 with SPE_FLAG_EXACT (no FP32 scoring at all)
 * reproj_err  15.0 in the reference; confidence 0.99 (cv2 default)
 * conf_floor  a landmark takes part iff conf > conf_floor; pass a negative value to run the
 *             reference's adaptive 0.95*0.8^k filter per frame (export_predicted_poses_real.py:186-197)
 * pose7       [B,7] float32 = (qw,qx,qy,qz,tx,ty,tz).  float32 resolves a rotation to ~1e-3 deg (less near pi)
 *             and a translation to 6e-8 relative: parity-grade consumers read `rt`
 * inlier_mask [B] uint32 over the J landmarks (bit j); status [B] int32 (spe_frame_status)
 * winner_hyp  [B] int32 index of the accepted hypothesis (-1 none); may be NULL
 * rt          [B,12] float64 = row-major R (9) then t (3) of the final refit; may be NULL
 *
 * Which hypothesis wins (flags):
 *   default         cv2's sequential acceptance rule replayed over the FP32 inlier counts of the first `hypotheses`
 *                   draws.  Equals cv2's result whenever FP32 and cv2's float64 agree on the hypotheses cv2 looks at
 *                   and cv2 stops within `hypotheses` draws; spe_ransac_read_budget tells when it would not have.
 *   SPE_FLAG_EXACT  cv2's loop itself, replayed in float64 up to the model's max_hypotheses (iterationsCount):
 *                   the hypotheses cv2 looks at (6.5 per frame on the benchmark data) are re-evaluated in float64,
 *                   whatever the FP32 scores say.  This is the parity path.  Beyond the first 32 draws a minimal set
 *                   that repeats an earlier draw (OpenCV's RNG does so heavily: 462 distinct sets among 10000 draws
 *                   at 11 points) is not evaluated again: the same five points give the same pose up to rounding,
 *                   and cv2's loop only accepts a strictly larger inlier count.
 */
#define SPE_FLAG_REFINE_LM 1
/* SPE_FLAG_ADAPTIVE: score the first 32 minimal sets of every frame, replay cv2's acceptance loop
 * over them, then score only the hypotheses cv2's (shrinking) iteration budget could still reach.
 * The result is identical to scoring all `hypotheses` (the selection never reads the skipped
 * entries); the work is what cv2 itself would do, rounded up to blocks of 32.  Off by default. */
#define SPE_FLAG_ADAPTIVE 2
/* SPE_FLAG_BACKGROUND_TAIL (spe_ransac_select_refit_f32): the caller overlaps this call with the
 * scoring of the next batch on another stream; the refit is then launched as whole-SM CTAs on few
 * SMs instead of one warp on every SM. */
#define SPE_FLAG_BACKGROUND_TAIL 4
/* SPE_FLAG_JACOBI_SVD (scoring): take EPnP's four vectors from a full one-sided Jacobi SVD of M^T
 * (the reference algorithm's eigensolve, 2.3x slower) instead of the default Householder QR + block
 * inverse iteration.  Same results up to the chaos of near-degenerate hypotheses; a development variant
 * for A/B measurements (tests/test_pnp_gpu.py compares the two when the library was built with it). */
#define SPE_FLAG_JACOBI_SVD 8 /* development builds (-DSPE_DEV) only; SPE_ERR_UNSUPPORTED otherwise */
#define SPE_FLAG_EXACT 16
int spe_ransac_epnp_f32(const spe_model_t* model, const float* kpts, int B, int hypotheses,
                        float reproj_err, double confidence, float conf_floor, float* pose7,
                        uint32_t* inlier_mask, int32_t* status, int32_t* winner_hyp, double* rt,
                        void* workspace, size_t workspace_bytes, int flags, void* stream);

/* The two halves of spe_ransac_epnp_f32, for callers that pipeline batches: the first scores every
 * hypothesis (frame preparation + FP32 hypothesis kernel, throughput-bound), the second replays
 * cv2's sequential acceptance and runs the float64 refit (latency-bound: ~0.3 ms whatever B).
 * Issued on different streams (with an event in between) the second half of batch i overlaps
 * the first half of batch i+1.  Both must see the same workspace, B and hypotheses. */
int spe_ransac_score_f32(const spe_model_t* model, const float* kpts, int B, int hypotheses,
                         float reproj_err, double confidence, float conf_floor, void* workspace,
                         size_t workspace_bytes, int flags, void* stream);
int spe_ransac_select_refit_f32(const spe_model_t* model, int B, int hypotheses, double confidence,
                                float* pose7, uint32_t* inlier_mask, int32_t* status,
                                int32_t* winner_hyp, double* rt, void* workspace,
                                size_t workspace_bytes, int flags, void* stream);

/* SPE_FLAG_EXACT as its own stage, between spe_ransac_score_f32 and spe_ransac_select_refit_f32 (which must then be
 * given SPE_FLAG_EXACT too): the float64 replay of cv2's loop for every frame of the workspace.  It needs the frame
 * preparation of spe_ransac_score_f32 (which may have been called with hypotheses = 0) and the same B / hypotheses. */
int spe_ransac_replay_f64(const spe_model_t* model, int B, int hypotheses, float reproj_err, double confidence,
                          void* workspace, size_t workspace_bytes, void* stream);

/* budget [B] int32 DEVICE, valid after spe_ransac_select_refit_f32 / spe_ransac_epnp_f32 on this workspace: cv2's
 * iteration budget when its loop ends = the number of hypotheses it looks at.  Without SPE_FLAG_EXACT a value above
 * `hypotheses` means cv2 would have kept drawing beyond what was scored (re-run those frames with SPE_FLAG_EXACT). */
int spe_ransac_read_budget(const spe_model_t* model, const void* workspace, int B, int hypotheses, int32_t* budget,
                           void* stream);

/* Per-hypothesis FP32 scores of the most recent spe_ransac_epnp_f32 call on this workspace, for the
 * parity tests: counts [B,hypotheses] int32 and masks [B,hypotheses] uint32 (DEVICE), in draw order. */
int spe_ransac_debug_scores(const spe_model_t* model, const void* workspace, int B, int hypotheses,
                            int32_t* counts, uint32_t* masks, void* stream);

/* decode + pose with the keypoints kept in HBM; kpts_out [B,J,3] may be NULL only if workspace
 * has room (it always does: spe_pipeline_workspace_bytes accounts for it). */
size_t spe_pipeline_workspace_bytes(const spe_model_t* model, int B, int J, int hypotheses);
int spe_heatmap_to_pose_f32(const spe_model_t* model, const float* hm, int B, int J, int H, int W,
                            const float* center, const float* scale, int post_process,
                            int hypotheses, float reproj_err, double confidence, float conf_floor,
                            float* pose7, uint32_t* inlier_mask, int32_t* status, float* kpts_out,
                            void* workspace, size_t workspace_bytes, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPE_B200_H */
