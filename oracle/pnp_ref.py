"""Oracle (test infrastructure): the reference's per-frame pose solve on the CPU.

Restates pose_estimation/export_predicted_poses_real.py:177-226 without its file and image I/O:
confidence filter (:186-197), cv2.solvePnPRansac with the script's arguments (:199-201),
cv2.Rodrigues (:203) and the quaternion helper (:22-57).

`solve_pnp_ransac_cv2` is the black box (OpenCV itself).  `ransac_epnp_whitebox` is OpenCV's
RANSAC loop restated around cv2's own solvePnP/projectPoints primitives so that intermediates
(per-hypothesis inlier counts and masks, the winner, the shrinking budget) are visible; it is
asserted bit-identical to the black box in tests/test_oracle_pnp.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import cv2
import numpy as np

from . import ocv_rng
from .epnp_ref import rotation_matrix_to_quat

REPROJECTION_ERROR = 15.0  # export_predicted_poses_real.py:201
ITERATIONS_COUNT = 10000  # export_predicted_poses_real.py:201
CONFIDENCE = 0.99  # cv2 default, not overridden by the script


def confidence_filter(conf: np.ndarray) -> np.ndarray:
    """export_predicted_poses_real.py:186-197: threshold 0.95, multiplied by 0.8 while fewer than
    15 landmarks pass, at most 100 times.  For J <= 14 it always ends at 0.95*0.8**100."""
    conf = np.asarray(conf, np.float32)
    thr = 0.95
    good = conf > thr
    rounds = 0
    while np.sum(good) < 15:
        thr *= 0.8
        good = conf > thr
        rounds += 1
        if rounds >= 100:
            break
    return good


def confidence_floor(num_landmarks: int) -> float:
    """The threshold the filter ends at when it can never find 15 landmarks (J < 15)."""
    thr = 0.95
    for _ in range(100):
        thr *= 0.8
    return thr if num_landmarks < 15 else float("nan")


def solve_pnp_ransac_cv2(obj, img, K, dist, iterations=ITERATIONS_COUNT, reproj=REPROJECTION_ERROR):
    """The reference's call, verbatim in its arguments (:199-201).  obj float64 [n,3] (as read by
    pandas), img float32 [n,2].  Returns (ok, rvec[3], tvec[3], inliers int32[k] or None)."""
    ok, rvec, tvec, inl = cv2.solvePnPRansac(
        np.ascontiguousarray(obj, np.float64), np.ascontiguousarray(img, np.float32), K,
        distCoeffs=dist, flags=cv2.SOLVEPNP_EPNP, iterationsCount=int(iterations), reprojectionError=float(reproj))
    return bool(ok), np.asarray(rvec, np.float64).ravel(), np.asarray(tvec, np.float64).ravel(), (None if inl is None else inl.ravel())


@dataclass
class RansacTrace:
    ok: bool
    rvec: np.ndarray
    tvec: np.ndarray
    inliers: np.ndarray | None
    counts: np.ndarray  # [H_eval] inlier count of every hypothesis evaluated (all H if exhaustive)
    masks: np.ndarray  # [H_eval] uint32 bit mask over the n input points
    winner: int  # index of the accepted hypothesis, -1 if none
    evaluated: int  # hypotheses cv2 would have evaluated before its budget ran out


def hypothesis_masks(obj32, img32, K, dist, subsets, reproj=REPROJECTION_ERROR):
    """Score every minimal set the way cv2's PnPRansacCallback does (App. B.3, B.5): EPnP on the 5
    points, project all n points (float32 object points, float32 output), squared error in
    float32, inlier <=> err <= reproj^2."""
    n = obj32.shape[0]
    thr = np.float32(reproj * reproj)
    counts = np.zeros(len(subsets), np.int32)
    masks = np.zeros(len(subsets), np.uint32)
    poses = np.zeros((len(subsets), 6))
    for h, s in enumerate(subsets):
        ok, rv, tv = cv2.solvePnP(obj32[s], img32[s], K, dist, flags=cv2.SOLVEPNP_EPNP)
        proj, _ = cv2.projectPoints(obj32, rv, tv, K, dist)
        d = img32 - proj.reshape(n, 2).astype(np.float32)
        err = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
        good = err <= thr
        counts[h] = int(good.sum())
        masks[h] = np.uint32(sum(1 << i for i in range(n) if good[i]))
        poses[h, :3], poses[h, 3:] = rv.ravel(), tv.ravel()
    return counts, masks, poses


def ransac_epnp_whitebox(obj, img, K, dist, iterations=ITERATIONS_COUNT, reproj=REPROJECTION_ERROR,
                         confidence=CONFIDENCE, exhaustive: int = 0) -> RansacTrace:
    """OpenCV's solvePnPRansac(EPNP) restated (App. B.1, B.6, B.7).

    `exhaustive=H` additionally scores the first H minimal sets even past cv2's early exit, which
    is what the GPU does; the winner/budget are still the sequential ones.
    """
    obj32 = np.ascontiguousarray(obj, np.float32)  # cv2 converts object points to float32 on entry
    img32 = np.ascontiguousarray(img, np.float32)
    n = obj32.shape[0]
    if n < 4:
        raise ValueError("solvePnPRansac needs at least 4 points")
    if n == 4:
        raise NotImplementedError("n == 4 takes OpenCV's P3P kernel (out of scope, SURVEY App. B.1)")
    if n == 5:  # model_points == npoints: plain solvePnP, all points inliers
        ok, rv, tv = cv2.solvePnP(obj32, img32, K, dist, flags=cv2.SOLVEPNP_EPNP)
        return RansacTrace(bool(ok), rv.ravel(), tv.ravel(), np.arange(5, dtype=np.int32), np.array([5]),
                           np.array([31], np.uint32), 0, 1)
    # sequential loop: evaluate lazily until the (shrinking) budget is exhausted
    cap = min(int(iterations), 100000)
    counts, masks = [], []
    niters, best, max_good, h = cap, -1, 0, 0
    block = 32
    subsets = ocv_rng.minimal_sets(n, min(cap, block))
    while h < niters:
        if h >= len(subsets):
            subsets = ocv_rng.minimal_sets(n, min(cap, len(subsets) * 2))
        c, m, _ = hypothesis_masks(obj32, img32, K, dist, subsets[h:h + 1], reproj)
        counts.append(int(c[0]))
        masks.append(int(m[0]))
        if c[0] > max(max_good, 4):
            best, max_good = h, int(c[0])
            niters = ocv_rng.update_num_iters(confidence, (n - max_good) / n, 5, niters)
        h += 1
    evaluated = h
    if exhaustive > evaluated:
        subsets = ocv_rng.minimal_sets(n, exhaustive)
        c, m, _ = hypothesis_masks(obj32, img32, K, dist, subsets[evaluated:exhaustive], reproj)
        counts += [int(v) for v in c]
        masks += [int(v) for v in m]
    counts = np.array(counts, np.int32)
    masks = np.array(masks, np.uint32)
    if best < 0:
        return RansacTrace(False, np.zeros(3), np.zeros(3), None, counts, masks, -1, evaluated)
    inl = np.array([i for i in range(n) if (int(masks[best]) >> i) & 1], np.int32)
    ok, rv, tv = cv2.solvePnP(obj32[inl].astype(np.float64), img32[inl].astype(np.float64), K, dist, flags=cv2.SOLVEPNP_EPNP)
    return RansacTrace(bool(ok), rv.ravel(), tv.ravel(), inl, counts, masks, best, evaluated)


def pose_from_keypoints(kpts, landmarks, K, dist, iterations=ITERATIONS_COUNT, reproj=REPROJECTION_ERROR):
    """One frame of the reference's loop (:180-204): kpts [J,3] = (x, y, conf) as stored in
    pred.mat.  Returns (ok, pose7 = [qw,qx,qy,qz,tx,ty,tz], inlier mask over the J landmarks,
    rvec, tvec)."""
    kpts = np.asarray(kpts).reshape(-1, 3)
    img = kpts[:, :2].astype(np.float32)
    good = confidence_filter(kpts[:, 2])
    ok, rvec, tvec, inl = solve_pnp_ransac_cv2(np.asarray(landmarks, np.float64)[good], img[good], K, dist, iterations, reproj)
    R, _ = cv2.Rodrigues(rvec)
    q = rotation_matrix_to_quat(R)
    mask = 0
    if inl is not None:
        idx = np.flatnonzero(good)
        for i in inl:
            mask |= 1 << int(idx[i])
    return ok, np.concatenate([q, tvec]), mask, rvec, tvec


def refine_lm_cv2(obj, img, K, dist, rvec, tvec, inliers):
    """Optional extra: cv2.solvePnPRefineLM on the RANSAC inliers.  The reference's
    own call has no such step (flags=SOLVEPNP_EPNP ends with EPnP on the inliers)."""
    o = np.ascontiguousarray(np.asarray(obj, np.float64)[inliers])
    i = np.ascontiguousarray(np.asarray(img, np.float32)[inliers])
    r, t = cv2.solvePnPRefineLM(o, i, K, dist, np.asarray(rvec, np.float64).reshape(3, 1).copy(), np.asarray(tvec, np.float64).reshape(3, 1).copy())
    return r.ravel(), t.ravel()


# ----------------------------------------------------------------------------- parity metrics
def rotation_angle_deg(Ra: np.ndarray, Rb: np.ndarray) -> float:
    """Geodesic angle between two rotations via atan2(|vee|, (tr-1)/2): resolves 1e-7 deg where
    arccos((tr-1)/2) floors at ~1e-2 deg on float32 matrices."""
    D = np.asarray(Ra, np.float64).T @ np.asarray(Rb, np.float64)
    w = 0.5 * np.array([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])
    return float(np.degrees(np.arctan2(np.linalg.norm(w), (np.trace(D) - 1.0) / 2.0)))


def quat_to_matrix(q: np.ndarray) -> np.ndarray:
    w, x, y, z = (float(v) for v in np.asarray(q, np.float64) / np.linalg.norm(q))
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def pose_errors(pose_a: np.ndarray, pose_b: np.ndarray):
    """(rotation error in degrees, relative translation error) between two [7] poses."""
    ra, rb = quat_to_matrix(pose_a[:4]), quat_to_matrix(pose_b[:4])
    ta, tb = np.asarray(pose_a[4:], np.float64), np.asarray(pose_b[4:], np.float64)
    return rotation_angle_deg(ra, rb), float(np.linalg.norm(ta - tb) / max(np.linalg.norm(tb), 1e-30))


# ----------------------------------------------------------------------------- float64 white box without cv2
def ransac_epnp_numpy(obj, img, K, dist, iterations=ITERATIONS_COUNT, reproj=REPROJECTION_ERROR, confidence=CONFIDENCE):
    """cv2.solvePnPRansac(EPNP) restated end to end in NumPy float64 — OpenCV's RNG and loop (ocv_rng), this package's
    own EPnP (epnp_ref, incl. the port of OpenCV's Jacobi SVD), undistortion and projection; no cv2 call.  It is the
    "float64 white box" of the parity tests: an independent float64 implementation of the same algorithm.  Where it
    and cv2 disagree on a frame's inlier set, the frame's answer depends on rounding noise (the 2-D null space of
    5-point EPnP).  Returns (ok, R, t, inlier indices or None, winner, hypotheses looked at)."""
    from . import epnp_ref

    obj32 = np.ascontiguousarray(obj, np.float32)
    img32 = np.ascontiguousarray(img, np.float32)
    n = obj32.shape[0]
    if n < 6:
        raise ValueError("the white box covers the RANSAC branch (n >= 6)")
    thr = np.float32(reproj * reproj)
    und32 = epnp_ref.undistort_points(img32, K, dist)  # float32 out, as cv2 does for float32 input
    niters, best, max_good, h, best_good = int(iterations), -1, 0, 0, None
    subsets = ocv_rng.minimal_sets(n, min(int(iterations), 64))
    while h < niters:
        if h >= len(subsets):
            subsets = ocv_rng.minimal_sets(n, min(int(iterations), len(subsets) * 4))
        s = subsets[h]
        R, t = epnp_ref.epnp(obj32[s].astype(np.float64), und32[s].astype(np.float64), K)
        proj = epnp_ref.project_points(obj32.astype(np.float64), R, t, K, dist).astype(np.float32)
        d = img32 - proj
        with np.errstate(over="ignore", invalid="ignore"):
            good = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) <= thr
        c = int(good.sum())
        if c > max(max_good, 4):
            best, max_good, best_good = h, c, good.copy()
            niters = ocv_rng.update_num_iters(confidence, (n - c) / n, 5, niters)
        h += 1
    if best < 0:
        return False, None, None, None, -1, h
    inl = np.flatnonzero(best_good).astype(np.int32)
    und64 = epnp_ref.undistort_points(img32[inl].astype(np.float64), K, dist)
    R, t = epnp_ref.epnp(obj32[inl].astype(np.float64), und64, K)
    return True, R, t, inl, best, h


def cv2_is_unstable(obj, img, K, dist, trials: int = 8, seed: int = 0, iterations=ITERATIONS_COUNT, reproj=REPROJECTION_ERROR):
    """Does cv2.solvePnPRansac itself return a different inlier set when every image coordinate is moved by at most one
    float32 ulp (~1e-4 px: nothing a measurement could resolve)?  Then the frame's answer is decided by rounding noise
    inside OpenCV, not by the data: no second implementation — nor cv2 on another CPU — can be expected to reproduce it."""
    img32 = np.ascontiguousarray(img, np.float32)
    ok0, _, _, inl0 = solve_pnp_ransac_cv2(obj, img32, K, dist, iterations, reproj)
    ref = None if inl0 is None else tuple(int(i) for i in inl0)
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        step = rng.integers(-1, 2, img32.shape)
        p = np.where(step > 0, np.nextafter(img32, np.float32(np.inf)), np.where(step < 0, np.nextafter(img32, np.float32(-np.inf)), img32))
        ok, _, _, inl = solve_pnp_ransac_cv2(obj, p.astype(np.float32), K, dist, iterations, reproj)
        got = None if inl is None else tuple(int(i) for i in inl)
        if ok != ok0 or got != ref:
            return True
    return False
