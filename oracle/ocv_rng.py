"""Oracle (test infrastructure): OpenCV's RANSAC random-number stream and iteration budget.

OpenCV is a third-party dependency of the reference (opencv-python==3.4.11.41, environment.yml:37;
cv2 4.13.0 in this image) and is not vendored under /root/reference, so its published algorithm is
restated here and pinned against cv2 itself by
tests/test_oracle_pnp.py (bit-identical rvec/tvec/inliers) and the known answers of App. E.2.

cv::RNG is a multiply-with-carry generator; calib3d's RANSACPointSetRegistrator seeds it with
(uint64)-1 on every run, so the 5-point minimal sets used by cv2.solvePnPRansac are a pure
function of the number of points.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

_MASK64 = (1 << 64) - 1
_CV_RNG_COEFF = 4164903690
_DBL_MIN = 2.2250738585072014e-308


class CvRNG:
    def __init__(self, state: int = _MASK64):
        self.state = state

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * _CV_RNG_COEFF + (self.state >> 32)) & _MASK64
        return self.state & 0xFFFFFFFF

    def uniform(self, lo: int, hi: int) -> int:
        return lo if lo == hi else self.next() % (hi - lo) + lo


@lru_cache(maxsize=None)
def _subsets_cached(count: int, model_points: int, num: int):
    rng = CvRNG()
    out = np.empty((num, model_points), np.int32)
    for h in range(num):
        chosen = []
        for _ in range(model_points):
            v = rng.uniform(0, count)
            while v in chosen:  # redraw duplicates within the subset; RNG state carries on
                v = rng.uniform(0, count)
            chosen.append(v)
        out[h] = chosen
    out.setflags(write=False)
    return out


def minimal_sets(count: int, num: int, model_points: int = 5) -> np.ndarray:
    """First `num` minimal sets (draw order preserved) cv2.solvePnPRansac uses for `count` points."""
    return _subsets_cached(int(count), int(model_points), int(num))


def update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    """RANSACUpdateNumIters: budget after a model with outlier ratio `ep` (App. B.6)."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, _DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < _DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def select_sequential(counts, n_points: int, iterations: int, confidence: float = 0.99, model_points: int = 5):
    """Replay of the RANSAC acceptance loop over per-hypothesis inlier counts (App. B.6).

    Returns (best_index or -1, hypotheses_evaluated).  A hypothesis wins only if its count is
    strictly greater than max(best_so_far, model_points - 1); each win shrinks the budget.
    """
    niters = min(int(iterations), len(counts))
    best, max_good, h = -1, 0, 0
    while h < niters:
        g = int(counts[h])
        if g > max(max_good, model_points - 1):
            best, max_good = h, g
            niters = min(niters, update_num_iters(confidence, (n_points - g) / n_points, model_points, niters))
        h += 1
    return best, h
