"""Oracle (test infrastructure): the PCK-style training/validation metric built on get_max_preds, restated.

`accuracy(output, target, hm_type, thr)` — landmark_regression/lib/core/evaluate.py:16-80 (`calc_dists`, `dist_acc`,
`accuracy`), as called from lib/core/function.py:61-62 / :395-396 with the network output and the target heatmaps:
argmax coordinates of both ([B,J,2] float32), per-joint distances in float64 after dividing (x, y) by (H, W) / 10
— the reference's order, kept — with -1 for a joint whose target coordinates are not both > 1, then per joint the share
of valid distances below `thr`, and their mean over the joints that have one.

Pinned by tests/golden/accuracy_golden.npz, produced by importing the reference's own core.evaluate
(tests/golden/make_accuracy_golden.py).  Vectorised over the two reference loops; same float arithmetic per element.
"""
from __future__ import annotations

import numpy as np

from . import decode_ref


def calc_dists(preds, target, normalize):
    """preds, target [B,J,2]; normalize [B,2] float64 -> dists [J,B] float64 (evaluate.py:16-29)."""
    preds = preds.astype(np.float32)
    target = target.astype(np.float32)
    valid = (target[:, :, 0] > 1) & (target[:, :, 1] > 1)
    d = preds / normalize[:, None, :] - target / normalize[:, None, :]  # float32 / float64 -> float64, as in the loop
    dist = np.sqrt(d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1])  # np.linalg.norm of a 2-vector
    return np.where(valid, dist, -1.0).T


def dist_acc(dists, thr=0.5):
    """evaluate.py:32-39."""
    dist_cal = np.not_equal(dists, -1)
    num_dist_cal = dist_cal.sum()
    if num_dist_cal > 0:
        return np.less(dists[dist_cal], thr).sum() * 1.0 / num_dist_cal
    return -1


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """evaluate.py:42-80 -> (acc [J+1] float64, avg_acc, cnt, pred [B,J,2] float32)."""
    J = output.shape[1]
    norm = 1.0
    if hm_type == "gaussian":
        pred, _ = decode_ref.get_max_preds(output)
        target, _ = decode_ref.get_max_preds(target)
        h, w = output.shape[2], output.shape[3]
        norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    else:
        raise ValueError("the reference defines `pred` for hm_type == 'gaussian' only")
    dists = calc_dists(pred, target, norm)
    acc = np.zeros((J + 1))
    avg_acc, cnt = 0, 0
    for i in range(J):
        acc[i + 1] = dist_acc(dists[i])  # evaluate.py:71: `thr` is NOT passed on — dist_acc always runs at its default 0.5
        if acc[i + 1] >= 0:
            avg_acc = avg_acc + acc[i + 1]
            cnt += 1
    avg_acc = avg_acc / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg_acc
    return acc, avg_acc, cnt, pred
