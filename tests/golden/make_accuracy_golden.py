"""Golden vectors for accuracy(), produced by IMPORTING THE REFERENCE here.

Run in the build container only (needs /root/reference):
    python tests/golden/make_accuracy_golden.py
Writes tests/golden/accuracy_golden.npz: seeded (output, target) heatmap pairs and what the reference's own
core.evaluate.accuracy (landmark_regression/lib/core/evaluate.py:42-80) returns for them.  Cases cover joints whose target
peak lies at x <= 1 or y <= 1 (distance -1: not counted), a joint invalid in every frame (acc -1), an all-invalid batch
(cnt = 0), non-square maps, and a `thr` argument other than 0.5 (which the reference does not pass on to dist_acc).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/landmark_regression/lib")

from core.evaluate import accuracy  # noqa: E402  (the reference)


def peaks_to_heatmaps(rng, xy, H, W, sigma=1.5, noise=0.02):
    """[B,J,2] integer peak positions -> [B,J,H,W] float32 Gaussian blobs on a little noise."""
    B, J, _ = xy.shape
    yy, xx = np.mgrid[0:H, 0:W]
    hm = rng.normal(scale=noise, size=(B, J, H, W))
    for b in range(B):
        for j in range(J):
            hm[b, j] += np.exp(-((xx - xy[b, j, 0]) ** 2 + (yy - xy[b, j, 1]) ** 2) / (2 * sigma ** 2))
    return (np.rint(hm * 1024.0) / 1024.0).astype(np.float32)  # multiples of 2^-10: exact in float32, and the file stays small


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    cases = []
    for name, (B, J, H, W) in (("square32", (8, 11, 32, 32)), ("hubble48x36", (4, 17, 48, 36)), ("tiny", (3, 4, 9, 7))):
        tgt = np.stack([rng.integers(0, W, (B, J)), rng.integers(0, H, (B, J))], -1)
        tgt[:, 0, 0] = rng.integers(0, 2, B)  # joint 0: target x in {0, 1} -> never counted (acc -1)
        tgt[0, 1, 1] = 1  # a single invalid (frame, joint)
        err = rng.normal(scale=[W / 25.0, H / 25.0], size=(B, J, 2))  # around the 0.5 threshold of (x / (H/10), y / (W/10))
        prd = np.clip(np.rint(tgt + err), 0, [W - 1, H - 1]).astype(int)
        cases.append((name, peaks_to_heatmaps(rng, prd, H, W), peaks_to_heatmaps(rng, tgt, H, W), 0.5))
    B, J, H, W = 4, 5, 16, 16
    tgt = np.zeros((B, J, 2), int)  # every target at (0, 0): nothing is counted, cnt = 0
    cases.append(("all_invalid", peaks_to_heatmaps(rng, rng.integers(0, 16, (B, J, 2)), H, W), peaks_to_heatmaps(rng, tgt, H, W), 0.5))
    name, o, t, _ = cases[0]
    cases.append(("square32_thr02", o, t, 0.2))  # the reference ignores thr inside accuracy(): same numbers as thr = 0.5
    for name, o, t, thr in cases:
        acc, avg_acc, cnt, pred = accuracy(o, t, "gaussian", thr)
        if not name.endswith("_thr02"):  # (same heat maps as the case it is named after)
            out[f"{name}/output"], out[f"{name}/target"] = o, t
        out[f"{name}/thr"] = np.float64(thr)
        out[f"{name}/acc"], out[f"{name}/avg_acc"], out[f"{name}/cnt"], out[f"{name}/pred"] = acc, np.float64(avg_acc), np.int64(cnt), pred
        print(f"{name}: avg_acc {avg_acc:.4f} cnt {cnt} acc {np.round(acc, 3)}")
    out["cases"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "accuracy_golden.npz"), **out)


if __name__ == "__main__":
    main()
