"""On-disk formats of the reference pipeline around the heatmap->pose stage.

The reference's stages talk through files:
  * `pred.mat`  — written by lib/dataset/PEdataset.py:116-125 (`savemat(pred_file, mdict={'preds': preds})`,
    preds float32 [N,J,3] = x, y, maxval) and read by pose_estimation/export_predicted_poses_real.py:172-173;
  * `opencv_poses.json` — written by export_predicted_poses_real.py:224-236: a list of
    {"image_name", "T" (3x1 nested list), "rotation_matrix" (3x3 nested list)}, json.dumps(indent=2).
These helpers write/read exactly those structures from this package's outputs so the rest of the
reference pipeline (and its evaluation scripts) can consume them unchanged.
"""
from __future__ import annotations

import json
import os

import numpy as np


def save_pred_mat(path: str, kpts) -> str:
    """kpts [N,J,3] float32 (StageOutput.kpts / get_final_preds output stacked with maxvals)."""
    from scipy.io import savemat

    arr = np.ascontiguousarray(_to_numpy(kpts), dtype=np.float32)
    if arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError("preds must be [N,J,3]")
    if not path.endswith(".mat"):
        path += ".mat"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    savemat(path, mdict={"preds": arr})
    return path


def load_pred_mat(path: str) -> np.ndarray:
    from scipy.io import loadmat

    return np.array(loadmat(path)["preds"])


def pose_records(image_names, rt, status=None):
    """The reference's per-frame records from PoseBatch.rt ([N,12] float64: row-major R then t).
    Frames without a pose (status != 0) get the zero R / T the GPU path reports."""
    rt = np.asarray(_to_numpy(rt), np.float64).reshape(-1, 12)
    if len(image_names) != rt.shape[0]:
        raise ValueError("one image name per frame")
    out = []
    for i, name in enumerate(image_names):
        R = rt[i, :9].reshape(3, 3)
        T = rt[i, 9:].reshape(3, 1)
        rec = {"image_name": name, "T": T.tolist(), "rotation_matrix": R.tolist()}
        if status is not None and int(status[i]) != 0:
            rec["status"] = int(status[i])  # extra key; the reference's readers ignore unknown keys
        out.append(rec)
    return out


def save_opencv_poses_json(path: str, image_names, rt, status=None) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write(json.dumps(pose_records(image_names, rt, status), indent=2))
    return path


def _to_numpy(x):
    if isinstance(x, np.ndarray):
        return x
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)
