"""Oracle (test infrastructure, see oracle/__init__.py): heatmap decoding on the CPU.

NumPy restatement of the reference's decode half:
  get_max_preds       landmark_regression/lib/core/inference.py:18-46
  get_final_preds     landmark_regression/lib/core/inference.py:49-79
  transform_preds     landmark_regression/lib/utils/transforms.py:49-54
  get_affine_transform / get_3rd_point / get_dir   transforms.py:57-110
  affine_transform    transforms.py:92-95

Two flavours of get_final_preds are provided:
  * `get_final_preds`        — same loop structure as the reference (a Python loop over frames
    and landmarks, one cv2.getAffineTransform per frame).  This is what `bench.py` times as the
    CPU baseline, because its cost profile is the reference's.
  * `get_final_preds_fast`   — vectorised over the batch; used by tests at sizes where the
    looped version would take minutes.  tests/test_oracle_decode.py asserts both agree bit for bit.
Both return (preds float32 [B,J,2], maxvals float32 [B,J,1]); `return_index=True` adds the flat
argmax index int64 [B,J] that the GPU parity tests compare bit-exactly.
"""
from __future__ import annotations

import math

import cv2
import numpy as np


# ----------------------------------------------------------------------------- argmax
def get_max_preds(batch_heatmaps, return_index: bool = False):
    """inference.py:18-46 — flat argmax (first index on ties, NaN wins) -> (x, y) float32,
    zeroed unless maxval > 0."""
    assert isinstance(batch_heatmaps, np.ndarray), "batch_heatmaps should be numpy.ndarray"
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    B, J, _, W = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(B, J, -1)
    idx = flat.argmax(axis=2)  # inference.py:31
    maxvals = flat.max(axis=2).reshape(B, J, 1)  # inference.py:32
    fidx = idx.astype(np.float32)
    preds = np.empty((B, J, 2), np.float32)
    preds[..., 0] = fidx % W  # inference.py:39
    preds[..., 1] = np.floor(fidx / W)  # inference.py:40
    preds *= (maxvals > 0.0).astype(np.float32)  # inference.py:42-45 (NaN > 0 is False)
    if return_index:
        return preds, maxvals, idx
    return preds, maxvals


# ----------------------------------------------------------------------------- affine
def _third_point(a, b):
    """transforms.py:98-100 (float32 in, float32 out)."""
    d = a - b
    return b + np.array([-d[1], d[0]], dtype=np.float32)


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """transforms.py:57-89.  Anchors are built in float32; only scale[0] enters (src_w)."""
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale])
    scale_px = scale * 200.0  # transforms.py:65
    src_w = scale_px[0]
    dst_w, dst_h = output_size[0], output_size[1]
    ang = np.pi * rot / 180
    sn, cs = np.sin(ang), np.cos(ang)
    p = [0, src_w * -0.5]
    src_dir = [p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs]  # get_dir, transforms.py:103-110
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_px * shift
    src[1, :] = center + src_dir + scale_px * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    src[2, :] = _third_point(src[0, :], src[1, :])
    dst[2, :] = _third_point(dst[0, :], dst[1, :])
    if inv:
        return cv2.getAffineTransform(np.float32(dst), np.float32(src))  # transforms.py:85
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def transform_preds(coords, center, scale, output_size):
    """transforms.py:49-54 — heatmap pixels -> image pixels through the inverse box affine (f64)."""
    trans = get_affine_transform(center, scale, 0, output_size, inv=1)
    out = np.zeros(coords.shape)
    for p in range(coords.shape[0]):
        out[p, 0:2] = trans @ np.array([coords[p, 0], coords[p, 1], 1.0])
    return out


# ----------------------------------------------------------------------------- full decode
def get_final_preds(post_process: bool, batch_heatmaps, center, scale, return_index: bool = False):
    """inference.py:49-79 with the reference's loop structure (`config.TEST.POST_PROCESS` is the
    only thing the reference reads from its config, passed here as a bool)."""
    coords, maxvals, idx = get_max_preds(batch_heatmaps, return_index=True)
    H, W = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    if post_process:
        for n in range(coords.shape[0]):
            for p in range(coords.shape[1]):
                hm = batch_heatmaps[n][p]
                px = int(math.floor(coords[n][p][0] + 0.5))
                py = int(math.floor(coords[n][p][1] + 0.5))
                if 1 < px < W - 1 and 1 < py < H - 1:  # inference.py:62
                    d = np.array([hm[py][px + 1] - hm[py][px - 1], hm[py + 1][px] - hm[py - 1][px]])
                    coords[n][p] += np.sign(d) * 0.25
    preds = coords.copy()
    for i in range(coords.shape[0]):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [W, H])  # f64 -> f32 on store
    if return_index:
        return preds, maxvals, idx
    return preds, maxvals


def get_final_preds_fast(post_process: bool, batch_heatmaps, center, scale, return_index: bool = False):
    """Same results as get_final_preds, vectorised over the batch (one getAffineTransform per
    frame remains — it is the bit-defining step)."""
    coords, maxvals, idx = get_max_preds(batch_heatmaps, return_index=True)
    B, J, H, W = batch_heatmaps.shape
    if post_process:
        px = np.floor(coords[..., 0] + 0.5).astype(np.int64)
        py = np.floor(coords[..., 1] + 0.5).astype(np.int64)
        ok = (px > 1) & (px < W - 1) & (py > 1) & (py < H - 1)
        pxc, pyc = np.clip(px, 1, W - 2), np.clip(py, 1, H - 2)
        bi, ji = np.meshgrid(np.arange(B), np.arange(J), indexing="ij")
        dx = batch_heatmaps[bi, ji, pyc, pxc + 1] - batch_heatmaps[bi, ji, pyc, pxc - 1]
        dy = batch_heatmaps[bi, ji, pyc + 1, pxc] - batch_heatmaps[bi, ji, pyc - 1, pxc]
        coords[..., 0] += np.where(ok, np.sign(dx) * np.float32(0.25), np.float32(0)).astype(np.float32)
        coords[..., 1] += np.where(ok, np.sign(dy) * np.float32(0.25), np.float32(0)).astype(np.float32)
    preds = np.empty_like(coords)
    for i in range(B):
        t = get_affine_transform(center[i], scale[i], 0, [W, H], inv=1)
        # np.dot(t, [x, y, 1]) as the reference does it: products summed left to right in float64
        x = coords[i, :, 0].astype(np.float64)
        y = coords[i, :, 1].astype(np.float64)
        preds[i, :, 0] = (t[0, 0] * x + t[0, 1] * y + t[0, 2]).astype(np.float32)
        preds[i, :, 1] = (t[1, 0] * x + t[1, 1] * y + t[1, 2]).astype(np.float32)
    if return_index:
        return preds, maxvals, idx
    return preds, maxvals


def inverse_affine_closed_form(center, scale, W, H):
    """closed form that replays the float32 roundings of get_affine_transform.
    Returns (ax, bx, ay, by) float64 with X = ax*x + bx, Y = ay*y + by.  This is what the CUDA
    kernel implements; kept here so tests can check kernel == closed form == reference."""
    cx, cy = np.float32(center[0]), np.float32(center[1])
    sw = np.float32(np.float32(scale[0]) * np.float32(200.0))
    q1y = np.float32(cy - np.float32(sw * np.float32(0.5)))
    d = np.float32(cy - q1y)
    q2x = np.float32(cx - d)
    ax = (float(cx) - float(q2x)) / (W / 2.0)
    ay = (float(cy) - float(q1y)) / (W / 2.0)
    return ax, float(cx) - ax * (W / 2.0), ay, float(cy) - ay * (H / 2.0)
