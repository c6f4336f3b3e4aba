"""Seeded synthetic workload for the heatmap->pose stage.

One generator feeds the GPU path, the oracle and the CPU baseline so they all see identical
bytes.  Everything here is host-side NumPy; `device_heatmaps` is the torch/CUDA variant used only
to fill HBM for throughput runs (same formula, different random stream).

Frame recipe (per frame):
  * pose: rotation axis uniform on S^2, angle U(0, pi); t = (U(-.3,.3), U(-.2,.2), U(zmin,zmax)) m
  * image landmarks: pinhole + 5-coefficient distortion — the formula of
    object_detection/speed_plus_utils/utils.py:108-139 (== cv2.projectPoints)
  * detection box: min/max of the landmarks grown by 10 % of the width on every side
    (object_detection/speedplus_to_coco_dicts.py:106-117), widened to the heatmap aspect ratio so
    every landmark lands inside the map; center = xy + wh/2, scale = wh/200*1.5 in float32
    (landmark_regression/lib/dataset/PEdataset.py:98-113)
  * heatmap [J,H,W] float32: exp(-|p-h|^2 / (2 sigma^2)) + N(0, noise^2) with the peak at
    h = (u - center)/a + (W/2, H/2), a = scale_x*200/W (the inverse of transform_preds,
    landmark_regression/lib/utils/transforms.py:49-89, which is isotropic in scale_x)
  * with probability p_outlier a landmark's peak is moved to a uniformly random pixel (gross
    outlier), with probability p_masked the whole map is <= 0 (landmark not detected).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .models import CameraModel

BASE_SEED = 20261017


@dataclass
class Frames:
    heatmaps: np.ndarray  # [B,J,H,W] float32
    center: np.ndarray  # [B,2] float32
    scale: np.ndarray  # [B,2] float32
    rvec: np.ndarray  # [B,3] float64 ground truth
    tvec: np.ndarray  # [B,3] float64 ground truth
    image_points: np.ndarray  # [B,J,2] float64 exact projections (before heatmap quantisation)
    peak_hm: np.ndarray  # [B,J,2] float64 peak position in heatmap pixels (after outlier moves)
    outlier: np.ndarray  # [B,J] bool
    masked: np.ndarray  # [B,J] bool


def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """Batched axis-angle -> rotation matrix, [B,3] -> [B,3,3] (float64)."""
    rvec = np.asarray(rvec, np.float64).reshape(-1, 3)
    theta = np.linalg.norm(rvec, axis=1)
    k = rvec / np.where(theta > 0, theta, 1.0)[:, None]
    Kx = np.zeros((len(rvec), 3, 3))
    Kx[:, 0, 1], Kx[:, 0, 2] = -k[:, 2], k[:, 1]
    Kx[:, 1, 0], Kx[:, 1, 2] = k[:, 2], -k[:, 0]
    Kx[:, 2, 0], Kx[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(theta)[:, None, None], np.cos(theta)[:, None, None]
    return np.eye(3)[None] + s * Kx + (1.0 - c) * (Kx @ Kx)


def project(landmarks: np.ndarray, R: np.ndarray, t: np.ndarray, K: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """Pinhole + (k1,k2,p1,p2,k3) distortion. landmarks [J,3], R [B,3,3], t [B,3] -> [B,J,2]."""
    pc = np.einsum("bij,nj->bni", R, landmarks) + t[:, None, :]
    x = pc[..., 0] / pc[..., 2]
    y = pc[..., 1] / pc[..., 2]
    k1, k2, p1, p2, k3 = (float(v) for v in dist)
    r2 = x * x + y * y
    cd = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * cd + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y * cd + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
    return np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], axis=-1)


def random_poses(rng: np.random.Generator, n: int, z_range=(4.0, 10.0)):
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    rvec = axis * rng.uniform(0.0, np.pi, size=(n, 1))
    tvec = np.stack(
        [rng.uniform(-0.3, 0.3, n), rng.uniform(-0.2, 0.2, n), rng.uniform(z_range[0], z_range[1], n)], axis=1
    )
    return rvec, tvec


def boxes_to_center_scale(pts: np.ndarray, hm_w: int, hm_h: int):
    """Detection box -> (center, scale) exactly as the reference data layer does it, in float32."""
    mn, mx = pts.min(axis=1), pts.max(axis=1)
    w, h = mx[:, 0] - mn[:, 0], mx[:, 1] - mn[:, 1]
    x0, y0 = mn[:, 0] - 0.1 * w, mn[:, 1] - 0.1 * w  # both tolerances use the width (reference quirk)
    bw, bh = w + 0.2 * w, h + 0.2 * w
    cx, cy = x0 + 0.5 * bw, y0 + 0.5 * bh
    bw = np.maximum(bw, bh * hm_w / hm_h)  # keep every landmark inside the (isotropic) crop
    center = np.stack([cx, cy], axis=1).astype(np.float32)
    scale = (np.stack([bw, bh], axis=1) / 200.0).astype(np.float32) * np.float32(1.5)
    return center, scale


def make_frames(
    model: CameraModel,
    batch: int,
    hm_h: int = 64,
    hm_w: int = 64,
    seed: int = BASE_SEED,
    sigma: float = 2.0,
    noise: float = 0.01,
    p_outlier: float = 0.10,
    p_masked: float = 0.02,
    z_range=(4.0, 10.0),
) -> Frames:
    rng = np.random.default_rng(seed)
    J = model.num_landmarks
    rvec, tvec = random_poses(rng, batch, z_range)
    pts = project(model.landmarks, rodrigues(rvec), tvec, model.K, model.dist)
    center, scale = boxes_to_center_scale(pts, hm_w, hm_h)
    a = scale[:, 0].astype(np.float64) * 200.0 / hm_w
    peak = (pts - center[:, None, :].astype(np.float64)) / a[:, None, None] + np.array([hm_w / 2.0, hm_h / 2.0])
    outlier = rng.random((batch, J)) < p_outlier
    masked = rng.random((batch, J)) < p_masked
    rand_px = np.stack([rng.uniform(0, hm_w - 1, (batch, J)), rng.uniform(0, hm_h - 1, (batch, J))], axis=-1)
    peak = np.where(outlier[..., None], rand_px, peak)
    heatmaps = render_heatmaps(peak, hm_h, hm_w, sigma, noise, masked, rng)
    return Frames(heatmaps, center, scale, rvec, tvec, pts, peak, outlier, masked)


def render_heatmaps(peak, hm_h, hm_w, sigma, noise, masked, rng, chunk: int = 1024) -> np.ndarray:
    B, J = peak.shape[:2]
    out = np.empty((B, J, hm_h, hm_w), np.float32)
    ys = np.arange(hm_h, dtype=np.float64)[None, None, :, None]
    xs = np.arange(hm_w, dtype=np.float64)[None, None, None, :]
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        px = peak[s:e, :, 0][:, :, None, None]
        py = peak[s:e, :, 1][:, :, None, None]
        g = np.exp(-((xs - px) ** 2 + (ys - py) ** 2) / (2.0 * sigma * sigma))
        g += rng.normal(scale=noise, size=g.shape)
        m = masked[s:e][:, :, None, None]
        g = np.where(m, -np.abs(g) - 1e-3, g)  # undetected landmark: every value <= 0
        out[s:e] = g.astype(np.float32)
    return out


def device_heatmaps(model: CameraModel, batch: int, hm_h: int, hm_w: int, seed: int, device, sigma=2.0, noise=0.01,
                    p_outlier=0.10, p_masked=0.02, chunk: int = 4096):
    """Throughput-run variant: same frame recipe, heatmaps rendered directly in HBM with torch.

    Returns (heatmaps [B,J,H,W] f32 cuda, center [B,2] f32 cuda, scale [B,2] f32 cuda).
    Poses/boxes come from the NumPy recipe (cheap); only the B*J*H*W render runs on the device.
    """
    import torch

    rng = np.random.default_rng(seed)
    J = model.num_landmarks
    rvec, tvec = random_poses(rng, batch)
    pts = project(model.landmarks, rodrigues(rvec), tvec, model.K, model.dist)
    center, scale = boxes_to_center_scale(pts, hm_w, hm_h)
    a = scale[:, 0].astype(np.float64) * 200.0 / hm_w
    peak = (pts - center[:, None, :].astype(np.float64)) / a[:, None, None] + np.array([hm_w / 2.0, hm_h / 2.0])
    outlier = rng.random((batch, J)) < p_outlier
    masked = rng.random((batch, J)) < p_masked
    rand_px = np.stack([rng.uniform(0, hm_w - 1, (batch, J)), rng.uniform(0, hm_h - 1, (batch, J))], axis=-1)
    peak = np.where(outlier[..., None], rand_px, peak)

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    hm = torch.empty((batch, J, hm_h, hm_w), dtype=torch.float32, device=device)
    ys = torch.arange(hm_h, dtype=torch.float32, device=device)[None, None, :, None]
    xs = torch.arange(hm_w, dtype=torch.float32, device=device)[None, None, None, :]
    peak_t = torch.from_numpy(peak.astype(np.float32)).to(device)
    masked_t = torch.from_numpy(masked).to(device)
    for s in range(0, batch, chunk):
        e = min(batch, s + chunk)
        px = peak_t[s:e, :, 0][:, :, None, None]
        py = peak_t[s:e, :, 1][:, :, None, None]
        g = torch.exp(-((xs - px) ** 2 + (ys - py) ** 2) / (2.0 * sigma * sigma))
        g += noise * torch.randn(g.shape, generator=gen, device=device, dtype=torch.float32)
        g = torch.where(masked_t[s:e][:, :, None, None], -g.abs() - 1e-3, g)
        hm[s:e] = g
    return hm, torch.from_numpy(center).to(device), torch.from_numpy(scale).to(device)
