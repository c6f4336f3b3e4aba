"""Generate the committed golden vectors by running THE REFERENCE ITSELF (and cv2) here.

Run in the build container only (needs /root/reference; the GPU box has no copy):
    python tests/golden/make_golden.py
Writes tests/golden/decode_golden.npz and tests/golden/pnp_golden.npz.

decode_golden.npz — inputs and outputs of the reference's own
    core.inference.get_max_preds / get_final_preds  (landmark_regression/lib/core/inference.py:18-79)
    utils.transforms.transform_preds                (landmark_regression/lib/utils/transforms.py:49-54)
on small seeded cases that cover ties, NaNs, all-negative maps, border peaks, odd sizes and
non-square maps.
pnp_golden.npz — cv2.solvePnPRansac (cv2 version recorded inside) called with the arguments of
    pose_estimation/export_predicted_poses_real.py:199-201 on keypoints decoded by the reference
    from seeded synthetic Tango frames, plus the per-hypothesis trace of the white-box restatement.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
sys.path.insert(0, "/root/reference/landmark_regression/lib")

from core.inference import get_final_preds, get_max_preds  # noqa: E402  (the reference)
from utils.transforms import transform_preds  # noqa: E402  (the reference)

from oracle import pnp_ref  # noqa: E402
from spe_b200 import models, synth  # noqa: E402


class _Cfg:
    class TEST:
        POST_PROCESS = True


class _CfgOff:
    class TEST:
        POST_PROCESS = False


def edge_case_heatmaps(rng, H, W):
    """[1, 12, H, W] maps exercising the corner cases of App. A."""
    J = 12
    hm = rng.normal(scale=0.05, size=(1, J, H, W)).astype(np.float32)
    hm[0, 0] = -np.abs(hm[0, 0]) - 0.01  # all negative -> coords (0,0), maxval < 0
    hm[0, 1, :, :] = 0.0  # all zero -> argmax 0, not > 0
    hm[0, 2, H // 2, W // 2] = 1.0
    hm[0, 2, H // 2 + 1, W // 3] = 1.0  # tie -> first index wins
    hm[0, 3, 0, 0] = 2.0  # corner, never refined
    hm[0, 4, H - 1, W - 1] = 2.0  # last element
    hm[0, 5, 1, W // 2] = 2.0  # py == 1 -> not refined
    hm[0, 6, H // 2, W - 2] = 2.0  # px == W-2 -> refined
    hm[0, 7, H // 2, W - 1] = 2.0  # px == W-1 -> not refined
    hm[0, 8, H // 3, W // 3] = np.nan  # NaN wins the argmax; NaN > 0 is False
    hm[0, 8, H // 2, W // 2] = 5.0
    hm[0, 9, 2, 2] = 3.0  # smallest refinable position
    hm[0, 9, 2, 3] = hm[0, 9, 2, 1]  # equal neighbours -> sign(0) = 0 in x
    hm[0, 10, H - 2, 2] = 3.0
    hm[0, 11] = -0.0
    hm[0, 11, H // 2, W // 2 + 1] = 0.0  # +0 == -0: first index (0) wins
    return hm


def main():
    rng = np.random.default_rng(12345)
    out = {}
    cases = []
    tango = models.tango()
    fr = synth.make_frames(tango, 4, 64, 64, seed=synth.BASE_SEED)
    cases.append(("tango64", fr.heatmaps, fr.center, fr.scale))
    hub = models.hubble_synthetic(17)
    fr2 = synth.make_frames(hub, 2, 96, 72, seed=synth.BASE_SEED + 2, z_range=(3.0, 8.0))
    cases.append(("hubble96x72", fr2.heatmaps, fr2.center, fr2.scale))
    fr3 = synth.make_frames(tango, 1, 128, 128, seed=synth.BASE_SEED + 3)
    cases.append(("tango128", fr3.heatmaps, fr3.center, fr3.scale))
    for (H, W) in ((64, 64), (17, 19), (7, 5), (48, 36)):
        hm = edge_case_heatmaps(rng, H, W)
        c = np.array([[rng.uniform(100, 1800), rng.uniform(100, 1100)]], np.float32)
        s = np.array([[rng.uniform(0.3, 9.0), rng.uniform(0.3, 9.0)]], np.float32)
        cases.append((f"edge{H}x{W}", hm, c, s))
    names = []
    for name, hm, c, s in cases:
        p0, m0 = get_max_preds(hm)
        p1, m1 = get_final_preds(_Cfg, hm, c, s)
        p2, _ = get_final_preds(_CfgOff, hm, c, s)
        out[f"{name}/hm"], out[f"{name}/center"], out[f"{name}/scale"] = hm, c, s
        out[f"{name}/max_preds"], out[f"{name}/maxvals"] = p0, m0
        out[f"{name}/final_preds"], out[f"{name}/final_preds_nopp"] = p1, p2
        out[f"{name}/argmax"] = hm.reshape(hm.shape[0], hm.shape[1], -1).argmax(2).astype(np.int64)
        names.append(name)
    out["names"] = np.array(names)
    # transform_preds known answers (SURVEY App. E.1 regenerated from the reference)
    tp_in = np.array([[10.25, 20.75], [0, 0], [63, 63]], np.float32)
    out["tp/coords"] = tp_in
    out["tp/center"] = np.array([960.5, 600.25], np.float32)
    out["tp/scale"] = np.array([3.1, 2.5], np.float32)
    out["tp/out64"] = transform_preds(tp_in, out["tp/center"], out["tp/scale"], [64, 64])
    # wide sweep of centres/scales/sizes: float64 outputs of the reference's transform_preds
    sweep = []
    for (W, H) in ((64, 64), (72, 96), (128, 128), (384, 384), (768, 768)):
        for _ in range(40):
            c = np.array([rng.uniform(0, 1920), rng.uniform(0, 1200)], np.float32)
            s = np.array([rng.uniform(0.15, 14.0), rng.uniform(0.15, 14.0)], np.float32)
            xy = np.stack([rng.integers(0, W, 6) + rng.choice([0, 0.25, -0.25], 6),
                           rng.integers(0, H, 6) + rng.choice([0, 0.25, -0.25], 6)], 1).astype(np.float32)
            o = transform_preds(xy, c, s, [W, H])
            sweep.append(np.concatenate([[W, H], c, s, xy.ravel(), o.ravel()]))
    out["tp/sweep"] = np.array(sweep, np.float64)  # [W,H,cx,cy,sx,sy, 12 coords, 12 outputs]
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **out)

    # ---- PnP golden: the reference's call on reference-decoded keypoints
    fr = synth.make_frames(tango, 48, 64, 64, seed=synth.BASE_SEED + 7)
    preds, maxvals = get_final_preds(_Cfg, fr.heatmaps, fr.center, fr.scale)
    kpts = np.concatenate([preds, maxvals], axis=-1).astype(np.float32)  # pred.mat layout [N,J,3]
    H = 256
    g = dict(kpts=kpts, landmarks=tango.landmarks, K=tango.K, dist=tango.dist, cv2_version=np.array(cv2.__version__))
    oks, rvecs, tvecs, inl_masks, pose7 = [], [], [], [], []
    oks256, rvecs256, tvecs256 = [], [], []
    counts, masks, winners, evaluated = [], [], [], []
    for b in range(kpts.shape[0]):
        ok, p7, mask, rv, tv = pnp_ref.pose_from_keypoints(kpts[b], tango.landmarks, tango.K, tango.dist)
        oks.append(ok), rvecs.append(rv), tvecs.append(tv), inl_masks.append(mask), pose7.append(p7)
        ok2, _, _, rv2, tv2 = pnp_ref.pose_from_keypoints(kpts[b], tango.landmarks, tango.K, tango.dist, iterations=H)
        oks256.append(ok2), rvecs256.append(rv2), tvecs256.append(tv2)
        good = pnp_ref.confidence_filter(kpts[b, :, 2])
        n = int(good.sum())
        if n >= 6:
            tr = pnp_ref.ransac_epnp_whitebox(tango.landmarks[good], kpts[b, good, :2], tango.K, tango.dist,
                                              iterations=H, exhaustive=H)
            counts.append(tr.counts), masks.append(tr.masks), winners.append(tr.winner), evaluated.append(tr.evaluated)
        else:
            counts.append(np.zeros(H, np.int32)), masks.append(np.zeros(H, np.uint32)), winners.append(-2), evaluated.append(0)
    g.update(ok=np.array(oks), rvec=np.array(rvecs), tvec=np.array(tvecs), inlier_mask=np.array(inl_masks, np.uint32),
             pose7=np.array(pose7), ok256=np.array(oks256), rvec256=np.array(rvecs256), tvec256=np.array(tvecs256),
             hyp_counts=np.array(counts), hyp_masks=np.array(masks), winner=np.array(winners), evaluated=np.array(evaluated))
    # SURVEY App. E.3 known answer (regenerated)
    rv0, tv0 = np.array([0.3, -0.5, 1.0]), np.array([0.1, -0.1, 6.0])
    p, _ = cv2.projectPoints(tango.landmarks, rv0, tv0, tango.K, tango.dist)
    p = np.round(p.reshape(-1, 2), 2)
    p[3] += 150.0
    p = p.astype(np.float32)
    ok, rv, tv, inl = cv2.solvePnPRansac(tango.landmarks, p, tango.K, distCoeffs=tango.dist, flags=cv2.SOLVEPNP_EPNP,
                                         iterationsCount=256, reprojectionError=15.0)
    g.update(e3_img=p, e3_ok=np.array(ok), e3_rvec=rv.ravel(), e3_tvec=tv.ravel(), e3_inliers=inl.ravel())
    np.savez_compressed(os.path.join(HERE, "pnp_golden.npz"), **g)
    print("wrote golden vectors; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
