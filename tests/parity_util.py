"""Whole-population pose parity against cv2.solvePnPRansac(iterationsCount=10000) — the reference's call,
pose_estimation/export_predicted_poses_real.py:199-201 — used by the -m gpu tests and tools/parity_report.py.

Every status-OK frame is compared.  Frames whose winning inlier set equals cv2's must meet the north_star bars
(rotation <= 1e-3 deg with the atan2 metric, translation <= 1e-4 relative) on the float64 R|t.  A frame whose inlier
set differs has to be EXPLAINED, i.e. shown to be one where cv2's own answer is decided by rounding noise:
  * cv2 itself returns another inlier set when its image points move by <= 1 float32 ulp (oracle.pnp_ref.cv2_is_unstable), or
  * the independent float64 white box (oracle.pnp_ref.ransac_epnp_numpy: OpenCV's algorithm restated in NumPy) disagrees
    with cv2 on that frame too.
Anything else is an unexplained disagreement and fails the test.
"""
from dataclasses import dataclass, field

import numpy as np

ROT_TOL_DEG = 1e-3
T_TOL_REL = 1e-4


@dataclass
class ParityReport:
    name: str
    frames: int = 0  # frames cv2 or the GPU solved (n >= 6 after the confidence filter)
    same_mask: int = 0
    status_mismatch: list = field(default_factory=list)
    max_rot_same: float = 0.0
    max_t_same: float = 0.0
    disagree: list = field(default_factory=list)  # (frame, explanation, rot error, t error)
    unexplained: list = field(default_factory=list)
    whitebox_frames: int = 0  # sample on which the white box's own agreement with cv2 was measured
    whitebox_same: int = 0
    gpu_same_on_sample: int = 0

    @property
    def agreement(self):
        return self.same_mask / max(self.frames, 1)

    def line(self):
        wb = (f"; float64 white box agrees with cv2 on {self.whitebox_same}/{self.whitebox_frames} sampled frames, the GPU on "
              f"{self.gpu_same_on_sample}/{self.whitebox_frames} of the same") if self.whitebox_frames else ""
        worst = max([d[2] for d in self.disagree], default=0.0)
        return (f"{self.name}: {self.frames} frames, inlier-set agreement with cv2@10000 {self.same_mask}/{self.frames} = {self.agreement:.4f}; "
                f"agreeing frames: max rot {self.max_rot_same:.2e} deg, max t {self.max_t_same:.2e}; disagreeing {len(self.disagree)} "
                f"(explained {len(self.disagree) - len(self.unexplained)}, unexplained {len(self.unexplained)}, worst rot error {worst:.3g} deg)"
                f"; status mismatches {len(self.status_mismatch)}{wb}")


def population_parity(name, model, kpts, out, iterations=10000, reproj=15.0, whitebox_sample=0, seed=0, conf_floor=None) -> ParityReport:
    """kpts [B,J,3] float32 (pred.mat rows), out = PoseBatch of NumPy arrays from the GPU (status, inlier_mask, rt)."""
    import cv2

    from oracle import pnp_ref

    rep = ParityReport(name)
    rng = np.random.default_rng(seed)
    B = kpts.shape[0]
    sample = set(rng.choice(B, min(whitebox_sample, B), replace=False).tolist()) if whitebox_sample else set()
    for b in range(B):
        conf = kpts[b, :, 2]
        good = pnp_ref.confidence_filter(conf) if conf_floor is None else (conf > np.float32(conf_floor))
        n = int(good.sum())
        if n < 6:
            continue  # no RANSAC: covered by the status-code tests
        idx = np.flatnonzero(good)
        obj, img = np.asarray(model.landmarks, np.float64)[good], kpts[b, good, :2].astype(np.float32)
        ok, rv, tv, inl = pnp_ref.solve_pnp_ransac_cv2(obj, img, model.K, model.dist, iterations, reproj)
        gpu_ok = int(out.status[b]) == 0
        rep.frames += 1
        cv_mask = 0 if inl is None else sum(1 << int(idx[i]) for i in inl)
        gpu_mask = int(out.inlier_mask[b]) & 0xFFFFFFFF
        same = (ok == gpu_ok) and (not ok or cv_mask == gpu_mask)
        rot = tr = 0.0
        if ok and gpu_ok:
            rot = pnp_ref.rotation_angle_deg(out.rt[b, :9].reshape(3, 3), cv2.Rodrigues(rv)[0])
            tr = float(np.linalg.norm(out.rt[b, 9:] - tv) / np.linalg.norm(tv))
        wb_same = None
        if b in sample or not same:
            wok, wR, wt, winl, _, _ = pnp_ref.ransac_epnp_numpy(obj, img, model.K, model.dist, iterations, reproj)
            wb_mask = 0 if winl is None else sum(1 << int(idx[i]) for i in winl)
            wb_same = (wok == ok) and (not ok or wb_mask == cv_mask)
        if b in sample:
            rep.whitebox_frames += 1
            rep.whitebox_same += int(wb_same)
            rep.gpu_same_on_sample += int(same)
        if same:
            rep.same_mask += 1
            if ok and n > 5 and bin(cv_mask).count("1") > 5:  # exactly-5-inlier refits are chaotic by themselves (2-D null space)
                rep.max_rot_same = max(rep.max_rot_same, rot)
                rep.max_t_same = max(rep.max_t_same, tr)
            continue
        if ok != gpu_ok:
            rep.status_mismatch.append(b)
        why = []
        if not wb_same:
            why.append("float64 white box disagrees with cv2 too")
        if pnp_ref.cv2_is_unstable(obj, img, model.K, model.dist, trials=64, iterations=iterations, reproj=reproj, seed=b):
            why.append("cv2 changes its own answer under a 1-ulp input perturbation")
        rep.disagree.append((b, "; ".join(why) or "UNEXPLAINED", rot if ok and gpu_ok else float("inf"), tr if ok and gpu_ok else float("inf")))
        if not why:
            rep.unexplained.append(b)
    return rep


def assert_parity(rep: ParityReport, min_agreement: float, max_unexplained: int = 0):
    """max_unexplained: the chaos probes are statistical (64 one-ulp perturbations of cv2's input catch a frame whose
    hypotheses flip one time in ten with probability 0.999, one that flips one time in fifty only with 0.73), so large
    populations get a slack of one frame per thousand; small, well-separated ones none."""
    print(rep.line())
    for d in rep.disagree[:12]:
        print(f"    frame {d[0]}: {d[1]}; pose differs by {d[2]:.3g} deg, {d[3]:.3g} rel-t")
    assert rep.max_rot_same <= ROT_TOL_DEG and rep.max_t_same <= T_TOL_REL, rep.line()
    assert len(rep.unexplained) <= max_unexplained, rep.line()
    assert rep.agreement >= min_agreement, rep.line()


# ----------------------------------------------------------------------------- datasets (BASELINE.json configs + close range)
DATASETS = {
    # name: (model factory name, J, heatmap (H, W), z range in metres, FP32 hypotheses of the BASELINE config)
    "B_tango_64x64": ("tango", 11, (64, 64), (4.0, 10.0), 256),
    "C_hubble17_96x72": ("hubble", 17, (96, 72), (3.0, 8.0), 256),
    "C_hubble24_96x72": ("hubble", 24, (96, 72), (3.0, 8.0), 256),
    "D_tango_128x128": ("tango", 11, (128, 128), (4.0, 10.0), 1024),
    "close_range_tango_64x64": ("tango", 11, (64, 64), (1.5, 3.0), 256),
}


def make_dataset(name, frames, seed_offset=0, chunk=256):
    """(model, kpts [frames,J,3] float32): synthetic frames of the named config rendered to heatmaps and decoded by the
    oracle's get_final_preds (the reference's decode, bit-pinned by tests/golden), i.e. the pred.mat rows the pose stage
    of the reference would read."""
    import spe_b200
    from oracle import decode_ref
    from spe_b200 import synth

    kind, J, hw, zr, _ = DATASETS[name]
    model = spe_b200.models.tango() if kind == "tango" else spe_b200.models.hubble_synthetic(J)
    out = []
    for i, lo in enumerate(range(0, frames, chunk)):
        n = min(chunk, frames - lo)
        fr = synth.make_frames(model, n, hw[0], hw[1], seed=synth.BASE_SEED + 1000 * (1 + seed_offset) + 17 * i + sum(map(ord, name)), z_range=zr)
        p, mv = decode_ref.get_final_preds_fast(True, fr.heatmaps, fr.center, fr.scale)
        out.append(np.concatenate([p, mv], -1).astype(np.float32))
    return model, np.concatenate(out, 0)
