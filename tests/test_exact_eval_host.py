"""The arithmetic of the float64 replay (SPE_FLAG_EXACT), checked against cv2 WITHOUT a GPU.

csrc/ransac_exact_eval.cuh (one 5-point EPnP hypothesis in float64 + the reprojection test) and
csrc/ransac_common.cuh (RANSACUpdateNumIters) are __host__ __device__; tests/host/exact_eval_host.cu compiles the very
same functions for the host and drives them with the sequential loop replay_kernel runs per frame.  The result must be
cv2.solvePnPRansac's inlier set (iterationsCount = 10000, the reference's call) on the population — the GPU tests then
only have to show that the kernel reproduces this host build bit for bit (tests/test_parity_gpu.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from parity_util import make_dataset

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "exact_eval_host.cu")
OUT = os.path.join(HERE, "host", "exact_eval_host.so")


@pytest.fixture(scope="module")
def host_lib():
    csrc = os.path.join(os.path.dirname(HERE), "spacecraft-pose-estimation_b200", "csrc")
    deps = [SRC] + [os.path.join(csrc, f) for f in ("ransac_exact_eval.cuh", "ransac_common.cuh", "epnp_math.cuh", "epnp_f64.cuh", "p3p_f64.cuh", "ransac.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        cmd = ["nvcc", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
               "-o", OUT, SRC]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
    L = ctypes.CDLL(OUT)
    L.spe_host_replay_frame.restype = ctypes.c_int
    return L


def host_replay(L, model, obj, img, iterations=10000, want=0):
    from oracle import epnp_ref, ocv_rng

    n = len(obj)
    obj32 = np.ascontiguousarray(np.asarray(obj).astype(np.float32).astype(np.float64))
    img32 = np.ascontiguousarray(img, np.float32)
    und = np.ascontiguousarray(epnp_ref.undistort_points(img32.astype(np.float64), model.K, model.dist))  # what frame_prep_kernel stores
    cam = np.array([model.K[0, 0], model.K[1, 1], model.K[0, 2], model.K[1, 2], *model.dist[:5]], np.float64)
    subs = np.ascontiguousarray(ocv_rng.minimal_sets(n, iterations).astype(np.uint8))
    mask, vis = ctypes.c_uint32(), ctypes.c_int32()
    hm = np.zeros(max(want, 1), np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    w = L.spe_host_replay_frame(p(obj32), p(und), p(img32), n, p(cam), p(subs), iterations, ctypes.c_float(15.0), ctypes.c_double(0.99),
                                ctypes.byref(mask), ctypes.byref(vis), p(hm), want)
    return w, mask.value, vis.value, hm


@pytest.mark.parametrize("name,frames,min_agree", [("B_tango_64x64", 128, 0.99), ("close_range_tango_64x64", 128, 0.94), ("C_hubble17_96x72", 96, 0.97)])
def test_host_build_of_the_replay_arithmetic_matches_cv2(host_lib, name, frames, min_agree):
    from oracle import pnp_ref

    model, kpts = make_dataset(name, frames)
    same = tot = 0
    visited = []
    for b in range(frames):
        good = pnp_ref.confidence_filter(kpts[b, :, 2])
        if good.sum() < 6:
            continue
        obj, img = model.landmarks[good], kpts[b, good, :2].astype(np.float32)
        ok, rv, tv, inl = pnp_ref.solve_pnp_ransac_cv2(obj, img, model.K, model.dist)
        w, mask, vis, _ = host_replay(host_lib, model, obj, img)
        cv_mask = 0 if inl is None else sum(1 << int(i) for i in inl)
        tot += 1
        same += int(ok == (w >= 0) and (not ok or cv_mask == mask))
        visited.append(vis)
    print(f"{name}: host build of the replay arithmetic == cv2 on {same}/{tot} frames; cv2 looks at {np.mean(visited):.1f} hypotheses per frame "
          f"(90th percentile {np.percentile(visited, 90):.0f}, max {max(visited)})")
    assert tot >= 0.9 * frames and same / tot >= min_agree


def test_known_answer_frame(host_lib, pnp_golden):
    """one outlier of 11, hypothesis 2 wins with 10 inliers, the budget drops to 5."""
    import spe_b200

    g = pnp_golden
    m = spe_b200.models.tango()
    w, mask, vis, hm = host_replay(host_lib, m, g["landmarks"], g["e3_img"], iterations=256, want=8)
    assert w == 2 and vis == 5
    assert mask == sum(1 << int(i) for i in g["e3_inliers"])
    assert [bin(int(x)).count("1") for x in hm[:5]] == [0, 0, 10, 10, 10]


def test_host_build_of_the_p3p_branch_matches_cv2(host_lib):
    """cv2's n == 4 branch (solvePnP with SOLVEPNP_P3P on the four points): csrc/p3p_f64.cuh compiled for the host."""
    import cv2

    import spe_b200
    from oracle import epnp_ref, pnp_ref
    from spe_b200 import synth

    host_lib.spe_host_p3p.restype = ctypes.c_int
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    good = total = 0
    for model in (spe_b200.models.tango(), spe_b200.models.hubble_synthetic(17)):
        rng = np.random.default_rng(1)
        B, J = 300, model.num_landmarks
        rvec, tvec = synth.random_poses(rng, B, z_range=(2.0, 10.0))
        pts = synth.project(model.landmarks, synth.rodrigues(rvec), tvec, model.K, model.dist) + rng.normal(scale=1.0, size=(B, J, 2))
        cam = np.array([model.K[0, 0], model.K[1, 1], model.K[0, 2], model.K[1, 2], *model.dist[:5]])
        for b in range(B):
            idx = np.sort(rng.choice(J, 4, replace=False))
            obj, img = model.landmarks[idx], pts[b, idx].astype(np.float32)
            ok, rv, tv, inl = cv2.solvePnPRansac(obj, img, model.K, distCoeffs=model.dist, flags=cv2.SOLVEPNP_EPNP, iterationsCount=10000, reprojectionError=15.0)
            if not (ok and np.all(np.isfinite(rv)) and np.all(np.isfinite(tv))):
                continue  # cv2 reports NaN poses when the quartic has no real root
            obj32 = np.ascontiguousarray(obj.astype(np.float32).astype(np.float64))
            und = np.ascontiguousarray(epnp_ref.undistort_points(img.astype(np.float64), model.K, model.dist))
            Rt = np.zeros(12)
            total += 1
            if not host_lib.spe_host_p3p(p(obj32), p(und), p(cam), p(Rt)):
                continue
            r = pnp_ref.rotation_angle_deg(Rt[:9].reshape(3, 3), cv2.Rodrigues(rv)[0])
            t = np.linalg.norm(Rt[9:] - tv.ravel()) / np.linalg.norm(tv)
            good += int(r <= 1e-3 and t <= 1e-4)
    print(f"P3P branch (host build) within 1e-3 deg / 1e-4 of cv2 on {good}/{total} four-point frames")
    assert total >= 550 and good >= 0.99 * total
