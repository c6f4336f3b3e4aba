"""The pose oracle pinned against OpenCV (the un-vendored third-party dependency that holds the
reference's pose arithmetic) and against the committed golden vectors (cv2 4.13.0)."""
import cv2
import numpy as np
import pytest

from oracle import epnp_ref, ocv_rng, pnp_ref
from spe_b200 import models, synth


def test_rng_known_answers():
    """SURVEY App. E.2."""
    r = ocv_rng.CvRNG()
    assert [r.next() for _ in range(4)] == [130063605, 3133359004, 2578348940, 925327173]
    assert ocv_rng.minimal_sets(11, 6).tolist() == [[1, 6, 2, 0, 3], [2, 8, 3, 1, 6], [10, 9, 0, 1, 6], [2, 0, 7, 5, 9], [6, 1, 2, 4, 5], [5, 0, 2, 3, 1]]
    assert ocv_rng.minimal_sets(17, 3).tolist() == [[5, 9, 12, 3, 10], [2, 16, 7, 1, 12], [10, 2, 14, 16, 8]]
    assert ocv_rng.minimal_sets(24, 3).tolist() == [[21, 4, 20, 15, 8], [21, 8, 5, 6, 1], [10, 18, 1, 2, 15]]
    for n in (6, 11, 24):
        s = ocv_rng.minimal_sets(n, 300)
        assert all(len(set(row)) == 5 for row in s.tolist()) and s.min() >= 0 and s.max() < n


def test_budget_table():
    """SURVEY App. B.6."""
    got = {g: ocv_rng.update_num_iters(0.99, (11 - g) / 11, 5, 10000) for g in range(5, 12)}
    assert got == {5: 235, 6: 93, 7: 42, 8: 20, 9: 10, 10: 5, 11: 0}
    assert ocv_rng.update_num_iters(0.99, (17 - 5) / 17, 5, 10000) == 2090
    assert ocv_rng.update_num_iters(0.99, (24 - 5) / 24, 5, 10000) == 10000
    assert ocv_rng.update_num_iters(0.99, (24 - 6) / 24, 5, 10000) == 4713


def test_select_sequential_trace():
    """App. E.3 trace: hypothesis 2 wins with 10 inliers, budget 256 -> 5, loop ends at h = 5."""
    counts = [0, 0, 10, 10, 10, 0, 10, 0, 1, 10, 1, 10, 10, 0, 0, 0]
    assert ocv_rng.select_sequential(counts, 11, 256) == (2, 5)
    assert ocv_rng.select_sequential([4, 3, 4, 0], 11, 4) == (-1, 4)
    # strictly-better rule: a later, larger count inside the budget replaces the first
    assert ocv_rng.select_sequential([6, 0, 0, 9, 11], 11, 256)[0] == 4


def _frames(n=24, seed=3):
    m = models.tango()
    rng = np.random.default_rng(seed)
    rvec, tvec = synth.random_poses(rng, n)
    pts = synth.project(m.landmarks, synth.rodrigues(rvec), tvec, m.K, m.dist)
    pts += rng.normal(scale=1.0, size=pts.shape)
    for b in range(n):
        for j in rng.choice(11, rng.integers(0, 4), replace=False):
            pts[b, j] += rng.uniform(40, 300, 2) * rng.choice([-1, 1], 2)
    return m, pts.astype(np.float32)


def test_whitebox_equals_cv2_bitwise():
    m, pts = _frames()
    for b in range(len(pts)):
        for iters in (256, 10000):
            ok, rv, tv, inl = pnp_ref.solve_pnp_ransac_cv2(m.landmarks, pts[b], m.K, m.dist, iterations=iters)
            tr = pnp_ref.ransac_epnp_whitebox(m.landmarks, pts[b], m.K, m.dist, iterations=iters)
            assert ok == tr.ok
            if ok:
                np.testing.assert_array_equal(rv, tr.rvec)
                np.testing.assert_array_equal(tv, tr.tvec)
                np.testing.assert_array_equal(inl, tr.inliers)


def test_whitebox_subset_of_points_and_failure():
    m, pts = _frames(6, seed=9)
    keep = np.array([0, 1, 2, 4, 6, 7, 9, 10])
    for b in range(len(pts)):
        ok, rv, tv, inl = pnp_ref.solve_pnp_ransac_cv2(m.landmarks[keep], pts[b, keep], m.K, m.dist)
        tr = pnp_ref.ransac_epnp_whitebox(m.landmarks[keep], pts[b, keep], m.K, m.dist)
        assert ok == tr.ok
        if ok:
            np.testing.assert_array_equal(rv, tr.rvec)
            np.testing.assert_array_equal(inl, tr.inliers)
    rng = np.random.default_rng(0)
    junk = rng.uniform(0, 1200, (11, 2)).astype(np.float32)
    ok, *_ = pnp_ref.solve_pnp_ransac_cv2(m.landmarks, junk, m.K, m.dist, iterations=64)
    tr = pnp_ref.ransac_epnp_whitebox(m.landmarks, junk, m.K, m.dist, iterations=64)
    assert ok == tr.ok
    with pytest.raises(ValueError):
        pnp_ref.ransac_epnp_whitebox(m.landmarks[:3], junk[:3], m.K, m.dist)
    with pytest.raises(cv2.error):
        pnp_ref.solve_pnp_ransac_cv2(m.landmarks[:3], junk[:3], m.K, m.dist)


def test_jacobi_port_matches_cv_svdecomp():
    rng = np.random.default_rng(1)
    for n, rows in ((3, 5), (3, 11), (12, 22)):
        for _ in range(10):
            X = rng.normal(size=(rows, n)) * rng.uniform(0.1, 30, n)
            A = X.T @ X
            w, ut, _ = epnp_ref.jacobi_svd_rows(A)
            w_cv, u_cv, _ = cv2.SVDecomp(A)
            np.testing.assert_allclose(w, w_cv.ravel(), rtol=1e-10)
            np.testing.assert_allclose(ut, u_cv.T, atol=1e-9)  # same signs, not just same subspace


def test_camera_helpers_match_cv2():
    m, pts = _frames(8, seed=5)
    flat = pts.reshape(-1, 2)
    und = epnp_ref.undistort_points(flat, m.K, m.dist)
    und_cv = cv2.undistortPoints(flat.reshape(-1, 1, 2), m.K, m.dist).reshape(-1, 2)
    assert und.dtype == np.float32
    np.testing.assert_array_equal(und, und_cv)
    rv, tv = np.array([0.3, -0.5, 1.0]), np.array([0.1, -0.1, 6.0])
    R = epnp_ref.rodrigues_to_matrix(rv)
    np.testing.assert_allclose(R, cv2.Rodrigues(rv)[0], atol=1e-14)
    p = epnp_ref.project_points(m.landmarks, R, tv, m.K, m.dist)
    p_cv, _ = cv2.projectPoints(m.landmarks, rv, tv, m.K, m.dist)
    np.testing.assert_allclose(p, p_cv.reshape(-1, 2), atol=1e-9)


def test_numpy_epnp_matches_cv2_on_overdetermined_sets():
    """EPnP restated in NumPy vs cv2.solvePnP(EPNP) for n = 6..11 (the final-refit regime)."""
    m, pts = _frames(20, seed=11)
    rng = np.random.default_rng(2)
    worst_r, worst_t = 0.0, 0.0
    for b in range(len(pts)):
        n = int(rng.integers(7, 12))
        idx = np.sort(rng.choice(11, n, replace=False))
        obj = m.landmarks[idx].astype(np.float32).astype(np.float64)
        img = pts[b, idx].astype(np.float64)
        ok, rv, tv = cv2.solvePnP(obj, img, m.K, m.dist, flags=cv2.SOLVEPNP_EPNP)
        R, t = epnp_ref.solve_pnp_epnp(obj, img, m.K, m.dist)
        worst_r = max(worst_r, pnp_ref.rotation_angle_deg(R, cv2.Rodrigues(rv)[0]))
        worst_t = max(worst_t, np.linalg.norm(t - tv.ravel()) / np.linalg.norm(tv))
    assert worst_r < 1e-6 and worst_t < 1e-8, (worst_r, worst_t)


def test_quaternion_helper():
    rng = np.random.default_rng(4)
    for _ in range(50):
        rv = rng.normal(size=3) * rng.uniform(0, 3.1)
        R = epnp_ref.rodrigues_to_matrix(rv)
        q = epnp_ref.rotation_matrix_to_quat(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-12
        assert pnp_ref.rotation_angle_deg(pnp_ref.quat_to_matrix(q), R) < 1e-10


def test_confidence_filter():
    conf = np.array([0.99, 0.5, 1e-9, 0.0, -0.3, 1e-11] + [0.9] * 5, np.float32)
    good = pnp_ref.confidence_filter(conf)
    assert good.tolist() == [True, True, True, False, False, False] + [True] * 5
    assert 1.9e-10 < pnp_ref.confidence_floor(11) < 2.0e-10
    many = np.linspace(0.5, 0.99, 24).astype(np.float32)
    g = pnp_ref.confidence_filter(many)
    assert g.sum() >= 15 and g.sum() < 24  # stops as soon as 15 pass


def test_pnp_golden(pnp_golden):
    """cv2 in this environment reproduces the committed vectors (pins the cv2 build)."""
    g = pnp_golden
    assert str(g["cv2_version"]) == cv2.__version__
    ok, rv, tv, inl = pnp_ref.solve_pnp_ransac_cv2(g["landmarks"], g["e3_img"], g["K"], g["dist"], iterations=256)
    assert ok and inl.tolist() == g["e3_inliers"].tolist() == [0, 1, 2, 4, 5, 6, 7, 8, 9, 10]
    np.testing.assert_allclose(rv, g["e3_rvec"], atol=1e-12)
    np.testing.assert_allclose(rv, [0.30000040990744464, -0.5000183937489224, 0.9999999537238677], atol=1e-9)
    np.testing.assert_allclose(tv, [0.09999875404485173, -0.1000007860886065, 5.999970994017114], atol=1e-9)
    for b in range(0, len(g["kpts"]), 6):
        ok, p7, mask, rv, tv = pnp_ref.pose_from_keypoints(g["kpts"][b], g["landmarks"], g["K"], g["dist"])
        assert ok == bool(g["ok"][b]) and mask == int(g["inlier_mask"][b])
        np.testing.assert_allclose(p7, g["pose7"][b], atol=1e-10)
