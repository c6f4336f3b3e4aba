"""Oracle (test infrastructure): EPnP and its helpers restated in NumPy float64.

The reference calls cv2.solvePnPRansac(..., flags=cv2.SOLVEPNP_EPNP)
(pose_estimation/export_predicted_poses_real.py:199-201).  The arithmetic is OpenCV calib3d's,
a third-party dependency that is not vendored under /root/reference (opencv-python==3.4.11.41
pinned in environment.yml:37; this image has cv2 4.13.0).  This file restates the published
algorithm (Lepetit/Moreno-Noguer/Fua EPnP as implemented by OpenCV) so the
CUDA kernels have a white-box to be compared with step by step; tests/test_oracle_pnp.py pins it
against cv2.solvePnP(EPNP), cv2.undistortPoints, cv2.projectPoints, cv2.Rodrigues and cv2.SVDecomp.

Sign-defining detail: OpenCV's SVD is a one-sided (Hestenes) Jacobi; the signs of the PCA axes
that place EPnP's control points come out of it, so `jacobi_svd_rows` ports it rotation by rotation.
"""
from __future__ import annotations

import numpy as np

_PAIRS = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


# ----------------------------------------------------------------------------- OpenCV Jacobi SVD
def jacobi_svd_rows(A: np.ndarray):
    """One-sided Jacobi SVD as OpenCV runs it for cvSVD(A, W, Ut, 0, MODIFY_A | U_T) on a square
    matrix (App. B.4).  Works on the rows of A^T; returns (w descending, Ut, Vt) where row i of Ut
    is the i-th left singular vector.  float64 only."""
    At = np.array(A, dtype=np.float64).T.copy()
    n = At.shape[0]
    Vt = np.eye(n)
    W = (At * At).sum(axis=1)
    eps = np.finfo(np.float64).eps * 10
    for _ in range(max(n, 30)):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= eps * np.sqrt(a * b):
                    continue
                p *= 2.0
                beta = a - b
                gamma = np.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = np.sqrt(delta / gamma)
                    c = p / (gamma * s * 2.0)
                else:
                    c = np.sqrt((gamma + beta) / (gamma * 2.0))
                    s = p / (gamma * c * 2.0)
                ti = c * At[i] + s * At[j]
                tj = -s * At[i] + c * At[j]
                At[i], At[j] = ti, tj
                W[i], W[j] = ti @ ti, tj @ tj
                vi = c * Vt[i] + s * Vt[j]
                vj = -s * Vt[i] + c * Vt[j]
                Vt[i], Vt[j] = vi, vj
                changed = True
        if not changed:
            break
    w = np.sqrt((At * At).sum(axis=1))
    for i in range(n - 1):  # selection sort, descending (ties keep the earlier row)
        j = i + int(np.argmax(w[i:]))
        if w[j] > w[i]:
            w[[i, j]] = w[[j, i]]
            At[[i, j]] = At[[j, i]]
            Vt[[i, j]] = Vt[[j, i]]
    Ut = At / np.where(w > np.finfo(np.float64).tiny, w, np.inf)[:, None]
    return w, Ut, Vt


def _svd_solve(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """cv::solve(A, b, x, DECOMP_SVD): minimum-norm least squares; singular values at or below
    2*DBL_EPSILON*sum(w) are dropped."""
    U, w, Vt = np.linalg.svd(A, full_matrices=False)
    thr = np.finfo(np.float64).eps * 2 * w.sum()
    winv = np.where(w > thr, 1.0 / np.where(w > thr, w, 1.0), 0.0)
    return Vt.T @ (winv * (U.T @ b))


# ----------------------------------------------------------------------------- camera helpers
def undistort_points(uv: np.ndarray, K: np.ndarray, dist: np.ndarray, iters: int = 5, out_dtype=None) -> np.ndarray:
    """cv2.undistortPoints(uv, K, dist) without R/P: normalised coordinates after exactly 5
    fixed-point iterations (App. B.3a).  The result has the input's dtype, as in OpenCV."""
    uv = np.asarray(uv)
    out_dtype = out_dtype or (np.float32 if uv.dtype == np.float32 else np.float64)
    k1, k2, p1, p2, k3 = (float(v) for v in np.asarray(dist, np.float64).ravel()[:5])
    x0 = (uv[..., 0].astype(np.float64) - K[0, 2]) / K[0, 0]
    y0 = (uv[..., 1].astype(np.float64) - K[1, 2]) / K[1, 1]
    x, y = x0.copy(), y0.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icd = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
        dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
        x = (x0 - dx) * icd
        y = (y0 - dy) * icd
    return np.stack([x, y], axis=-1).astype(out_dtype)


def project_points(obj: np.ndarray, R: np.ndarray, t: np.ndarray, K: np.ndarray, dist: np.ndarray) -> np.ndarray:
    """cv2.projectPoints in float64 (App. B.5)."""
    k1, k2, p1, p2, k3 = (float(v) for v in np.asarray(dist, np.float64).ravel()[:5])
    pc = np.asarray(obj, np.float64) @ np.asarray(R, np.float64).T + np.asarray(t, np.float64).reshape(1, 3)
    x, y = pc[:, 0] / pc[:, 2], pc[:, 1] / pc[:, 2]
    r2 = x * x + y * y
    cd = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
    xd = x * cd + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
    yd = y * cd + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
    return np.stack([K[0, 0] * xd + K[0, 2], K[1, 1] * yd + K[1, 2]], axis=1)


def rodrigues_to_matrix(rvec: np.ndarray) -> np.ndarray:
    r = np.asarray(rvec, np.float64).ravel()
    th = np.linalg.norm(r)
    if th < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * Kx


def rotation_matrix_to_quat(r: np.ndarray) -> np.ndarray:
    """cv_rotation_matrix_to_quat, pose_estimation/export_predicted_poses_real.py:22-57:
    scalar-first quaternion, branch on the largest of the four candidate magnitudes."""
    r = np.asarray(r, np.float64)
    tr = [r[0, 0] + r[1, 1] + r[2, 2], r[0, 0] - r[1, 1] - r[2, 2], -r[0, 0] + r[1, 1] - r[2, 2], -r[0, 0] - r[1, 1] + r[2, 2]]
    e = [np.sqrt(max(1.0 + v, 0.0)) / 2.0 for v in tr]
    m = int(np.argmax(e))
    if m == 0:
        q = [e[0], (r[2, 1] - r[1, 2]) / (4 * e[0]), (r[0, 2] - r[2, 0]) / (4 * e[0]), (r[1, 0] - r[0, 1]) / (4 * e[0])]
    elif m == 1:
        q = [(r[2, 1] - r[1, 2]) / (4 * e[1]), e[1], (r[1, 0] + r[0, 1]) / (4 * e[1]), (r[2, 0] + r[0, 2]) / (4 * e[1])]
    elif m == 2:
        q = [(r[0, 2] - r[2, 0]) / (4 * e[2]), (r[1, 0] + r[0, 1]) / (4 * e[2]), e[2], (r[2, 1] + r[1, 2]) / (4 * e[2])]
    else:
        q = [(r[1, 0] - r[0, 1]) / (4 * e[3]), (r[2, 0] + r[0, 2]) / (4 * e[3]), (r[2, 1] + r[1, 2]) / (4 * e[3]), e[3]]
    return np.array(q)


# ----------------------------------------------------------------------------- EPnP
def _householder_lsq(A: np.ndarray, b: np.ndarray) -> np.ndarray:
    """The 6x4 least-squares step of EPnP's Gauss-Newton (OpenCV solves it with a hand-written
    Householder QR).  Any backward-stable QR gives the same answer to rounding."""
    q, r = np.linalg.qr(A)
    return np.linalg.solve(r, q.T @ b)


def epnp(obj: np.ndarray, uv_norm: np.ndarray, K: np.ndarray, debug: dict | None = None):
    """EPnP as OpenCV runs it inside solvePnP(EPNP) (App. B.3 b-k).

    obj      [n,3]  object points (already float32-rounded if the caller mimics cv2)
    uv_norm  [n,2]  *undistorted normalised* image points (output of undistort_points)
    K        [3,3]  intrinsics; EPnP works in pixel units, x*fx+cx, with K inside M
    Returns (R [3,3], t [3]).
    """
    pw = np.asarray(obj, np.float64)
    n = pw.shape[0]
    fu, fv, uc, vc = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    us = np.stack([np.asarray(uv_norm[:, 0], np.float64) * fu + uc, np.asarray(uv_norm[:, 1], np.float64) * fv + vc], axis=1)

    # control points: centroid + PCA axes scaled by sqrt(lambda / n)                      (B.3c)
    cws = np.zeros((4, 3))
    cws[0] = pw.sum(axis=0) / n
    pw0 = pw - cws[0]
    dc, uct, _ = jacobi_svd_rows(pw0.T @ pw0)
    for i in range(3):
        cws[i + 1] = cws[0] + np.sqrt(dc[i] / n) * uct[i]

    # barycentric coordinates                                                            (B.3d)
    cc = (cws[1:] - cws[0]).T
    cc_inv = np.linalg.pinv(cc)
    alphas = np.zeros((n, 4))
    alphas[:, 1:] = pw0 @ cc_inv.T
    alphas[:, 0] = 1.0 - alphas[:, 1] - alphas[:, 2] - alphas[:, 3]

    # M (2n x 12) in pixel units                                                         (B.3e)
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = alphas[:, j] * fu
        M[0::2, 3 * j + 2] = alphas[:, j] * (uc - us[:, 0])
        M[1::2, 3 * j + 1] = alphas[:, j] * fv
        M[1::2, 3 * j + 2] = alphas[:, j] * (vc - us[:, 1])

    # the four right-most singular vectors of MtM                                        (B.3f)
    _, ut, _ = jacobi_svd_rows(M.T @ M)
    v = [ut[11], ut[10], ut[9], ut[8]]

    # L (6x10) and rho                                                                   (B.3g)
    dv = np.array([[v[i][3 * a:3 * a + 3] - v[i][3 * b:3 * b + 3] for a, b in _PAIRS] for i in range(4)])
    L = np.zeros((6, 10))
    for k in range(6):
        d0, d1, d2, d3 = dv[0, k], dv[1, k], dv[2, k], dv[3, k]
        L[k] = [d0 @ d0, 2 * d0 @ d1, d1 @ d1, 2 * d0 @ d2, 2 * d1 @ d2, d2 @ d2, 2 * d0 @ d3, 2 * d1 @ d3, 2 * d2 @ d3, d3 @ d3]
    rho = np.array([np.sum((cws[a] - cws[b]) ** 2) for a, b in _PAIRS])

    def approx(N):  # three linearised initialisations                                    (B.3h)
        if N == 1:
            b4 = _svd_solve(L[:, [0, 1, 3, 6]], rho)
            sg = -1.0 if b4[0] < 0 else 1.0
            b0 = np.sqrt(sg * b4[0])
            return np.array([b0, sg * b4[1] / b0, sg * b4[2] / b0, sg * b4[3] / b0])
        cols = [0, 1, 2] if N == 2 else [0, 1, 2, 3, 4]
        bb = _svd_solve(L[:, cols], rho)
        if bb[0] < 0:
            b0 = np.sqrt(-bb[0])
            b1 = np.sqrt(-bb[2]) if bb[2] < 0 else 0.0
        else:
            b0 = np.sqrt(bb[0])
            b1 = np.sqrt(bb[2]) if bb[2] > 0 else 0.0
        if bb[1] < 0:
            b0 = -b0
        return np.array([b0, b1, (bb[3] / b0) if N == 3 else 0.0, 0.0])

    def gauss_newton(be):  # exactly five iterations                                      (B.3i)
        be = be.copy()
        for _ in range(5):
            A = np.stack(
                [
                    2 * L[:, 0] * be[0] + L[:, 1] * be[1] + L[:, 3] * be[2] + L[:, 6] * be[3],
                    L[:, 1] * be[0] + 2 * L[:, 2] * be[1] + L[:, 4] * be[2] + L[:, 7] * be[3],
                    L[:, 3] * be[0] + L[:, 4] * be[1] + 2 * L[:, 5] * be[2] + L[:, 8] * be[3],
                    L[:, 6] * be[0] + L[:, 7] * be[1] + L[:, 8] * be[2] + 2 * L[:, 9] * be[3],
                ],
                axis=1,
            )
            bb = np.array([be[0] * be[0], be[0] * be[1], be[1] * be[1], be[0] * be[2], be[1] * be[2], be[2] * be[2],
                           be[0] * be[3], be[1] * be[3], be[2] * be[3], be[3] * be[3]])
            be = be + _householder_lsq(A, rho - L @ bb)
        return be

    def pose_from_betas(be):  # camera-frame control points -> Procrustes               (B.3j)
        ccs = sum(be[k] * v[k].reshape(4, 3) for k in range(4))
        pcs = alphas @ ccs
        if pcs[0, 2] < 0:
            ccs, pcs = -ccs, -pcs
        pc0, pwc = pcs.sum(axis=0) / n, pw.sum(axis=0) / n
        abt = (pcs - pc0).T @ (pw - pwc)
        U, _, Vt = np.linalg.svd(abt)
        R = U @ Vt
        if np.linalg.det(R) < 0:
            R[2] = -R[2]  # OpenCV flips the third ROW of R, not a column of U
        t = pc0 - R @ pwc
        pc = pw @ R.T + t
        ue = uc + fu * pc[:, 0] / pc[:, 2]
        ve = vc + fv * pc[:, 1] / pc[:, 2]
        err = np.sum(np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2)) / n
        return R, t, err

    sols = [pose_from_betas(gauss_newton(approx(N))) for N in (1, 2, 3)]
    best = 0  # (B.3k)
    if sols[1][2] < sols[0][2]:
        best = 1
    if sols[2][2] < sols[best][2]:
        best = 2
    if debug is not None:
        debug.update(cws=cws, alphas=alphas, M=M, v=v, L=L, rho=rho, errs=[s[2] for s in sols], N=best + 1)
    return sols[best][0], sols[best][1]


def solve_pnp_epnp(obj, img, K, dist):
    """cv2.solvePnP(obj, img, K, dist, flags=SOLVEPNP_EPNP) restated: undistort (5 iterations, in
    the dtype of `img`), then EPnP in pixel units.  Returns (R, t)."""
    img = np.asarray(img)
    und = undistort_points(img.reshape(-1, 2), K, dist)
    return epnp(np.asarray(obj, np.float64).reshape(-1, 3), und.astype(np.float64), K)
