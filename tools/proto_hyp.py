"""Dev tool: NumPy model of the GPU hypothesis kernel's numerics (batched over hypotheses).

Mirrors the plan for csrc/ransac_epnp.cu — one-sided Jacobi on the columns of the 2m x 12 matrix M
(implicit eigensolve of MtM) with V accumulated, in float32 — so algorithmic choices (sweeps,
tolerance, precision) can be checked against cv2's per-hypothesis inlier masks before spending GPU
time.  Not part of the product or the oracle.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))

PAIRS = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))


def round_robin(n=12):
    ids = list(range(n))
    rounds = []
    for _ in range(n - 1):
        rounds.append([(min(ids[i], ids[n - 1 - i]), max(ids[i], ids[n - 1 - i])) for i in range(n // 2)])
        ids = [ids[0]] + [ids[-1]] + ids[1:-1]
    return rounds


def jacobi_cols(M, dt, sweeps=8, tol=3e-7, stats=None):
    """M [N,r,12] -> (M V, V) with orthogonal columns; one-sided, round-robin order."""
    N = M.shape[0]
    A = M.astype(dt).copy()
    V = np.broadcast_to(np.eye(12, dtype=dt), (N, 12, 12)).copy()
    rr = round_robin()
    for sw in range(sweeps):
        d = (A * A).sum(1)
        nrot = 0
        for rnd in rr:
            for (i, j) in rnd:
                p = (A[:, :, i] * A[:, :, j]).sum(1)
                rot = np.abs(p) > dt(tol) * np.sqrt(d[:, i] * d[:, j])
                nrot += int(rot.sum())
                ps = np.where(rot, p, dt(1))
                zeta = (d[:, j] - d[:, i]) / (dt(2) * ps)
                t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(dt(1) + zeta * zeta))
                t = np.where(zeta == 0, dt(1), t)
                t = np.where(rot, t, dt(0)).astype(dt)
                c = (dt(1) / np.sqrt(dt(1) + t * t)).astype(dt)
                s = (c * t).astype(dt)
                ai, aj = A[:, :, i].copy(), A[:, :, j].copy()
                A[:, :, i] = c[:, None] * ai - s[:, None] * aj
                A[:, :, j] = s[:, None] * ai + c[:, None] * aj
                vi, vj = V[:, :, i].copy(), V[:, :, j].copy()
                V[:, :, i] = c[:, None] * vi - s[:, None] * vj
                V[:, :, j] = s[:, None] * vi + c[:, None] * vj
                d[:, i] = d[:, i] - t * p
                d[:, j] = d[:, j] + t * p
        if stats is not None:
            stats.append(nrot)
        if nrot == 0:
            break
    return A, V


def lsq(A, b, dt):
    q, r = np.linalg.qr(A.astype(dt))
    y = np.einsum("nij,ni->nj", q, b.astype(dt))
    # guard rank deficiency
    k = r.shape[-1]
    r = r + np.eye(k, dtype=dt) * dt(1e-30)
    return np.linalg.solve(r, y[..., None])[..., 0].astype(dt)


def epnp_batch(pw, us, fu, fv, uc, vc, dt=np.float32, sweeps=8, tol=3e-7, stats=None):
    """pw [N,m,3] float64 object points, us [N,m,2] ideal pixels. Returns R [N,3,3], t [N,3] (dt)."""
    N, m = pw.shape[:2]
    # control points and barycentric coordinates in float64 (object-side, tiny)
    c0 = pw.mean(1)
    P0 = pw - c0[:, None]
    cov = np.einsum("nik,nil->nkl", P0, P0)
    w, U = np.linalg.eigh(cov)
    w, U = w[:, ::-1], U[:, :, ::-1]
    k = np.sqrt(np.maximum(w, 0) / m)
    cws = np.concatenate([c0[:, None], c0[:, None] + (k[:, None, :] * U).transpose(0, 2, 1)], 1)  # [N,4,3]
    al = np.zeros((N, m, 4))
    al[:, :, 1:] = np.einsum("nik,nkj->nij", P0, U) / k[:, None, :]
    al[:, :, 0] = 1 - al[:, :, 1:].sum(2)
    al = al.astype(dt)
    us = us.astype(dt)
    pwd = pw.astype(dt)
    M = np.zeros((N, 2 * m, 12), dt)
    for j in range(4):
        M[:, 0::2, 3 * j] = al[:, :, j] * dt(fu)
        M[:, 0::2, 3 * j + 2] = al[:, :, j] * (dt(uc) - us[:, :, 0])
        M[:, 1::2, 3 * j + 1] = al[:, :, j] * dt(fv)
        M[:, 1::2, 3 * j + 2] = al[:, :, j] * (dt(vc) - us[:, :, 1])
    A, V = jacobi_cols(M, dt, sweeps, tol, stats)
    d = (A.astype(np.float64) ** 2).sum(1)
    order = np.argsort(d, axis=1, kind="stable")[:, :4]
    v = np.take_along_axis(V, order[:, None, :], axis=2).transpose(0, 2, 1)  # [N,4,12]
    v4 = v.reshape(N, 4, 4, 3)
    dv = np.stack([v4[:, :, a] - v4[:, :, b] for a, b in PAIRS], 2)  # [N,4,6,3]
    dot = lambda i, j: (dv[:, i] * dv[:, j]).sum(-1)  # noqa: E731  [N,6]
    L = np.stack([dot(0, 0), 2 * dot(0, 1), dot(1, 1), 2 * dot(0, 2), 2 * dot(1, 2), dot(2, 2), 2 * dot(0, 3), 2 * dot(1, 3),
                  2 * dot(2, 3), dot(3, 3)], -1).astype(dt)  # [N,6,10]
    cw = cws.astype(dt)
    rho = np.stack([((cw[:, a] - cw[:, b]) ** 2).sum(-1) for a, b in PAIRS], -1).astype(dt)

    def approx(Nv):
        if Nv == 1:
            b4 = lsq(L[:, :, [0, 1, 3, 6]], rho, dt)
            sg = np.where(b4[:, 0] < 0, dt(-1), dt(1))
            b0 = np.sqrt(sg * b4[:, 0])
            return np.stack([b0, sg * b4[:, 1] / b0, sg * b4[:, 2] / b0, sg * b4[:, 3] / b0], -1)
        cols = [0, 1, 2] if Nv == 2 else [0, 1, 2, 3, 4]
        bb = lsq(L[:, :, cols], rho, dt)
        neg = bb[:, 0] < 0
        b0 = np.sqrt(np.abs(bb[:, 0]))
        b1 = np.where(neg, np.where(bb[:, 2] < 0, np.sqrt(np.abs(bb[:, 2])), 0), np.where(bb[:, 2] > 0, np.sqrt(np.abs(bb[:, 2])), 0))
        b0 = np.where(bb[:, 1] < 0, -b0, b0)
        b2 = bb[:, 3] / b0 if Nv == 3 else np.zeros_like(b0)
        return np.stack([b0, b1, b2, np.zeros_like(b0)], -1).astype(dt)

    def gn(be):
        be = be.astype(dt).copy()
        for _ in range(5):
            A_ = np.stack([
                2 * L[:, :, 0] * be[:, None, 0] + L[:, :, 1] * be[:, None, 1] + L[:, :, 3] * be[:, None, 2] + L[:, :, 6] * be[:, None, 3],
                L[:, :, 1] * be[:, None, 0] + 2 * L[:, :, 2] * be[:, None, 1] + L[:, :, 4] * be[:, None, 2] + L[:, :, 7] * be[:, None, 3],
                L[:, :, 3] * be[:, None, 0] + L[:, :, 4] * be[:, None, 1] + 2 * L[:, :, 5] * be[:, None, 2] + L[:, :, 8] * be[:, None, 3],
                L[:, :, 6] * be[:, None, 0] + L[:, :, 7] * be[:, None, 1] + L[:, :, 8] * be[:, None, 2] + 2 * L[:, :, 9] * be[:, None, 3],
            ], -1).astype(dt)
            bb = np.stack([be[:, 0] * be[:, 0], be[:, 0] * be[:, 1], be[:, 1] * be[:, 1], be[:, 0] * be[:, 2], be[:, 1] * be[:, 2],
                           be[:, 2] * be[:, 2], be[:, 0] * be[:, 3], be[:, 1] * be[:, 3], be[:, 2] * be[:, 3], be[:, 3] * be[:, 3]], -1)
            res = (rho - np.einsum("nij,nj->ni", L, bb)).astype(dt)
            with np.errstate(all="ignore"):
                x = lsq(A_, res, dt)
            be = (be + np.nan_to_num(x)).astype(dt)
        return be

    def pose(be):
        ccs = np.einsum("nk,nkj->nj", be, v).reshape(N, 4, 3).astype(dt)
        pcs = np.einsum("nij,njk->nik", al, ccs).astype(dt)
        flip = np.where(pcs[:, 0, 2] < 0, dt(-1), dt(1))
        pcs = pcs * flip[:, None, None]
        pc0, pw0 = pcs.mean(1), pwd.mean(1)
        abt = np.einsum("nik,nil->nkl", pcs - pc0[:, None], pwd - pw0[:, None]).astype(dt)
        with np.errstate(all="ignore"):
            abt = np.nan_to_num(abt)
            Uu, _, Vt = np.linalg.svd(abt)
        R = Uu @ Vt
        neg = np.linalg.det(R) < 0
        R[neg, 2] = -R[neg, 2]
        t = pc0 - np.einsum("nij,nj->ni", R, pw0)
        pc = np.einsum("nij,nkj->nki", R, pwd) + t[:, None]
        with np.errstate(all="ignore"):
            ue = dt(uc) + dt(fu) * pc[:, :, 0] / pc[:, :, 2]
            ve = dt(vc) + dt(fv) * pc[:, :, 1] / pc[:, :, 2]
            err = np.sqrt((us[:, :, 0] - ue) ** 2 + (us[:, :, 1] - ve) ** 2).sum(1) / m
        err = np.where(np.isfinite(err), err, np.inf)
        return R.astype(dt), t.astype(dt), err

    with np.errstate(all="ignore"):
        sols = [pose(gn(approx(Nv))) for Nv in (1, 2, 3)]
    best = np.zeros(N, np.int64)
    best = np.where(sols[1][2] < sols[0][2], 1, best)
    e_best = np.where(best == 1, sols[1][2], sols[0][2])
    best = np.where(sols[2][2] < e_best, 2, best)
    R = np.choose(best[:, None, None], [s[0] for s in sols])
    t = np.choose(best[:, None], [s[1] for s in sols])
    return R, t


def score(obj32, img32, R, t, K, dist, dt=np.float32, reproj=15.0):
    """obj32 [n,3], img32 [n,2] float32; R [N,3,3], t [N,3] -> counts [N], masks [N] uint32."""
    k1, k2, p1, p2, k3 = (dt(v) for v in dist)
    R = R.astype(dt)
    t = t.astype(dt)
    pc = np.einsum("nij,kj->nki", R, obj32.astype(dt)) + t[:, None]
    with np.errstate(all="ignore"):
        x, y = pc[..., 0] / pc[..., 2], pc[..., 1] / pc[..., 2]
        r2 = x * x + y * y
        cd = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2
        xd = x * cd + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        yd = y * cd + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        u = (dt(K[0, 0]) * xd + dt(K[0, 2])).astype(np.float32)
        v = (dt(K[1, 1]) * yd + dt(K[1, 2])).astype(np.float32)
        du, dv = img32[None, :, 0] - u, img32[None, :, 1] - v
        err = du * du + dv * dv
    good = err <= np.float32(reproj * reproj)
    counts = good.sum(1)
    masks = (good.astype(np.uint64) << np.arange(obj32.shape[0], dtype=np.uint64)[None]).sum(1).astype(np.uint32)
    return counts, masks


def main():
    import cv2  # noqa: F401

    from oracle import decode_ref, epnp_ref, ocv_rng, pnp_ref
    from spe_b200 import models, synth

    dt = np.float32 if "f64" not in sys.argv else np.float64
    nframes = int(os.environ.get("NF", "48"))
    H = 64
    m = models.tango()
    fr = synth.make_frames(m, nframes, 64, 64, seed=synth.BASE_SEED + 7)
    p, mv = decode_ref.get_final_preds_fast(True, fr.heatmaps, fr.center, fr.scale)
    kp = np.concatenate([p, mv], -1)
    tot = same_cnt = same_mask = 0
    win_same = pose_same = nf = 0
    sweeps_stats = []
    t0 = time.time()
    for b in range(nframes):
        good = pnp_ref.confidence_filter(kp[b, :, 2])
        n = int(good.sum())
        if n < 6:
            continue
        obj = m.landmarks[good]
        img = kp[b, good, :2].astype(np.float32)
        tr = pnp_ref.ransac_epnp_whitebox(obj, img, m.K, m.dist, iterations=H, exhaustive=H)
        sets = ocv_rng.minimal_sets(n, H)
        obj32 = obj.astype(np.float32)
        und = epnp_ref.undistort_points(img, m.K, m.dist)  # float32
        us = np.stack([und[:, 0].astype(np.float64) * m.K[0, 0] + m.K[0, 2], und[:, 1].astype(np.float64) * m.K[1, 1] + m.K[1, 2]], 1)
        st = []
        R, t = epnp_batch(obj32[sets].astype(np.float64), us[sets], m.K[0, 0], m.K[1, 1], m.K[0, 2], m.K[1, 2], dt, stats=st)
        sweeps_stats.append(len(st))
        counts, masks = score(obj32, img, R, t, m.K, m.dist, dt)
        tot += H
        same_cnt += int((counts == tr.counts[:H]).sum())
        same_mask += int((masks == tr.masks[:H]).sum())
        w, _ = ocv_rng.select_sequential(counts, n, H)
        nf += 1
        win_same += int(w == tr.winner)
        pose_same += int(w >= 0 and tr.winner >= 0 and masks[w] == tr.masks[tr.winner]) or int(w < 0 and tr.winner < 0)
    print(f"dtype {dt.__name__}: hypotheses {tot}, same count {same_cnt/tot:.4f}, same mask {same_mask/tot:.4f}; "
          f"frames {nf}: same winner {win_same/nf:.3f}, same winner mask {pose_same/nf:.3f}; sweeps {np.bincount(sweeps_stats)} ({time.time()-t0:.1f}s)")


if __name__ == "__main__":
    main()
