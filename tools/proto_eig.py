"""Dev tool: NumPy model of the hypothesis kernel's eigen stage (csrc/ransac_epnp.cu, eig_qr_inverse_iteration).

EPnP needs the four smallest right singular directions of the 10 x 12 matrix M.  Instead of a full
one-sided Jacobi SVD (tools/proto_hyp.py), Householder QR of M^T gives the exact 2-D null space (last
two columns of Q) and block inverse iteration on R R^T gives v2, v3.  This script compares both
against cv2's per-hypothesis inlier masks (same statistics as proto_hyp.py) for several iteration
counts / block sizes.  Not part of the product or the oracle.

    NF=100 python tools/proto_eig.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, 'tools'), ROOT, os.path.join(ROOT, 'spacecraft-pose-estimation_b200')):
    sys.path.insert(0, _p)
import proto_hyp as ph  # noqa: E402

def basis_qr(M, dt, iters=8, block=2):
    """M [N,10,12] -> v [N,4,12]: v0,v1 = null space (from Householder QR of M^T), v2,v3 by block inverse iteration on R R^T."""
    N = M.shape[0]
    A = M.astype(dt).transpose(0,2,1).copy()   # [N,12,10]
    Q = np.broadcast_to(np.eye(12, dtype=dt), (N,12,12)).copy()
    for k in range(10):
        x = A[:, k:, k].copy()
        nrm = np.sqrt((x*x).sum(1)).astype(dt)
        alpha = np.where(x[:,0] >= 0, -nrm, nrm).astype(dt)
        v = x.copy(); v[:,0] = v[:,0] - alpha
        vv = (v*v).sum(1).astype(dt)
        beta = np.where(vv > 0, dt(2)/np.maximum(vv, dt(1e-30)), dt(0)).astype(dt)
        # apply to A[k:, k:]
        w = np.einsum('ni,nij->nj', v, A[:, k:, k:]).astype(dt)
        A[:, k:, k:] = (A[:, k:, k:] - (beta[:,None]*v)[:, :, None]*w[:, None, :]).astype(dt)
        wq = np.einsum('nri,ni->nr', Q[:, :, k:], v).astype(dt)
        Q[:, :, k:] = (Q[:, :, k:] - wq[:, :, None]*(beta[:,None]*v)[:, None, :]).astype(dt)
    R = np.triu(A[:, :10, :10])   # [N,10,10]
    n0, n1 = Q[:, :, 10], Q[:, :, 11]
    # block inverse iteration on B = R R^T : x <- R^-T R^-1 x
    rng = np.random.default_rng(0)
    W = np.broadcast_to(rng.normal(size=(10, block)).astype(dt), (N,10,block)).copy()
    dg = np.einsum('nii->ni', R)
    tiny = dt(1e-20)
    Rs = R.copy()
    idx = np.arange(10)
    Rs[:, idx, idx] = np.where(np.abs(dg) < tiny, np.where(dg < 0, -tiny, tiny), dg)
    def solve_upper(Rm, X):  # R a = x
        Xo = np.zeros_like(X)
        for i in range(9, -1, -1):
            acc = X[:, i] - np.einsum('nj,njb->nb', Rm[:, i, i+1:], Xo[:, i+1:])
            Xo[:, i] = (acc / Rm[:, i, i][:, None]).astype(dt)
        return Xo
    def solve_lowerT(Rm, X):  # R^T y = a
        Xo = np.zeros_like(X)
        for i in range(10):
            acc = X[:, i] - np.einsum('nj,njb->nb', Rm[:, :i, i], Xo[:, :i])
            Xo[:, i] = (acc / Rm[:, i, i][:, None]).astype(dt)
        return Xo
    def orth(W):
        for b in range(W.shape[2]):
            for a in range(b):
                d = (W[:, :, a]*W[:, :, b]).sum(1)
                W[:, :, b] = W[:, :, b] - d[:, None]*W[:, :, a]
            nr = np.sqrt((W[:, :, b]**2).sum(1))
            W[:, :, b] = W[:, :, b]/np.maximum(nr, dt(1e-30))[:, None]
        return W.astype(dt)
    with np.errstate(all='ignore'):
        W = orth(W)
        for it in range(iters):
            # normalise each col by max abs to avoid overflow between solves
            W = solve_upper(Rs, W)
            W = W/np.maximum(np.abs(W).max(1, keepdims=True), dt(1e-30))
            W = solve_lowerT(Rs, W)
            W = W/np.maximum(np.abs(W).max(1, keepdims=True), dt(1e-30))
            W = np.nan_to_num(orth(W))
        # Rayleigh-Ritz on B = R R^T: G = R^T W
        G = np.einsum('nji,njb->nib', R, W).astype(dt)   # R^T W
        S = np.einsum('nia,nib->nab', G, G).astype(np.float64)
        ev, E = np.linalg.eigh(np.nan_to_num(S))
        W = np.einsum('nib,nbc->nic', W.astype(np.float64), E).astype(dt)
    Wf = np.zeros((N,12,2), dt); Wf[:, :10] = W[:, :, :2]
    V = np.einsum('nij,njb->nib', Q, Wf).astype(dt)
    v = np.stack([n0, n1, V[:, :, 0], V[:, :, 1]], 1)
    return v

def basis_jacobi(M, dt, sweeps=8, tol=3e-7):
    A, V = ph.jacobi_cols(M, dt, sweeps, tol)
    d = (A.astype(np.float64)**2).sum(1)
    order = np.argsort(d, axis=1, kind='stable')[:, :4]
    return np.take_along_axis(V, order[:, None, :], axis=2).transpose(0,2,1)

# patch epnp_batch to use a basis function
src = open(os.path.join(ROOT, 'tools', 'proto_hyp.py')).read()
a = src.index('def epnp_batch'); b = src.index('def score(')
body = src[a:b]
body = body.replace('def epnp_batch(pw, us, fu, fv, uc, vc, dt=np.float32, sweeps=8, tol=3e-7, stats=None):', 'def epnp_batch2(pw, us, fu, fv, uc, vc, dt=np.float32, basis=None):')
i0 = body.index('    A, V = jacobi_cols'); i1 = body.index('    v4 = v.reshape')
body = body[:i0] + '    v = basis(M, dt)\n' + body[i1:]
ns = dict(np=np, PAIRS=ph.PAIRS, lsq=ph.lsq)
exec(body, ns)
epnp_batch2 = ns['epnp_batch2']

def run(basis, nframes=48, H=64, dt=np.float32, seed_off=7):
    from oracle import decode_ref, epnp_ref, ocv_rng, pnp_ref
    from spe_b200 import models, synth
    m = models.tango()
    fr = synth.make_frames(m, nframes, 64, 64, seed=synth.BASE_SEED + seed_off)
    p, mv = decode_ref.get_final_preds_fast(True, fr.heatmaps, fr.center, fr.scale)
    kp = np.concatenate([p, mv], -1)
    tot = same_cnt = same_mask = 0; win_same = pose_same = nf = 0
    t0 = time.time()
    for b in range(nframes):
        good = pnp_ref.confidence_filter(kp[b, :, 2]); n = int(good.sum())
        if n < 6: continue
        obj = m.landmarks[good]; img = kp[b, good, :2].astype(np.float32)
        tr = pnp_ref.ransac_epnp_whitebox(obj, img, m.K, m.dist, iterations=H, exhaustive=H)
        sets = ocv_rng.minimal_sets(n, H)
        obj32 = obj.astype(np.float32)
        und = epnp_ref.undistort_points(img, m.K, m.dist)
        us = np.stack([und[:, 0].astype(np.float64)*m.K[0,0]+m.K[0,2], und[:, 1].astype(np.float64)*m.K[1,1]+m.K[1,2]], 1)
        R, t = epnp_batch2(obj32[sets].astype(np.float64), us[sets], m.K[0,0], m.K[1,1], m.K[0,2], m.K[1,2], dt, basis=basis)
        counts, masks = ph.score(obj32, img, R, t, m.K, m.dist, dt)
        tot += H; same_cnt += int((counts == tr.counts[:H]).sum()); same_mask += int((masks == tr.masks[:H]).sum())
        w, _ = ocv_rng.select_sequential(counts, n, H)
        nf += 1; win_same += int(w == tr.winner)
        pose_same += int(w >= 0 and tr.winner >= 0 and masks[w] == tr.masks[tr.winner]) or int(w < 0 and tr.winner < 0)
    return f"hyp {tot}: same count {same_cnt/tot:.4f}, same mask {same_mask/tot:.4f}; frames {nf}: same winner {win_same/nf:.3f}, winner mask {pose_same/nf:.3f} ({time.time()-t0:.1f}s)"

if __name__ == '__main__':
    nfr = int(os.environ.get('NF', '100'))
    print('jacobi f32     ', run(lambda M, dt: basis_jacobi(M, dt), nfr))
    for it, blk in ((4,2),(6,2),(8,2),(12,2),(6,3),(8,3)):
        print(f'qr it={it} blk={blk}', run(lambda M, dt: basis_qr(M, dt, it, blk), nfr))
