"""Batched RANSAC-EPnP on the B200: the pose half of the stage.

Mirrors the reference's per-frame loop, pose_estimation/export_predicted_poses_real.py:177-204:
    confidence filter (:186-197) -> cv2.solvePnPRansac(landmarks[good], pts[good], K, distCoeffs=dist,
    flags=cv2.SOLVEPNP_EPNP, iterationsCount=10000, reprojectionError=15.0) (:199-201)
    -> cv2.Rodrigues (:203) -> cv_rotation_matrix_to_quat (:22-57)
for a whole batch of frames at once.  `PnPSolver.solve` takes the `pred.mat` layout the reference
passes between its stages (preds float32 [N,J,3] = x, y, maxval; lib/dataset/PEdataset.py:121-123)
as a NumPy array or a torch CUDA tensor; `solvePnPRansac` keeps cv2's single-frame signature and
return tuple for call-site compatibility.

Two selections are available (include/spe_b200.h):
  exact=True (default)  cv2's own sequential, adaptive loop replayed in float64 up to `max_hypotheses`
                        (= cv2's iterationsCount, 10000 in the reference): the parity path.
  exact=False           every distinct one of the first `hypotheses` minimal sets of OpenCV's fixed-seed RNG is
                        scored in FP32 and cv2's acceptance rule is replayed over those counts: equal to cv2
                        whenever FP32 and float64 agree on the hypotheses cv2 looks at and cv2 stops within
                        `hypotheses` draws (`PoseBatch.budget` tells when it would not have).
All arithmetic is in libspe_b200.so (csrc/ransac_*.cu); there is no CPU path here.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

from . import _lib

SOLVEPNP_EPNP = 1  # cv2.SOLVEPNP_EPNP
FRAME_OK, FRAME_TOO_FEW_POINTS, FRAME_P3P_UNSUPPORTED, FRAME_NO_MODEL = 0, 1, 2, 3
ADAPTIVE_CONFIDENCE_FILTER = -1.0  # conf_floor value that selects the reference's 0.95*0.8^k filter


@dataclass
class PoseBatch:
    pose7: object  # [B,7] float32 (qw,qx,qy,qz,tx,ty,tz)
    inlier_mask: object  # [B] int32 holding the uint32 bit mask over the J landmarks
    status: object  # [B] int32, FRAME_*
    winner: object  # [B] int32 accepted hypothesis index (-1: none)
    rt: object  # [B,12] float64 row-major R then t
    budget: object = None  # [B] int32: hypotheses cv2's loop looks at (exact) / would still want (fast: > hypotheses = cut short)


def _dptr(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


class PnPSolver:
    """Immutable landmark/camera model on one device + reusable scratch space."""

    def __init__(self, landmarks, K, dist=None, max_hypotheses: int = 10000, device=None):
        """max_hypotheses = cv2's iterationsCount (10000 in the reference's call): the budget cv2's loop starts with
        and the upper bound of `hypotheses` in solve()."""
        torch = _lib.require_cuda()
        self._L = _lib.lib()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lm = np.ascontiguousarray(landmarks, dtype=np.float64).reshape(-1, 3)
        Km = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
        self.J = int(lm.shape[0])
        self.max_hypotheses = int(max_hypotheses)
        dp = None
        if dist is not None:
            d = np.zeros(5)
            dv = np.asarray(dist, np.float64).ravel()
            if dv.size > 5 and np.any(dv[5:] != 0):
                raise ValueError("only the 5-coefficient distortion model (k1,k2,p1,p2,k3) is supported")
            d[: min(5, dv.size)] = dv[:5]
            dp = _dptr(d, ctypes.c_double)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._L.spe_pnp_model_create(_dptr(lm, ctypes.c_double), self.J, _dptr(Km, ctypes.c_double), dp,
                                                    self.max_hypotheses, ctypes.byref(handle)), "spe_pnp_model_create")
        self._handle = handle
        self._workspace = None
        self._ws_key = None

    # -- lifetime
    def close(self):
        if getattr(self, "_handle", None):
            self._L.spe_pnp_model_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._handle

    def minimal_sets(self, n: int, count: int) -> np.ndarray:
        """First `count` 5-point minimal sets OpenCV's RANSAC draws for n points (int32 [count,5])."""
        out = np.empty((count, 5), np.int32)
        _lib.check(self._L.spe_pnp_model_minimal_sets(self._handle, int(n), int(count), _dptr(out, ctypes.c_int32)),
                   "spe_pnp_model_minimal_sets")
        return out

    def workspace(self, B: int, hypotheses: int):
        torch = _lib.require_cuda()
        need = int(self._L.spe_ransac_workspace_bytes(self._handle, B, hypotheses))
        if self._workspace is None or self._workspace.numel() < need or self._ws_key != (B, hypotheses):
            if self._workspace is None or self._workspace.numel() < need:
                self._workspace = torch.empty(max(need, 16), dtype=torch.uint8, device=self.device)
            self._ws_key = (B, hypotheses)
        return self._workspace

    # -- the batched solve
    def solve_device(self, kpts, hypotheses: int = 256, reproj_err: float = 15.0, confidence: float = 0.99,
                     conf_floor: float = ADAPTIVE_CONFIDENCE_FILTER, want_rt: bool = True, refine: str | None = None,
                     adaptive: bool = False, eig: str = "qr", exact: bool = True, want_budget: bool = True) -> PoseBatch:
        """kpts [B,J,3] float32 CUDA contiguous -> PoseBatch of CUDA tensors.  Enqueues on torch's
        current stream and does not synchronise.  `hypotheses` minimal sets per frame are scored in FP32 (0 with
        exact=True: none); exact=True selects by the float64 replay of cv2's loop.  eig="jacobi" (development builds
        of the library only) scores the hypotheses with the full Jacobi SVD of M^T.  refine="lm" adds a
        reprojection-error Levenberg-Marquardt step on the inliers (cv2.solvePnPRefineLM's result); the reference does
        not do that, so it is off by default."""
        torch = _lib.require_cuda()
        if refine not in (None, "lm"):
            raise ValueError("refine must be None or 'lm'")
        if eig not in ("qr", "jacobi"):
            raise ValueError("eig must be 'qr' or 'jacobi'")
        assert kpts.is_cuda and kpts.dtype == torch.float32 and kpts.is_contiguous()
        B, J, three = kpts.shape
        if J != self.J or three != 3:
            raise ValueError(f"kpts must be [B,{self.J},3]")
        hypotheses = min(int(hypotheses), self.max_hypotheses) if exact else int(hypotheses)
        if not ((0 if exact else 1) <= hypotheses <= self.max_hypotheses):
            raise ValueError(f"hypotheses must be in [{0 if exact else 1}, {self.max_hypotheses}]")
        dev = kpts.device
        pose7 = torch.empty((B, 7), dtype=torch.float32, device=dev)
        mask = torch.empty((B,), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        winner = torch.empty((B,), dtype=torch.int32, device=dev)
        rt = torch.empty((B, 12), dtype=torch.float64, device=dev) if want_rt else None
        budget = torch.empty((B,), dtype=torch.int32, device=dev) if want_budget else None
        ws = self.workspace(B, hypotheses)
        stream = torch.cuda.current_stream(dev).cuda_stream
        flags = ((_lib.FLAG_REFINE_LM if refine == "lm" else 0) | (_lib.FLAG_ADAPTIVE if adaptive else 0) |
                 (_lib.FLAG_JACOBI_SVD if eig == "jacobi" else 0) | (_lib.FLAG_EXACT if exact else 0))
        with torch.cuda.device(dev):
            _lib.check(self._L.spe_ransac_epnp_f32(self._handle, kpts.data_ptr(), B, int(hypotheses), float(reproj_err),
                                                   float(confidence), float(conf_floor), pose7.data_ptr(), mask.data_ptr(),
                                                   status.data_ptr(), winner.data_ptr(), rt.data_ptr() if want_rt else None,
                                                   ws.data_ptr(), ws.numel(), flags, stream),
                       "spe_ransac_epnp_f32")
            if want_budget:
                _lib.check(self._L.spe_ransac_read_budget(self._handle, ws.data_ptr(), B, int(hypotheses), budget.data_ptr(), stream),
                           "spe_ransac_read_budget")
        return PoseBatch(pose7, mask, status, winner, rt, budget)

    def solve(self, kpts, **kw) -> PoseBatch:
        """NumPy in -> NumPy out, torch CUDA in -> torch CUDA out."""
        torch = _lib.require_cuda()
        if isinstance(kpts, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(kpts, np.float32)).to(self.device)
            out = self.solve_device(t, **kw)
            return PoseBatch(*[None if x is None else x.cpu().numpy() for x in (out.pose7, out.inlier_mask, out.status, out.winner, out.rt, out.budget)])
        return self.solve_device(kpts.to(self.device, torch.float32).contiguous(), **kw)

    def hypothesis_scores(self, B: int, hypotheses: int):
        """(counts int32 [B,H], masks int32 [B,H]) of the most recent solve on this solver
        (parity tests only)."""
        torch = _lib.require_cuda()
        counts = torch.empty((B, hypotheses), dtype=torch.int32, device=self.device)
        masks = torch.empty((B, hypotheses), dtype=torch.int32, device=self.device)
        ws = self.workspace(B, hypotheses)
        with torch.cuda.device(self.device):
            _lib.check(self._L.spe_ransac_debug_scores(self._handle, ws.data_ptr(), B, hypotheses, counts.data_ptr(), masks.data_ptr(),
                                                       torch.cuda.current_stream(self.device).cuda_stream), "spe_ransac_debug_scores")
        return counts, masks


def matrix_to_rvec(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> axis-angle (what cv2.Rodrigues(R) returns), float64 [3]."""
    R = np.asarray(R, np.float64)
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.linalg.norm(w)
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    theta = np.arctan2(s, c)
    if s < 1e-12:
        if c > 0:
            return w  # theta ~ 0
        # theta ~ pi: axis from the diagonal
        axis = np.sqrt(np.maximum((np.diag(R) + 1.0) * 0.5, 0.0))
        if R[0, 1] < 0:
            axis[1] = -axis[1]
        if R[0, 2] < 0:
            axis[2] = -axis[2]
        return axis / max(np.linalg.norm(axis), 1e-300) * theta
    return w / s * theta


_solver_cache: dict = {}


def solvePnPRansac(objectPoints, imagePoints, cameraMatrix, distCoeffs=None, flags=SOLVEPNP_EPNP, iterationsCount=100,
                   reprojectionError=8.0, confidence=0.99):
    """cv2.solvePnPRansac's signature and return tuple for one frame, computed on the GPU.

    Returns (ret, rvec (3,1) float64, tvec (3,1) float64, inliers (k,1) int32 or None).
    Raises ValueError for fewer than 4 points (cv2 raises cv2.error); exactly 4 points take cv2's P3P branch.  `iterationsCount` is capped at 16384.  The result is cv2's loop replayed
    in float64 (exact=True, no FP32 scoring).
    """
    if flags != SOLVEPNP_EPNP:
        raise NotImplementedError("only flags=cv2.SOLVEPNP_EPNP is implemented (the reference's setting)")
    obj = np.ascontiguousarray(objectPoints, np.float64).reshape(-1, 3)
    img = np.ascontiguousarray(imagePoints, np.float32).reshape(-1, 2)
    n = obj.shape[0]
    if n != img.shape[0]:
        raise ValueError("objectPoints and imagePoints need the same number of points")
    if n < 4:
        raise ValueError("solvePnPRansac needs at least 4 points")
    if n > 32:
        raise ValueError("at most 32 points per frame")
    H = int(min(max(iterationsCount, 1), _lib.MAX_HYPOTHESES))
    K = np.ascontiguousarray(cameraMatrix, np.float64)
    d = None if distCoeffs is None else np.asarray(distCoeffs, np.float64).ravel()
    key = (obj.tobytes(), K.tobytes(), None if d is None else d.tobytes(), H)
    solver = _solver_cache.get(key)
    if solver is None:
        if len(_solver_cache) > 16:
            _solver_cache.clear()
        solver = _solver_cache[key] = PnPSolver(obj, K, d, max_hypotheses=H)
    kpts = np.concatenate([img, np.ones((n, 1), np.float32)], axis=1)[None]
    out = solver.solve(kpts, hypotheses=0, exact=True, reproj_err=float(reprojectionError), confidence=float(confidence), conf_floor=0.5)
    ok = int(out.status[0]) == FRAME_OK
    rt = out.rt[0]
    rvec = matrix_to_rvec(rt[:9].reshape(3, 3)).reshape(3, 1) if ok else np.zeros((3, 1))
    tvec = rt[9:].reshape(3, 1).copy() if ok else np.zeros((3, 1))
    inliers = None
    if ok:
        m = int(out.inlier_mask[0]) & 0xFFFFFFFF
        inliers = np.array([i for i in range(n) if (m >> i) & 1], np.int32).reshape(-1, 1)
    return ok, rvec, tvec, inliers
