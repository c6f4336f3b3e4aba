"""Heatmaps -> 6-DoF poses without leaving the device, sharded by frames across ranks.

The reference crosses a process and a file boundary between its two halves
(`pred.mat`: lib/dataset/PEdataset.py:121-123 -> pose_estimation/export_predicted_poses_real.py:172-173).
Here decode and pose solve run back to back on one stream through spe_heatmap_to_pose_f32
(include/spe_b200.h); the keypoints ([B,J,3], the pred.mat row layout) stay in HBM and are returned
so the rest of the reference pipeline can still write its files from them.

Multi-GPU: frames are independent, so rank r owns the contiguous slice
[r*N/G, (r+1)*N/G) and there is no collective in the hot loop; one all_gather of the [N/G,7]
pose tensor (28 B/frame) at the end.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib
from .models import CameraModel
from .pnp import ADAPTIVE_CONFIDENCE_FILTER, PnPSolver


@dataclass
class StageOutput:
    pose7: object  # [B,7] float32 (qw,qx,qy,qz,tx,ty,tz)
    inlier_mask: object  # [B] int32 (uint32 bits over the J landmarks)
    status: object  # [B] int32 (pnp.FRAME_*)
    kpts: object  # [B,J,3] float32 (x, y, maxval): what the reference stores in pred.mat


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced partition of n frames: the first n % world ranks get one extra."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total: int, group=None):
    """all_gather of row-sharded tensors produced with shard_bounds (uneven shards are padded to
    the largest one for the collective and trimmed afterwards).  Works on NCCL (CUDA tensors) and
    gloo (CPU tensors).  Returns the [n_total, ...] tensor on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    per = (n_total + world - 1) // world
    if n_total == per * world:  # even shards: one collective, no padding, no trimming
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


class HeatmapToPose:
    """The whole stage for one landmark/camera model on one GPU.

    hypotheses  minimal sets per frame scored by the FP32 hypothesis kernel (BASELINE.json: 256)
    iterations  cv2's iterationsCount (10000 in the reference, export_predicted_poses_real.py:201)
    exact       True (default): the winner comes from the float64 replay of cv2's own sequential loop
                (SPE_FLAG_EXACT, the parity path; hypotheses may then be 0); False: cv2's acceptance rule over the
                FP32 counts of the first `hypotheses` draws
    """

    def __init__(self, model: CameraModel, hypotheses: int = 256, reproj_err: float = 15.0, confidence: float = 0.99,
                 conf_floor: float = ADAPTIVE_CONFIDENCE_FILTER, post_process: bool = True, device=None, refine: str | None = None,
                 adaptive: bool = False, exact: bool = True, iterations: int = 10000):
        torch = _lib.require_cuda()
        self.model = model
        self.hypotheses = int(hypotheses)
        self.iterations = max(int(iterations), self.hypotheses)
        self.exact = bool(exact)
        self.reproj_err = float(reproj_err)
        self.confidence = float(confidence)
        self.conf_floor = float(conf_floor)
        self.post_process = bool(post_process)
        if refine not in (None, "lm"):
            raise ValueError("refine must be None or 'lm'")
        if self.hypotheses < (0 if self.exact else 1):
            raise ValueError("hypotheses must be positive unless exact=True")
        # adaptive=True scores only the hypotheses cv2's shrinking budget could reach (identical results)
        self.flags = ((_lib.FLAG_REFINE_LM if refine == "lm" else 0) | (_lib.FLAG_ADAPTIVE if adaptive else 0) |
                      (_lib.FLAG_EXACT if self.exact else 0))
        self.solver = PnPSolver(model.landmarks, model.K, model.dist, max_hypotheses=self.iterations, device=device)
        self.device = self.solver.device
        self._L = _lib.lib()
        self._ws = None
        self._torch = torch

    def _workspace(self, B: int):
        need = int(self._L.spe_pipeline_workspace_bytes(self.solver.handle, B, self.solver.J, self.hypotheses))
        if need == 0:
            raise _lib.SpeError("spe_pipeline_workspace_bytes rejected the configuration")
        if self._ws is None or self._ws.numel() < need:
            self._ws = self._torch.empty(need, dtype=self._torch.uint8, device=self.device)
        return self._ws

    def run_device(self, hm, center, scale, out: StageOutput | None = None) -> StageOutput:
        """hm [B,J,H,W], center/scale [B,2]: float32 CUDA contiguous.  Enqueues decode + prep +
        hypotheses + selection/refit on torch's current stream; no host synchronisation."""
        torch = self._torch
        B, J, H, W = hm.shape
        assert hm.is_cuda and hm.dtype == torch.float32 and hm.is_contiguous() and J == self.solver.J
        dev = hm.device
        if out is None:
            out = StageOutput(torch.empty((B, 7), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
                              torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B, J, 3), dtype=torch.float32, device=dev))
        ws = self._workspace(B)
        with torch.cuda.device(dev):
            _lib.check(self._L.spe_heatmap_to_pose_f32(
                self.solver.handle, hm.data_ptr(), B, J, H, W, center.data_ptr(), scale.data_ptr(), int(self.post_process),
                self.hypotheses, self.reproj_err, self.confidence, self.conf_floor, out.pose7.data_ptr(), out.inlier_mask.data_ptr(),
                out.status.data_ptr(), out.kpts.data_ptr(), ws.data_ptr(), ws.numel(), self.flags, torch.cuda.current_stream(dev).cuda_stream),
                "spe_heatmap_to_pose_f32")
        return out

    def run_host(self, heatmaps, center, scale, chunk: int = 1024) -> StageOutput:
        """Host buffers in, host buffers out: the call a user of the reference would make with
        the arrays it already has (`output.cpu().numpy()`, meta['center'], meta['scale']).

        Frames stream through the GPU in chunks on two streams so the host->device copy of chunk
        i+1 overlaps the kernels of chunk i; pinned inputs (torch CPU tensors with pin_memory, or
        NumPy views of them) are copied asynchronously, pageable ones go through the driver's
        staging path.  Returns NumPy arrays.
        """
        torch = self._torch
        hm = torch.from_numpy(heatmaps) if isinstance(heatmaps, np.ndarray) else heatmaps
        c = torch.from_numpy(np.ascontiguousarray(center, np.float32)) if not isinstance(center, torch.Tensor) else center
        s = torch.from_numpy(np.ascontiguousarray(scale, np.float32)) if not isinstance(scale, torch.Tensor) else scale
        assert hm.ndim == 4, "batch_images should be 4-ndim"
        hm = hm.contiguous().float()
        N, J, H, W = hm.shape
        dev = self.device
        chunk = max(1, min(int(chunk), N))
        copy_stream, run_stream = self._streams()
        bufs = self._staging(chunk, J, H, W)
        pose7 = torch.empty((N, 7), dtype=torch.float32).pin_memory()
        mask = torch.empty((N,), dtype=torch.int32).pin_memory()
        status = torch.empty((N,), dtype=torch.int32).pin_memory()
        kpts = torch.empty((N, J, 3), dtype=torch.float32).pin_memory()
        for i, lo in enumerate(range(0, N, chunk)):
            hi = min(N, lo + chunk)
            b = hi - lo
            buf = bufs[i % len(bufs)]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(buf["free"])  # the kernels that last read this buffer are done
                buf["hm"][:b].copy_(hm[lo:hi], non_blocking=True)
                buf["c"][:b].copy_(c[lo:hi], non_blocking=True)
                buf["s"][:b].copy_(s[lo:hi], non_blocking=True)
                buf["ready"].record(copy_stream)
            with torch.cuda.stream(run_stream):
                run_stream.wait_event(buf["ready"])
                out = StageOutput(buf["pose7"][:b], buf["mask"][:b], buf["status"][:b], buf["kpts"][:b])
                self.run_device(buf["hm"][:b], buf["c"][:b], buf["s"][:b], out)
                pose7[lo:hi].copy_(out.pose7, non_blocking=True)
                mask[lo:hi].copy_(out.inlier_mask, non_blocking=True)
                status[lo:hi].copy_(out.status, non_blocking=True)
                kpts[lo:hi].copy_(out.kpts, non_blocking=True)
                buf["free"].record(run_stream)
        run_stream.synchronize()
        return StageOutput(pose7.numpy(), mask.numpy(), status.numpy(), kpts.numpy())

    def __call__(self, heatmaps, center, scale):
        torch = self._torch
        if isinstance(heatmaps, torch.Tensor) and heatmaps.is_cuda:
            return self.run_device(heatmaps.contiguous().float(), center.to(heatmaps.device).contiguous().float(),
                                   scale.to(heatmaps.device).contiguous().float())
        return self.run_host(heatmaps, center, scale)

    # -- internals
    def _streams(self):
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = self._torch.cuda.Stream(self.device)
            self._run_stream = self._torch.cuda.Stream(self.device)
        return self._copy_stream, self._run_stream

    def _staging(self, chunk, J, H, W, depth: int = 3):
        torch = self._torch
        key = (chunk, J, H, W)
        if getattr(self, "_staging_key", None) != key:
            self._bufs = []
            for _ in range(depth):
                ev_free, ev_ready = torch.cuda.Event(), torch.cuda.Event()
                ev_free.record(torch.cuda.current_stream(self.device))
                self._bufs.append({
                    "hm": torch.empty((chunk, J, H, W), dtype=torch.float32, device=self.device),
                    "c": torch.empty((chunk, 2), dtype=torch.float32, device=self.device),
                    "s": torch.empty((chunk, 2), dtype=torch.float32, device=self.device),
                    "pose7": torch.empty((chunk, 7), dtype=torch.float32, device=self.device),
                    "mask": torch.empty((chunk,), dtype=torch.int32, device=self.device),
                    "status": torch.empty((chunk,), dtype=torch.int32, device=self.device),
                    "kpts": torch.empty((chunk, J, 3), dtype=torch.float32, device=self.device),
                    "free": ev_free, "ready": ev_ready,
                })
            self._staging_key = key
        return self._bufs


class StreamedHeatmapToPose:
    """Software-pipelined executor for a STREAM of device-resident batches.

    The stage has a throughput-bound front (decode + FP32 hypothesis scoring) and a latency-bound tail (float64 replay
    of cv2's loop in a few dependent phases, selection, float64 refit: ~0.6 ms whatever the batch size, at low SM
    occupancy).  Every slot has its own side stream: batch i's tail runs under the fronts of batches i+1 ... i+depth-1
    on the caller's stream AND next to the tails of those batches.  spe_ransac_score_f32 / spe_ransac_replay_f64 /
    spe_ransac_select_refit_f32 are the stages of the C ABI call.
    Results of a submit() are valid after wait(slot) or drain().  `depth` batches are in flight, each with its own
    workspace.  Outputs go to the slot's own tensors or to caller-provided ones (submit(..., out=StageOutput(...)), e.g.
    views of one [K*B,7] buffer that is all_gathered once at the end of the job).
    """

    def __init__(self, stage: HeatmapToPose, batch: int, depth: int = 4, want_rt: bool = False, tail_priority: int = 0, graph_tail: bool = True):
        torch = stage._torch
        self.stage, self.B, self.depth = stage, int(batch), int(depth)
        dev, J, H = stage.device, stage.solver.J, stage.hypotheses
        self._L = stage._L
        self.ws_bytes = int(self._L.spe_ransac_workspace_bytes(stage.solver.handle, self.B, H))
        self.slots = []
        cur = torch.cuda.current_stream(dev)
        for _ in range(self.depth):
            done = torch.cuda.Event()
            done.record(cur)
            self.slots.append({
                "own": StageOutput(torch.empty((self.B, 7), dtype=torch.float32, device=dev), torch.empty((self.B,), dtype=torch.int32, device=dev),
                                   torch.empty((self.B,), dtype=torch.int32, device=dev), torch.empty((self.B, J, 3), dtype=torch.float32, device=dev)),
                "out": None,
                "rt": torch.empty((self.B, 12), dtype=torch.float64, device=dev) if want_rt else None,
                "ws": torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev),
                "scored": torch.cuda.Event(), "done": done,
                "side": torch.cuda.Stream(dev, priority=tail_priority),  # (priority makes no measurable difference)
            })
        self._next = 0
        self._pending = None
        # graph_tail: a slot's tail is a chain of ~14 small dependent launches (replay phases, select/refit) with the same
        # arguments every time (the slot's workspace and its own output tensors).  After one plain run it is captured into a
        # CUDA graph and replayed with ONE launch per chunk: no host work between the phases, and eight ranks sharing a
        # host no longer stretch the chain (measured at 8 GPUs per host, profiles/step_r2.md).  Caller-provided outputs are filled
        # by three small copies behind the graph.
        self.graph_tail = bool(graph_tail)

    def submit(self, hm, center, scale, out: StageOutput | None = None, decode_events=None):
        """Enqueue one batch on torch's CURRENT stream (the stream that produced hm / center / scale).  Returns the slot
        dict: slot['out'] (StageOutput; `out` if given, with out.kpts optional), slot['done'] (event).  The inputs are only
        read by kernels of the current stream, so the caller's usual stream-ordered reuse of them stays safe."""
        torch = self.stage._torch
        st = self.stage
        B, J, H, W = hm.shape
        assert B == self.B and J == st.solver.J and hm.is_contiguous() and hm.dtype == torch.float32
        main = torch.cuda.current_stream(st.device)
        slot = self.slots[self._next]
        self._next = (self._next + 1) % self.depth
        if slot is self._pending:  # depth 1: this slot's previous tail has to run first
            self.flush()
        own = slot["own"]
        if out is None:
            out = own
        elif out.kpts is None:
            out = StageOutput(out.pose7, out.inlier_mask, out.status, own.kpts)
        slot["out"] = out
        ws = slot["ws"]
        main.wait_event(slot["done"])  # the tail that last used this slot's workspace has finished
        if decode_events is not None:
            decode_events[0].record(main)
        _lib.check(self._L.spe_decode_kpts_f32(hm.data_ptr(), B, J, H, W, center.data_ptr(), scale.data_ptr(), int(st.post_process),
                                               out.kpts.data_ptr(), None, main.cuda_stream), "spe_decode_kpts_f32")
        if decode_events is not None:
            decode_events[1].record(main)
        _lib.check(self._L.spe_ransac_score_f32(st.solver.handle, out.kpts.data_ptr(), B, st.hypotheses, st.reproj_err, st.confidence,
                                                st.conf_floor, ws.data_ptr(), ws.numel(), st.flags, main.cuda_stream), "spe_ransac_score_f32")
        slot["scored"].record(main)
        self._pending = slot
        self.flush()
        return slot

    def _tail_launches(self, slot, out, side):
        st, ws = self.stage, slot["ws"]
        if st.exact:
            _lib.check(self._L.spe_ransac_replay_f64(st.solver.handle, self.B, st.hypotheses, st.reproj_err, st.confidence, ws.data_ptr(),
                                                     ws.numel(), side.cuda_stream), "spe_ransac_replay_f64")
        _lib.check(self._L.spe_ransac_select_refit_f32(st.solver.handle, self.B, st.hypotheses, st.confidence, out.pose7.data_ptr(),
                                                       out.inlier_mask.data_ptr(), out.status.data_ptr(), None,
                                                       slot["rt"].data_ptr() if slot["rt"] is not None else None, ws.data_ptr(), ws.numel(),
                                                       st.flags | _lib.FLAG_BACKGROUND_TAIL, side.cuda_stream), "spe_ransac_select_refit_f32")

    def _enqueue_tail(self, slot):
        torch = self.stage._torch
        side, out, own = slot["side"], slot["out"], slot["own"]
        if self.graph_tail and slot.get("graph") is None and slot.get("plain_runs", 0) >= 1:
            # capture once per slot, after a plain run has done every lazy one-time set-up (kernel attributes ...)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                self._tail_launches(slot, own, side)
            slot["graph"] = g
        side.wait_event(slot["scored"])
        if slot.get("graph") is not None:
            with torch.cuda.stream(side):
                slot["graph"].replay()
                if out.pose7.data_ptr() != own.pose7.data_ptr():
                    out.pose7.copy_(own.pose7, non_blocking=True)
                    out.inlier_mask.copy_(own.inlier_mask, non_blocking=True)
                    out.status.copy_(own.status, non_blocking=True)
        else:
            self._tail_launches(slot, out, side)
            slot["plain_runs"] = slot.get("plain_runs", 0) + 1
        slot["done"].record(side)
        # outputs (and a caller-provided `out`) were allocated on other streams and are written here: tell the allocator
        for t in (out.pose7, out.inlier_mask, out.status, out.kpts, slot["ws"], slot["rt"]):
            if t is not None:
                t.record_stream(side)
        self._pending = None

    def warm_up(self, hm, center, scale):
        """Run 2 x depth batches so that every slot has done its plain run and captured its tail graph (a capture
        synchronises the device: keep it out of timed or latency-sensitive regions)."""
        for _ in range(2 * self.depth):
            self.submit(hm, center, scale)
        self.drain()
        self.stage._torch.cuda.current_stream(self.stage.device).synchronize()

    def flush(self):
        """Enqueue the tail of the most recent batch now."""
        if self._pending is not None:
            self._enqueue_tail(self._pending)

    def wait(self, slot):
        """Block the host until the results of `slot` are complete."""
        self.flush()
        slot["done"].synchronize()

    def drain(self):
        """Make torch's current stream wait for every tail in flight."""
        self.flush()
        main = self.stage._torch.cuda.current_stream(self.stage.device)
        for slot in self.slots:
            main.wait_event(slot["done"])
