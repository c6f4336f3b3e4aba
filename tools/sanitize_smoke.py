"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the
stage on small shapes, including unaligned maps, multi-chunk maps, NaNs, masked landmarks, the
adaptive second pass, LM refinement and the fused combine decode."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
import spe_b200  # noqa: E402
from spe_b200.pipeline import HeatmapToPose  # noqa: E402

rng = np.random.default_rng(0)
for shape in ((3, 11, 64, 64), (2, 5, 17, 19), (2, 3, 96, 72), (1, 2, 128, 128), (40, 1, 8, 8)):
    hm = rng.normal(size=shape).astype(np.float32)
    hm[0, 0, 1, 1] = np.nan
    c = rng.uniform(100, 900, (shape[0], 2)).astype(np.float32)
    s = rng.uniform(0.5, 3, (shape[0], 2)).astype(np.float32)
    spe_b200.get_final_preds(True, hm, c, s, return_index=True)
    spe_b200.get_max_preds(hm)
    spe_b200.accuracy(hm, np.roll(hm, 1, axis=3))
    spe_b200.get_final_preds_combined(True, [hm, hm[..., ::-1].copy()], c, s, mode="flip", shift_heatmap=True)
    spe_b200.get_final_preds_combined(True, [hm, hm, hm], c, s, mode="mean")
m = spe_b200.models.tango()
fr = spe_b200.synth.make_frames(m, 48, 64, 64, seed=3, p_outlier=0.3, p_masked=0.1)
for kw in ({}, {"adaptive": True, "exact": False}, {"refine": "lm"}, {"exact": False}, {"hypotheses": 0}):
    kw = dict({"hypotheses": 96, "iterations": 600}, **kw)  # exact replay up to 600 draws: several phases, junk frames run them all
    st = HeatmapToPose(m, **kw)
    out = st(fr.heatmaps, fr.center, fr.scale)
    dev = st(torch.from_numpy(fr.heatmaps).cuda(), torch.from_numpy(fr.center).cuda(), torch.from_numpy(fr.scale).cuda())
    torch.cuda.synchronize()
# junk frames (no model): the replay walks every phase to the end of the budget
junk = fr.heatmaps.copy()
junk[:8] = rng.normal(size=junk[:8].shape).astype(np.float32)
HeatmapToPose(m, hypotheses=64, iterations=3000)(junk, fr.center, fr.scale)
torch.cuda.synchronize()
# software-pipelined executor: every slot's float64 tail on its own side stream
from spe_b200.pipeline import StreamedHeatmapToPose  # noqa: E402

st = HeatmapToPose(m, hypotheses=96, iterations=600)
pipe = StreamedHeatmapToPose(st, 48, depth=3)
dhm, dc, ds = torch.from_numpy(fr.heatmaps).cuda(), torch.from_numpy(fr.center).cuda(), torch.from_numpy(fr.scale).cuda()
for _ in range(9):  # past the first pass every slot replays its tail as a CUDA graph
    slot = pipe.submit(dhm, dc, ds)
pipe.drain()
torch.cuda.synchronize()
# detection boxes -> (center, scale)
from spe_b200 import boxes  # noqa: E402

boxes.pick_boxes(rng.uniform(0, 500, (9, 3, 4)).astype(np.float32), rng.uniform(0, 1, (9, 3)).astype(np.float32), np.arange(9, dtype=np.int32) % 4, 1920, 1200)
boxes.xywh2cs(rng.uniform(0, 500, (7, 4)))
h = spe_b200.models.hubble_synthetic(24)
fr = spe_b200.synth.make_frames(h, 8, 96, 72, seed=4, z_range=(3.0, 8.0))
HeatmapToPose(h, hypotheses=64)(fr.heatmaps, fr.center, fr.scale)
torch.cuda.synchronize()
print("sanitize smoke done, ok frames:", int((out.status == 0).sum()))
