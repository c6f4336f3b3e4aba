"""Dev tool: pipelined step time for the tail placement options, several repetitions in one process."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
from spe_b200 import models, synth  # noqa: E402
from spe_b200.pipeline import HeatmapToPose, StreamedHeatmapToPose  # noqa: E402

B, J, H, W, HYP = 4096, 11, 64, 64, 256
dev = torch.device("cuda", 0)
model = models.tango()
fr = synth.make_frames(model, B, H, W, seed=synth.BASE_SEED + 1)
hm = torch.from_numpy(fr.heatmaps).to(dev)
c, s = torch.from_numpy(fr.center).to(dev), torch.from_numpy(fr.scale).to(dev)
stage = HeatmapToPose(model, hypotheses=HYP, device=dev)
K = 100
res = {}
for rep in range(3):
    for tad in (False, "ovl2", "ovl3", "ovl4"):
        if tad in (False, True):
            pipe = StreamedHeatmapToPose(stage, B, depth=2, tail_after_decode=tad)
        else:
            pipe = StreamedHeatmapToPose(stage, B, depth=int(tad[3:]), overlap_decode=True)
        for _ in range(10):
            pipe.submit(hm, c, s)
        pipe.drain()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dec = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
        e0.record()
        for k in range(K):
            pipe.submit(hm, c, s, decode_events=dec[k])
        pipe.drain()
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(tad, []).append((e0.elapsed_time(e1) / K, sum(a.elapsed_time(b) for a, b in dec) / K))
for tad, v in res.items():
    print(f"decode_variant={os.environ.get('SPE_DECODE_VARIANT', 'dyn')} mode={tad}: step ms " + " ".join(f"{x[0]:.4f}" for x in v) +
          " | decode-in-step ms " + " ".join(f"{x[1]:.4f}" for x in v))
