"""Dev probe (profiles/step_r2.md): ms per pipelined 4096-frame step at config B for pipeline depths 1..6 and the three
selections (FP32 + float64 replay, replay only, FP32 only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "B"]
frames = min(cfg["frames"], 4 * bench.CHUNK)
for name, kw in (("fp32+replay", dict(exact=True)), ("replay only", dict(exact=True, hypotheses=0)), ("fp32 only", dict(exact=False))):
    row = []
    for depth in (1, 2, 3, 4, 5, 6):
        bench.PIPE_DEPTH = depth
        job = bench.Job(cfg, frames, 0, dev, **kw)
        ms, _, _, _ = bench.timed_regions(job, 20, 5, 1, dev, gather=False)
        row.append(float(np.median(ms)) / 20 / (frames // job.chunk))
        del job
    print(f"{name:12s} ms per 4096-frame chunk at depth 1..6: " + "  ".join(f"{x:.4f}" for x in row), flush=True)
