"""Dev probe: ms per pipelined chunk of one config for a list of rank seeds (bench.py gives rank r the data seed
BASE_SEED + 101 + r, so this shows on ONE GPU whether some rank's batch is a harder one), pipeline depths and library builds.

    python tools/step_probe.py --config B --ranks 0,1,2,3,4,5,6,7
    SPE_T1_WARPS=6 python tools/step_probe.py --lib spacecraft-pose-estimation_b200/spe_b200/libspe_b200_dev.so
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="B")
ap.add_argument("--ranks", default="0")
ap.add_argument("--depths", default="4")
ap.add_argument("--modes", default="value", help="comma list of value,replay,fp32")
ap.add_argument("--lib", default="")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--repeats", type=int, default=5)
ap.add_argument("--tag", default="")
args = ap.parse_args()
if args.lib:
    from spe_b200 import _lib

    _lib.LIB_PATH = os.path.abspath(args.lib)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
cfg = bench.CONFIGS[args.config]
frames = min(cfg["frames"], 4 * bench.CHUNK)
MODES = {"value": dict(exact=True), "replay": dict(exact=True, hypotheses=0), "fp32": dict(exact=False)}
for mode in args.modes.split(","):
    for depth in (int(d) for d in args.depths.split(",")):
        bench.PIPE_DEPTH = depth
        for rank in (int(r) for r in args.ranks.split(",")):
            job = bench.Job(cfg, frames, rank, dev, **MODES[mode])
            ms, _, out, _ = bench.timed_regions(job, args.steps, args.repeats, 1, dev, gather=False)
            per = np.array(ms) / args.steps / (frames // job.chunk)
            ok = float((out.status == 0).float().mean())
            print(f"{args.tag} {mode:6s} depth {depth} rank-seed {rank}: {np.median(per):.4f} ms per {job.chunk}-frame chunk "
                  f"(min {per.min():.4f} max {per.max():.4f}), status OK {ok:.4f}", flush=True)
            del job, out
            torch.cuda.empty_cache()
