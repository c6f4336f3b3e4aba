"""Throughput sweep of the heatmap->pose stage on one GPU (BASELINE.json configs[2..4]).

    python tools/sweep.py [--quick] [--out gpurun_out/sweep.jsonl]

Lines: config C (Hubble-like 96x72 maps, J = 17 and 24, batch 16384, 256 hypotheses), config D
(Tango 128x128, 1024 hypotheses, batch 65536 on ONE GPU = 47 GB of heatmaps) and config E (batch
1 ... 1 M frames x 64 ... 2048 hypotheses at 11 x 64x64).  Inputs are resident in HBM; a base batch
from the seeded host generator (spe_b200.synth) is tiled on the device up to the batch size, batches
above 131072 frames run as 131072-frame chunks over the same buffer (23.6 GB, far beyond the L2).
Timing: CUDA events around `reps` back-to-back HeatmapToPose.run_device calls, after 2 warm-up calls.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
from spe_b200 import models, synth  # noqa: E402
from spe_b200.pipeline import HeatmapToPose, StageOutput  # noqa: E402

CHUNK = 131072


def tiled(fr, frames, dev):
    base = fr.heatmaps.shape[0]
    rep = (frames + base - 1) // base
    hm = torch.from_numpy(fr.heatmaps).to(dev)
    c, s = torch.from_numpy(fr.center).to(dev), torch.from_numpy(fr.scale).to(dev)
    if rep > 1:
        hm = hm.repeat(rep, 1, 1, 1)[:frames].contiguous()
        c, s = c.repeat(rep, 1)[:frames].contiguous(), s.repeat(rep, 1)[:frames].contiguous()
    else:
        hm, c, s = hm[:frames].contiguous(), c[:frames].contiguous(), s[:frames].contiguous()
    return hm, c, s


def measure(model, fr, frames, hyp, dev, label, adaptive=False):
    resident = min(frames, CHUNK)
    hm, c, s = tiled(fr, resident, dev)
    B, J, H, W = hm.shape
    stage = HeatmapToPose(model, hypotheses=hyp, device=dev, adaptive=adaptive)
    out = StageOutput(torch.empty((B, 7), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
                      torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B, J, 3), dtype=torch.float32, device=dev))
    calls = (frames + resident - 1) // resident
    for _ in range(2):
        stage.run_device(hm, c, s, out)
    torch.cuda.synchronize(dev)
    # enough repetitions for ~50 ms of work, at least 3 and at most 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stage.run_device(hm, c, s, out)
    e1.record()
    torch.cuda.synchronize(dev)
    one = max(e0.elapsed_time(e1), 1e-3)
    reps = int(min(200, max(3, 50.0 / (one * calls))))
    e0.record()
    for _ in range(reps * calls):
        stage.run_device(hm, c, s, out)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    ok = float((out.status == 0).float().mean())
    rec = {"label": label, "frames": frames, "landmarks": J, "heatmap": [H, W], "hypotheses": hyp, "adaptive": adaptive, "ms": ms,
           "frames_per_s": frames / (ms * 1e-3), "resident_frames": resident, "resident_gb": hm.numel() * 4 / 1e9, "calls": calls,
           "reps": reps, "solved_frac": ok}
    del hm, c, s, out, stage
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    recs = []

    def emit(r):
        recs.append(r)
        print(json.dumps(r), flush=True)
        with open(args.out, "a") as f:
            f.write(json.dumps(r) + "\n")

    tango = models.tango()
    base64 = synth.make_frames(tango, 2048, 64, 64, seed=synth.BASE_SEED + 4)
    # config E
    batches = [1, 64, 4096, 65536] if args.quick else [1, 16, 256, 4096, 16384, 65536, 262144, 1048576]
    hyps = [64, 2048] if args.quick else [64, 256, 1024, 2048]
    for frames in batches:
        for hyp in hyps:
            emit(measure(tango, base64, frames, hyp, dev, "E"))
    for frames in ([4096] if args.quick else [4096, 65536, 1048576]):
        emit(measure(tango, base64, frames, 256, dev, "E-adaptive", adaptive=True))
    del base64
    # config C
    for Jn in (17, 24):
        hub = models.hubble_synthetic(Jn)
        fr = synth.make_frames(hub, 512, 96, 72, seed=synth.BASE_SEED + 2, z_range=(3.0, 8.0))
        emit(measure(hub, fr, 16384, 256, dev, "C"))
    # config D on one GPU (its 2/4/8-GPU form is the same call on 32768/16384/8192 frames per rank)
    fr = synth.make_frames(tango, 256, 128, 128, seed=synth.BASE_SEED + 3)
    for frames in ([8192] if args.quick else [8192, 16384, 32768, 65536]):
        emit(measure(tango, fr, frames, 1024, dev, "D"))


if __name__ == "__main__":
    main()
