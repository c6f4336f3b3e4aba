"""Dev probe: how long does one decode launch take, measured five ways (profiles/decode_alone_r2.md)?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
import bench  # noqa: E402
from spe_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
cfg = bench.CONFIGS["B"]
job = bench.Job(cfg, 4 * 4096, 0, dev)
L = _lib.lib()
B, J = 4096, 11
kpts = torch.empty((B, J, 3), device=dev)
s = torch.cuda.current_stream(dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


def decode(lo):
    _lib.check(L.spe_decode_kpts_f32(job.hm[lo:lo + B].data_ptr(), B, J, 64, 64, job.c[lo:lo + B].data_ptr(), job.s[lo:lo + B].data_ptr(), 1, kpts.data_ptr(), None,
                                     s.cuda_stream), "decode")


# the same GPU's plain copy and read rates, for comparison (MEASURED_PEAKS.json is the pool's figure)
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
dst = torch.empty_like(src)
for _ in range(3):
    dst.copy_(src)
e0, e1 = ev(), ev()
torch.cuda.synchronize()
e0.record(s)
for _ in range(10):
    dst.copy_(src)
e1.record(s)
torch.cuda.synchronize()
print(f"torch copy of 1 GiB: {2 * src.numel() * 10 / e0.elapsed_time(e1) / 1e6:.0f} GB/s (read + write)")
x = src.view(torch.float32)
for _ in range(3):
    x.sum()
e0, e1 = ev(), ev()
torch.cuda.synchronize()
e0.record(s)
for _ in range(10):
    x.sum()
e1.record(s)
torch.cuda.synchronize()
print(f"torch sum of 1 GiB: {src.numel() * 10 / e0.elapsed_time(e1) / 1e6:.0f} GB/s (read only)")
print(torch.cuda.get_device_name(0), torch.cuda.get_device_properties(0).total_memory >> 20, "MiB")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.mem,clocks.max.mem,clocks.sm,ecc.mode.current,power.draw,temperature.gpu,pci.bus_id,serial", "--format=csv,noheader"], capture_output=True, text=True).stdout)

for lo in (0, 4096, 8192, 12288):
    for _ in range(3):
        decode(lo)
    torch.cuda.synchronize()
    # (a) one launch after an idle GPU, events around it
    a = []
    for _ in range(10):
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record(s); decode(lo); e1.record(s)
        torch.cuda.synchronize()
        a.append(e0.elapsed_time(e1))
    # (b) a dummy kernel first so the GPU is busy when the decode is enqueued
    b = []
    filler = torch.empty(64 << 20, device=dev)
    for _ in range(10):
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        filler.zero_(); e0.record(s); decode(lo); e1.record(s)
        torch.cuda.synchronize()
        b.append(e0.elapsed_time(e1))
    # (c) 20 back-to-back launches, divided by 20
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(20):
        decode(lo)
    e1.record(s)
    torch.cuda.synchronize()
    c = e0.elapsed_time(e1) / 20
    gb = B * bench.decode_bytes_per_frame(cfg) / 1e9
    print(f"chunk at frame {lo}: after idle {np.median(a):.4f} ms (min {min(a):.4f}), behind a running kernel {np.median(b):.4f} ms (min {min(b):.4f}), "
          f"20 back to back {c:.4f} ms = {gb / c * 1e3:.0f} GB/s")
