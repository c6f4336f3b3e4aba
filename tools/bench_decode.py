"""Quick decode-only timing (dev tool): GB/s of spe_decode_f32 on synthetic heatmaps in HBM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
import spe_b200  # noqa: E402


def run(B, J, H, W, iters=20):
    hm = torch.randn((B, J, H, W), device="cuda")
    c = torch.rand((B, 2), device="cuda") * 1000 + 100
    s = torch.rand((B, 2), device="cuda") * 3 + 0.5
    for _ in range(3):
        spe_b200.decode_device(hm, c, s, True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        spe_b200.decode_device(hm, c, s, True)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    byt = B * (J * H * W * 4 + J * 12 + 16)
    med = ts[len(ts) // 2]
    print(f"B={B} J={J} {H}x{W}: median {med*1e3:.1f} us  best {ts[0]*1e3:.1f} us  -> {byt/med/1e6:.0f} GB/s (best {byt/ts[0]/1e6:.0f})")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        run(4096, 11, 64, 64)
        run(16384, 17, 96, 72)
        run(2048, 11, 128, 128)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run(4096, 11, 64, 64, iters=3)
        sys.exit(0)
    run(4096, 11, 64, 64)
    run(16384, 17, 96, 72)
    run(2048, 11, 128, 128)
    run(64, 11, 64, 64)
    run(256, 11, 384, 384)
