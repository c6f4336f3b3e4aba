"""Dev tool: decode of few, large maps (the reference's 384x384 / 768x768 heatmaps at small batch)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spacecraft-pose-estimation_b200"))
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
    med = ts[len(ts) // 2]
    print(f"variant {os.environ.get('SPE_DECODE_VARIANT', '0')}: B={B} J={J} {H}x{W}: {med*1e3:.1f} us -> {B*J*H*W*4/med/1e6:.0f} GB/s")


if __name__ == "__main__":
    for shape in ((8, 11, 768, 768), (4, 11, 768, 768), (1, 11, 768, 768), (32, 11, 384, 384), (8, 11, 384, 384), (64, 11, 384, 384)):
        run(*shape)
