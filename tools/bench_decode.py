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


def run_combined(B, J, H, W, K, mode, iters=20):
    """f2: fused combine+decode vs the reference's sequence (torch ops, then decode) on the device."""
    hms = [torch.randn((B, J, H, W), device="cuda") for _ in range(K)]
    c = torch.rand((B, 2), device="cuda") * 1000 + 100
    s = torch.rand((B, 2), device="cuda") * 3 + 0.5

    def fused():
        spe_b200.get_final_preds_combined(True, hms, c, s, mode=mode, shift_heatmap=True)

    def unfused():
        if mode == "mean":
            out = hms[0].clone()
            for h in hms[1:]:
                out += h
            out = out / K
        else:
            of = hms[1].flip(3)
            of[:, :, :, 1:] = of.clone()[:, :, :, 0:-1]
            out = (hms[0] + of) * 0.5
        spe_b200.decode_device(out, c, s, True)

    res = {}
    for name, fn in (("fused", fused), ("unfused", unfused)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        ev[0].record()
        for i in range(iters):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
        res[name] = ts[len(ts) // 2]
    byt = B * J * H * W * 4 * K
    print(f"combine {mode} K={K} B={B} J={J} {H}x{W}: fused {res['fused']*1e3:.1f} us = {byt/res['fused']/1e6:.0f} GB/s of input; "
          f"torch ops + decode {res['unfused']*1e3:.1f} us  ({res['unfused']/res['fused']:.2f}x)")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "combined":
        run_combined(4096, 11, 64, 64, 2, "flip")
        run_combined(4096, 11, 64, 64, 2, "mean")
        run_combined(4096, 11, 64, 64, 6, "mean")
        run_combined(2048, 11, 128, 128, 2, "flip")
        sys.exit(0)
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
