"""Dev probe (profiles/decode_alone_r2.md): the decode launch timed with CUDA events for two builds of the library
(--libs a.so b.so ...), on data produced two ways (uploaded host frames / rendered on the device), in several L2 states."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
from spe_b200 import models, synth  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, J = 4096, 11
model = models.tango()
fr = synth.make_frames(model, B, 64, 64, seed=synth.BASE_SEED + 1)
data = {
    "host frames uploaded": (torch.from_numpy(fr.heatmaps).to(dev), torch.from_numpy(fr.center).to(dev), torch.from_numpy(fr.scale).to(dev)),
    "rendered on device": synth.device_heatmaps(model, B, 64, 64, seed=7, device=dev),
    "randn": (torch.randn((B, J, 64, 64), device=dev), torch.from_numpy(fr.center).to(dev), torch.from_numpy(fr.scale).to(dev)),
}
kpts = torch.empty((B, J, 3), device=dev)
s = torch.cuda.current_stream(dev)
other = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


for path in sys.argv[1:]:
    L = ctypes.CDLL(os.path.abspath(path))
    L.spe_decode_kpts_f32.restype = ctypes.c_int
    L.spe_decode_kpts_f32.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 3

    def decode(hm, c, sc):
        assert L.spe_decode_kpts_f32(hm.data_ptr(), B, J, 64, 64, c.data_ptr(), sc.data_ptr(), 1, kpts.data_ptr(), None, s.cuda_stream) == 0

    for name, (hm, c, sc) in data.items():
        for _ in range(3):
            decode(hm, c, sc)
        torch.cuda.synchronize()
        res = {}
        for state in ("back to back", "after idle", "after reading 512 MB", "after writing 512 MB"):
            t = []
            for _ in range(10):
                e0, e1 = ev(), ev()
                if state == "after idle":
                    torch.cuda.synchronize()
                elif state == "after reading 512 MB":
                    other.view(torch.float32).sum()
                elif state == "after writing 512 MB":
                    other.zero_()
                e0.record(s); decode(hm, c, sc); e1.record(s)
                if state != "back to back":
                    torch.cuda.synchronize()
                    t.append(e0.elapsed_time(e1))
                else:
                    t.append((e0, e1))
            if state == "back to back":
                torch.cuda.synchronize()
                t = [a.elapsed_time(b) for a, b in t]
            res[state] = float(np.median(t))
        print(f"{os.path.basename(path):22s} {name:22s} " + "  ".join(f"{k}: {v:.4f} ms" for k, v in res.items()), flush=True)
