"""Dev tool: where does the pipelined step time go?  Times, per 4096-frame step (config B):
  front        decode + prep + hypothesis scoring back to back (no selection/refit at all)
  score        prep + hypothesis scoring only
  serial       front + tail on ONE stream
  pipelined    StreamedHeatmapToPose (tail of step k on a side stream under the front of step k+1)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))
from spe_b200 import _lib, models, synth  # noqa: E402
from spe_b200.pipeline import HeatmapToPose, StreamedHeatmapToPose  # noqa: E402

B, J, H, W, HYP = 4096, 11, 64, 64, 256
dev = torch.device("cuda", 0)
model = models.tango()
fr = synth.make_frames(model, B, H, W, seed=synth.BASE_SEED + 1)
hm = torch.from_numpy(fr.heatmaps).to(dev)
c, s = torch.from_numpy(fr.center).to(dev), torch.from_numpy(fr.scale).to(dev)
stage = HeatmapToPose(model, hypotheses=HYP, device=dev)
L = _lib.lib()
st = stage
kpts = torch.empty((B, J, 3), dtype=torch.float32, device=dev)
pose7 = torch.empty((B, 7), dtype=torch.float32, device=dev)
mask = torch.empty((B,), dtype=torch.int32, device=dev)
status = torch.empty((B,), dtype=torch.int32, device=dev)
ws_bytes = int(L.spe_ransac_workspace_bytes(st.solver.handle, B, HYP))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)
K = 50


def decode():
    _lib.check(L.spe_decode_kpts_f32(hm.data_ptr(), B, J, H, W, c.data_ptr(), s.data_ptr(), 1, kpts.data_ptr(), None, stream.cuda_stream), "decode")


def score():
    _lib.check(L.spe_ransac_score_f32(st.solver.handle, kpts.data_ptr(), B, HYP, 15.0, 0.99, -1.0, ws.data_ptr(), ws_bytes, 0, stream.cuda_stream), "score")


def tail(flags=0):
    _lib.check(L.spe_ransac_select_refit_f32(st.solver.handle, B, HYP, 0.99, pose7.data_ptr(), mask.data_ptr(), status.data_ptr(), None, None,
                                             ws.data_ptr(), ws_bytes, flags, stream.cuda_stream), "tail")


def timeit(fn, name):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:28s} {e0.elapsed_time(e1) / K:.4f} ms/step", flush=True)


decode()
timeit(decode, "decode")
timeit(score, "score (prep + hypotheses)")
timeit(lambda: (decode(), score()), "front (decode + score)")
timeit(lambda: tail(0), "tail, shared-memory matrix")
timeit(lambda: tail(_lib.FLAG_BACKGROUND_TAIL), "tail, background variant")
timeit(lambda: (decode(), score(), tail(0)), "serial, one stream")
for tad in (False, True):
    for prio in (0, -1):
        pipe = StreamedHeatmapToPose(stage, B, depth=2, tail_after_decode=tad, tail_priority=prio)

        def piped():
            pipe.submit(hm, c, s)

        timeit(piped, f"pipelined tad={int(tad)} prio={prio}")
        pipe.drain()
pipe = StreamedHeatmapToPose(stage, B, depth=2)
torch.cuda.synchronize()
# CUDA graph of 10 pipelined steps
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(dev)
with torch.cuda.stream(side):
    pipe2 = StreamedHeatmapToPose(stage, B, depth=2)
    for _ in range(3):
        pipe2.submit(hm, c, s)
    pipe2.drain()
    torch.cuda.synchronize()
    try:
        with torch.cuda.graph(g, stream=side):
            pipe3 = StreamedHeatmapToPose(stage, B, depth=2)
            for _ in range(10):
                pipe3.submit(hm, c, s)
            pipe3.drain()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g.replay()
        torch.cuda.synchronize()
        e0.record(side)
        for _ in range(5):
            g.replay()
        e1.record(side)
        torch.cuda.synchronize()
        print(f"{'pipelined, CUDA graph x10':28s} {e0.elapsed_time(e1) / 50:.4f} ms/step")
    except Exception as ex:  # noqa: BLE001
        print("graph capture failed:", repr(ex)[:300])

# ---- timeline of the pipelined schedule (events on both streams, steady state) ----------------
torch.cuda.synchronize()
main = torch.cuda.current_stream(dev)
side2 = torch.cuda.Stream(dev)
slots = []
for _ in range(2):
    slots.append({"kpts": torch.empty_like(kpts), "ws": torch.empty_like(ws), "pose7": torch.empty_like(pose7), "mask": torch.empty_like(mask),
                  "status": torch.empty_like(status), "done": torch.cuda.Event()})
    slots[-1]["done"].record(main)
NS = 12
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(NS)]
t0 = torch.cuda.Event(enable_timing=True)
t0.record(main)
for k in range(NS):
    sl = slots[k % 2]
    main.wait_event(sl["done"])
    ev[k][0].record(main)
    _lib.check(L.spe_decode_kpts_f32(hm.data_ptr(), B, J, H, W, c.data_ptr(), s.data_ptr(), 1, sl["kpts"].data_ptr(), None, main.cuda_stream), "d")
    ev[k][1].record(main)
    _lib.check(L.spe_ransac_score_f32(st.solver.handle, sl["kpts"].data_ptr(), B, HYP, 15.0, 0.99, -1.0, sl["ws"].data_ptr(), ws_bytes, 0, main.cuda_stream), "s")
    ev[k][2].record(main)
    side2.wait_event(ev[k][2])
    ev[k][3].record(side2)
    _lib.check(L.spe_ransac_select_refit_f32(st.solver.handle, B, HYP, 0.99, sl["pose7"].data_ptr(), sl["mask"].data_ptr(), sl["status"].data_ptr(), None, None,
                                             sl["ws"].data_ptr(), ws_bytes, _lib.FLAG_BACKGROUND_TAIL, side2.cuda_stream), "t")
    ev[k][4].record(side2)
    sl["done"].record(side2)
torch.cuda.synchronize()
print("step: decode_start decode_end hyp_end | tail_start tail_end   (ms since t0)")
for k in range(NS):
    print(k, " ".join(f"{t0.elapsed_time(ev[k][i]):8.3f}" for i in range(5)))
