"""Profiling target: one chunk of a BASELINE config through the four stages of the C ABI (decode, FP32 scoring, float64
replay, select/refit), after a warm-up pass, bracketed by cudaProfilerStart/Stop.

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/stage_B python tools/ncu_target.py --config B
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_B.csv python tools/ncu_target.py --config B --pipelined 3
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "spacecraft-pose-estimation_b200"))

import torch  # noqa: E402

import bench  # noqa: E402
from spe_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="B")
    ap.add_argument("--landmarks", type=int, default=0)
    ap.add_argument("--pipelined", type=int, default=0, help="profile this many pipelined steps instead of one un-pipelined chunk")
    ap.add_argument("--lib", default="", help="A/B runs: load this build of the library instead of spe_b200/libspe_b200.so")
    args = ap.parse_args()
    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)
    cfg = dict(bench.CONFIGS[args.config])
    if args.config == "C" and args.landmarks:
        cfg["J"] = args.landmarks
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    frames = min(cfg["frames"], bench.CHUNK) if not args.pipelined else min(cfg["frames"], 4 * bench.CHUNK)
    job = bench.Job(cfg, frames, 0, dev)
    L = _lib.lib()
    J, (H_, W_), H = cfg["J"], cfg["hm"], cfg["H"]
    B = job.chunk
    stream = torch.cuda.current_stream(dev)
    if args.pipelined:
        out = job.outputs(args.pipelined)
        job.run_steps(2, job.outputs(2))
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        job.run_steps(args.pipelined, out)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    kpts = torch.empty((B, J, 3), device=dev)
    pose7 = torch.empty((B, 7), device=dev)
    mask = torch.empty((B,), dtype=torch.int32, device=dev)
    status = torch.empty((B,), dtype=torch.int32, device=dev)
    handle = job.stage.solver.handle
    nbytes = int(L.spe_ransac_workspace_bytes(handle, B, H))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def one_pass():
        _lib.check(L.spe_decode_kpts_f32(job.hm.data_ptr(), B, J, H_, W_, job.c.data_ptr(), job.s.data_ptr(), 1, kpts.data_ptr(), None, stream.cuda_stream), "decode")
        _lib.check(L.spe_ransac_score_f32(handle, kpts.data_ptr(), B, H, 15.0, 0.99, -1.0, ws.data_ptr(), nbytes, _lib.FLAG_EXACT, stream.cuda_stream), "score")
        _lib.check(L.spe_ransac_replay_f64(handle, B, H, 15.0, 0.99, ws.data_ptr(), nbytes, stream.cuda_stream), "replay")
        _lib.check(L.spe_ransac_select_refit_f32(handle, B, H, 0.99, pose7.data_ptr(), mask.data_ptr(), status.data_ptr(), None, None, ws.data_ptr(), nbytes,
                                                 _lib.FLAG_EXACT, stream.cuda_stream), "refit")
        torch.cuda.synchronize()

    one_pass()
    torch.cuda.profiler.start()
    one_pass()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
