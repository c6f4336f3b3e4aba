"""Whole-population pose parity report (profiles/parity_r2.md): every BASELINE config + close range, N frames each, every
solved frame against cv2.solvePnPRansac(iterationsCount=10000), for both selections of the library.

    python tools/parity_report.py [--frames 2048] [--out gpurun_out/parity_r2.md] [--whitebox 128]
"""
import argparse
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "spacecraft-pose-estimation_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import spe_b200  # noqa: E402
from parity_util import DATASETS, make_dataset, population_parity  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2048)
    ap.add_argument("--whitebox", type=int, default=128)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_r2.md"))
    ap.add_argument("--seed-offset", type=int, default=0, help="another population of the same datasets (tests use 0)")
    ap.add_argument("--exact-only", action="store_true", help="skip the FP32-only selection")
    args = ap.parse_args()
    commit = os.environ.get("SPE_COMMIT", "")
    if not commit:
        try:
            commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or "n/a"
        except Exception:
            commit = "n/a"
    lines = [f"# Pose parity over whole populations (commit {commit}, cv2 {__import__('cv2').__version__}, population seed offset {args.seed_offset})", "",
             f"{args.frames} frames per dataset, decoded by the oracle's get_final_preds; every frame with >= 6 visible landmarks compared with "
             "`cv2.solvePnPRansac(EPNP, iterationsCount=10000, reprojectionError=15)`.  Tolerance on frames with cv2's inlier set: 1e-3 deg / 1e-4 rel-t "
             "(float64 R|t).  A differing frame is *explained* when the independent float64 NumPy white box disagrees with cv2 there too, or cv2 changes "
             "its own answer under a 1-float32-ulp perturbation of its image points.", "",
             "| dataset | selection | frames | same inlier set as cv2 | max rot / t on those | differing (explained / unexplained) | worst pose difference | white box vs cv2 (sample) | GPU vs cv2 (same sample) | mean / max cv2 budget |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    detail = []
    for name in DATASETS:
        t0 = time.time()
        model, kpts = make_dataset(name, args.frames, seed_offset=args.seed_offset)
        H = DATASETS[name][4]
        solver = spe_b200.PnPSolver(model.landmarks, model.K, model.dist, max_hypotheses=10000)
        for exact in ((True,) if args.exact_only else (True, False)):
            out = solver.solve(kpts, hypotheses=H, exact=exact)
            rep = population_parity(name, model, kpts, out, iterations=10000, whitebox_sample=args.whitebox if exact else 0)
            worst = max([d[2] for d in rep.disagree], default=0.0)
            wb = f"{rep.whitebox_same}/{rep.whitebox_frames}" if rep.whitebox_frames else "—"
            gs = f"{rep.gpu_same_on_sample}/{rep.whitebox_frames}" if rep.whitebox_frames else "—"
            sel = "float64 replay (SPE_FLAG_EXACT)" if exact else f"FP32 scores, H = {H}"
            lines.append(f"| {name} | {sel} | {rep.frames} | {rep.same_mask} = {rep.agreement:.4f} | {rep.max_rot_same:.1e} deg / {rep.max_t_same:.1e} | "
                         f"{len(rep.disagree)} ({len(rep.disagree) - len(rep.unexplained)} / {len(rep.unexplained)}) | {worst:.3g} deg | {wb} | {gs} | "
                         f"{out.budget[out.status == 0].mean():.1f} / {out.budget.max()} |")
            print(rep.line(), f"[{time.time() - t0:.0f} s]", flush=True)
            for d in rep.disagree:
                detail.append(f"- {name}, {sel}, frame {d[0]}: {d[1]}; pose differs from cv2's by {d[2]:.3g} deg, {d[3]:.3g} rel-t")
        solver.close()
    lines += ["", "## Every differing frame", ""] + (detail or ["none"])
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write("\n".join(lines) + "\n")
    print("wrote", args.out)


if __name__ == "__main__":
    main()
