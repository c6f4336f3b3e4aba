"""A/B builds of the library with other register budgets for the FP32 hypothesis kernel (profiles/hyp_regs_r2.md):
spacecraft-pose-estimation_b200/build/ab/libspe_r{128,208}.so (git-ignored, travels to the GPU box) next to the shipped 168-register build.  Run here (CPU), profile on the GPU box with
    ncu ... python tools/ncu_target.py --config B --lib spacecraft-pose-estimation_b200/build/ab/libspe_r128.so
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("spe_build", os.path.join(ROOT, "spacecraft-pose-estimation_b200", "build.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
AB = os.path.join(ROOT, "spacecraft-pose-estimation_b200", "build", "ab")
os.makedirs(AB, exist_ok=True)
for arg in sys.argv[1:] or ("128", "208"):
    if arg.startswith("replay"):  # replay3 / replay4: CTAs per SM of the float64 replay kernels (168 / 128 registers)
        n = int(arg[6:])
        print(mod.build(out=os.path.join(AB, f"libspe_replay{n}.so"), extra_flags=[f"-DSPE_REPLAY_CTAS_PER_SM={n}"]))
    elif arg.startswith("refit"):  # refit168 / refit128: register budget of select_refit_kernel (12 / 16 warps per background CTA)
        print(mod.build(out=os.path.join(AB, f"libspe_{arg}.so"), extra_flags=[f"-DSPE_REFIT_REGS={int(arg[5:])}"]))
    elif arg == "alldraws":  # the float64 replay evaluates every draw (no distinct-set phases)
        print(mod.build(out=os.path.join(AB, "libspe_alldraws.so"), extra_flags=["-DSPE_REPLAY_ALL_DRAWS"]))
    elif arg.startswith("shape"):  # shape<RW>x<RC>t<TW>: replay kernels of RW warps x RC CTAs per SM, background refit CTAs of TW warps
        rw, rest = arg[5:].split("x")
        rc, tw = rest.split("t")
        print(mod.build(out=os.path.join(AB, f"libspe_{arg}.so"),
                        extra_flags=[f"-DSPE_REPLAY_WARPS={rw}", f"-DSPE_REPLAY_CTAS_PER_SM={rc}", f"-DSPE_TAIL_WARPS={tw}"]))
    else:
        regs = int(arg)
        print(mod.build(out=os.path.join(AB, f"libspe_r{regs}.so"), extra_flags=[f"-DSPE_T1_REGS={regs}"]))
