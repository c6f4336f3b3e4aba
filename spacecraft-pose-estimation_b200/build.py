"""Build libspe_b200.so (sm_100a) in-tree with nvcc.  No JIT cache, no torch extension machinery:
the library is a plain C-ABI shared object loaded with ctypes (include/spe_b200.h).

Every .cu under csrc/ is compiled to an object file (in parallel, only when it or a header changed) and
linked.  `--dev` adds -DSPE_DEV: the development kernel variants and their environment knobs
(csrc/dev_variants.cuh), written to libspe_b200_dev.so so that the shipped library never contains them.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "spe_b200", "libspe_b200.so")
OUT_DEV = os.path.join(HERE, "spe_b200", "libspe_b200_dev.so")
SOURCES = ["decode.cu", "ransac_model.cu", "ransac_score.cu", "ransac_exact.cu", "ransac_refit.cu", "boxes.cu", "evaluate.cu", "capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
]


def _headers():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "spe_b200.h")]


def _stamp(paths, flags):
    h = hashlib.sha1(" ".join(flags).encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, dev: bool = False, out: str | None = None, extra_flags=()) -> str:
    """out / extra_flags: A/B builds (tools/ab_builds.py), e.g. another register budget into another file."""
    tag = "_dev" if dev else ""
    if out is not None:
        tag = "_" + os.path.splitext(os.path.basename(out))[0]
    out = out or (OUT_DEV if dev else OUT)
    extra = os.environ.get("SPE_NVCC_EXTRA", "").split() + (["-DSPE_DEV"] if dev else []) + list(extra_flags)
    flags = NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else [])
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = _headers()
    jobs, objs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", tag + ".o"))
        stamp_file = obj + ".stamp"
        stamp = _stamp([src] + hdrs, flags)
        objs.append(obj)
        fresh = os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp
        if force or verbose or not fresh:
            jobs.append((src, obj, stamp_file, stamp))

    def compile_one(job):
        src, obj, stamp_file, stamp = job
        cmd = ["nvcc", *flags, "-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        with open(stamp_file, "w") as f:
            f.write(stamp)
        return res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for log in pool.map(compile_one, jobs):
                if verbose:
                    sys.stderr.write(log)
    if jobs or not os.path.exists(out) or any(os.path.getmtime(o) > os.path.getmtime(out) for o in objs):
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    return out


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "spe_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, dev="--dev" in sys.argv))
