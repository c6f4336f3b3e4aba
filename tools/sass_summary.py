"""profiles/sass_summary.txt: per kernel of libspe_b200.so, the SASS mnemonics that characterise it (bulk copies and
mbarrier waits of the decode, REDUX of its warp argmax, FFMA/DFMA of the solvers; no tensor-core instruction anywhere:
nothing on this path is a dense contraction).  CPU only (cuobjdump).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spacecraft-pose-estimation_b200", "spe_b200", "libspe_b200.so")
WATCH = ["UBLKCP", "SYNCS", "REDUX", "CREDUX", "LDS", "LDG", "STG", "FFMA", "FMUL", "FADD", "MUFU", "DFMA", "DMUL", "DADD", "LDL", "STL", "ATOMG", "REDG", "SHFL", "VOTE",
         "HMMA", "IMMA", "UTCHMMA", "UTCQMMA", "UTMALDG"]


def main():
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = dict(re.findall(r"Function (\S+):\n\s+REG:(\d+)", res))
    print(f"# SASS summary of libspe_b200.so (sm_100a), commit {commit}: static instruction counts per kernel")
    print("# columns: kernel | registers | total | " + " ".join(WATCH))
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    print(f"# cubin architectures: {sorted(archs)}")
    for chunk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = chunk.split("\n")[0].strip()
        ops = collections.Counter(m.split(".")[0] for m in re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", chunk, flags=re.M))
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = demangled.replace("spe::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"^void ", "", short)
        short = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", short)
        print(f"{short} | {regs.get(name, '?')} | {sum(ops.values())} | " + " ".join(f"{k}={ops.get(k, 0)}" for k in WATCH if ops.get(k, 0)))


if __name__ == "__main__":
    sys.exit(main())
