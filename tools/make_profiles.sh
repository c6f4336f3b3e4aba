#!/bin/bash
# Regenerates the evidence under profiles/ at HEAD.  Run ON THE GPU BOX (one B200) from the repo root, e.g.
#   gpurun --timeout 3000 -- 'bash tools/make_profiles.sh r2'
# Everything lands in gpurun_out/profiles_<tag>/ (merged back by gpurun); copy what should be judged into profiles/.
# Each file carries the commit (or the kernel-source hash) it was produced at.
set -u
TAG=${1:-r2}
OUT=gpurun_out/profiles_$TAG
mkdir -p $OUT
# the GPU box has no .git: pass the commit in, e.g.  gpurun -- "SPE_COMMIT=$(git rev-parse --short HEAD) bash tools/make_profiles.sh r2"
COMMIT=${SPE_COMMIT:-$(git rev-parse --short HEAD 2>/dev/null || echo "n/a")}
echo "commit $COMMIT, $(nvidia-smi --query-gpu=name,driver_version --format=csv,noheader | head -1), $(date -u +%FT%TZ)" > $OUT/STAMP.txt

# 1. parity over whole populations (both selections), every differing frame listed
timeout 1500 python tools/parity_report.py --frames 2048 --out $OUT/parity_$TAG.md > $OUT/parity.log 2>&1

# 2. the bench lines: config B on one GPU, then the other configs
for c in B A C D; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 > $OUT/bench_${c}_n1_$TAG.json 2> $OUT/bench_${c}_n1.err
done
timeout 900 python bench.py --config C --landmarks 24 --steps 20 --warmup 3 > $OUT/bench_C24_n1_$TAG.json 2> $OUT/bench_C24_n1.err
timeout 900 python bench.py --config E > $OUT/bench_E_n1_$TAG.json 2> $OUT/bench_E_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref.err

# 3. launch list of three pipelined steps + of one un-pipelined chunk (per-launch times are cold-cache and serialised)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_pipelined_$TAG.csv \
  python tools/ncu_target.py --config B --pipelined 3 > /dev/null 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_single_$TAG.csv \
  python tools/ncu_target.py --config B > /dev/null 2>&1

# 4. ncu --set full of every kernel of one chunk, configs B and D (the counters bench.py quotes come from these)
#    (the .ncu-rep files are ~45 MB each and gpurun brings back 64 MiB at most: keep the raw-metric CSV, drop the report)
for c in B D; do
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/stage_$c python tools/ncu_target.py --config $c > $OUT/ncu_$c.log 2>&1
  ncu -i $OUT/stage_$c.ncu-rep --page raw --csv > $OUT/stage_${c}_raw.csv 2>/dev/null
  rm -f $OUT/stage_$c.ncu-rep
done

# 5. sanitizer pass over the smoke path
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $OUT/sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $OUT/sanitizer_racecheck.log 2>&1
for f in $OUT/sanitizer_*.log; do echo "== $f"; tail -n 4 $f; done
ls -la $OUT
