#!/bin/bash
# Driver-style bench lines on N GPUs of one box: bash tools/scale_runs.sh <N> <configs...>   (e.g. 8 B D E)
# Run ON THE GPU BOX: gpurun --gpus 8 --timeout 900 -- 'bash tools/scale_runs.sh 8 B D E'.  Lines land in gpurun_out/scale/.
set -u
N=$1
shift
mkdir -p gpurun_out/scale
PORT=29511
for c in "$@"; do
  PORT=$((PORT + 1))
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --config $c --steps 20 --warmup 3 > gpurun_out/scale/bench_${c}_n1.json 2> gpurun_out/scale/bench_${c}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --config $c --steps 20 --warmup 3 \
      > gpurun_out/scale/bench_${c}_n$N.json 2> gpurun_out/scale/bench_${c}_n$N.err
  fi
  echo "n$N $c rc=$?"
done
python - "$N" "$@" <<'PY'
import json, sys
n = sys.argv[1]
for c in sys.argv[2:]:
    try:
        d = json.loads(open(f"gpurun_out/scale/bench_{c}_n{n}.json").read().strip().splitlines()[-1])
        print(f"bench_{c}_n{n}: value {d['value']:.0f} ms/step {d['ms_per_step']:.4f} spread {d.get('ms_per_step_spread')} e2e {d.get('e2e', {}).get('value')}")
    except Exception as e:
        print(f"bench_{c}_n{n}: {e}")
PY
