#!/bin/bash
mkdir -p gpurun_out
for cfg in "auto auto" "cublas cudnn" "tcgen05 tcgen05" "tcgen05 auto" "auto cudnn"; do
  set -- $cfg
  B2RL_LINEAR=$1 B2RL_CONV=$2 timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 2 --warmup 3 --passes 8 > gpurun_out/rb_bench_$1_$2.log 2>&1
  echo "== linear=$1 conv=$2 rc=$?"
  python - <<PY
import json
for line in open('gpurun_out/rb_bench_$1_$2.log'):
    if line.startswith('{'):
        d = json.loads(line)['rainbow']; print(d['env_steps_per_sec'], d['e2e_env_steps_per_sec'], d['ms_per_update_incl_acting'])
PY
done
