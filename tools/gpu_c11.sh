#!/bin/bash
mkdir -p gpurun_out
for s in 1 4 16 64; do
  echo "== sleep $s" >> gpurun_out/c11.log
  B2RL_V6_SLEEP=$s timeout -s KILL 300 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['sampler'], d['roofline']['phases_of_one_launch'])" >> gpurun_out/c11.log
done
cat gpurun_out/c11.log
