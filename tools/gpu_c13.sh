#!/bin/bash
mkdir -p gpurun_out
T="tests/test_per_gpu.py tests/test_headline_shapes_gpu.py tests/test_fused_step_gpu.py"
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/c13.log
  timeout -s KILL $to "$@" > gpurun_out/c13_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/c13_${name}.log | tr '\n' ' ' | cut -c1-300)" | tee -a gpurun_out/c13.log; }
run default 900 python -m pytest $T -x -q
B2RL_V6_SLOW_EVERY=3 run slow3 600 python -m pytest $T -x -q
B2RL_V6_EPS_SCALE=1e7 run eps1e7 600 python -m pytest tests/test_per_gpu.py tests/test_headline_shapes_gpu.py -x -q
B2RL_V6_CYCLES=1 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 3 --warmup 1 2>gpurun_out/c13.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['sampler']); print(json.dumps(d['roofline']['phases_of_one_launch'], indent=1))" >> gpurun_out/c13.log
tail -3 gpurun_out/c13.err; cat gpurun_out/c13.log
