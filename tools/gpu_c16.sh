#!/bin/bash
mkdir -p gpurun_out
T="tests/test_per_gpu.py tests/test_headline_shapes_gpu.py tests/test_fused_step_gpu.py tests/test_replay_buffers_gpu.py"
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/c16.log
  timeout -s KILL $to "$@" > gpurun_out/c16_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/c16_${name}.log | tr '\n' ' ' | cut -c1-300)" | tee -a gpurun_out/c16.log; }
run default 900 python -m pytest $T -x -q
B2RL_V6_SLOW_EVERY=3 run slow3 600 python -m pytest $T -x -q
B2RL_V6_EPS_SCALE=1e7 run eps1e7 600 python -m pytest tests/test_per_gpu.py tests/test_headline_shapes_gpu.py -x -q
B2RL_V6_EPS_SCALE=1e10 run eps1e10 600 python -m pytest tests/test_per_gpu.py tests/test_headline_shapes_gpu.py -x -q
timeout -s KILL 300 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 5 --warmup 2 2>gpurun_out/c16.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['sampler']); print(json.dumps(d['roofline']['phases_of_one_launch'])); print(json.dumps(d['throughput_mode']['roofline']['phases_of_one_launch']), d['throughput_mode']['ms_per_pass'], d['throughput_mode']['u8_out']['ms_per_pass'])" >> gpurun_out/c16.log
tail -3 gpurun_out/c16.err; cat gpurun_out/c16.log
