#!/bin/bash
mkdir -p gpurun_out
B2RL_V6_CYCLES=1 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 3 --warmup 1 2>gpurun_out/c12.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['sampler']); print(json.dumps(d['roofline']['phases_of_one_launch'], indent=1))" > gpurun_out/c12.log
tail -5 gpurun_out/c12.err; cat gpurun_out/c12.log
