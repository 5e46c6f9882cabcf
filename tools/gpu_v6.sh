#!/bin/bash
# v6 exact sampler: bit-identity with the default margin, with forced slow paths and with a
# widened margin; then the quick bench and ncu captures.
set -u
mkdir -p gpurun_out
T="tests/test_per_gpu.py tests/test_headline_shapes_gpu.py tests/test_fused_step_gpu.py tests/test_replay_buffers_gpu.py"
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/v6.log
  timeout -s KILL $to "$@" > gpurun_out/v6_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/v6_${name}.log | tr '\n' ' ' | cut -c1-600)" | tee -a gpurun_out/v6.log; }
run default 600 python -m pytest $T -x -q
B2RL_V6_SLOW_EVERY=3 run slow3 600 python -m pytest $T -x -q
B2RL_V6_EPS_SCALE=1e7 run eps1e7 600 python -m pytest $T -x -q
B2RL_V6_EPS_SCALE=1e10 run eps1e10 600 python -m pytest tests/test_per_gpu.py -x -q
B2RL_SAMPLER=v5 run v5 600 python -m pytest tests/test_per_gpu.py tests/test_fused_step_gpu.py -x -q
run benchq 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary
B2RL_SAMPLER=v5 run benchq_v5 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 5 --warmup 2
# launch list + full captures (short runs)
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 600 --csv \
  --log-file gpurun_out/r02_launches_ncu.csv python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 8 > gpurun_out/v6_ncu_list.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_replay_step|k_update_paths" -s 24 -c 6 \
  -o gpurun_out/r02_step python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 8 > gpurun_out/v6_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
