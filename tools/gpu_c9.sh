#!/bin/bash
set -u
mkdir -p gpurun_out
T="tests/test_per_gpu.py tests/test_headline_shapes_gpu.py tests/test_fused_step_gpu.py tests/test_replay_buffers_gpu.py"
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/c9.log
  timeout -s KILL $to "$@" > gpurun_out/c9_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/c9_${name}.log | tr '\n' ' ' | cut -c1-300)" | tee -a gpurun_out/c9.log; }
run default 900 python -m pytest $T -x -q
B2RL_V6_SLOW_EVERY=3 run slow3 600 python -m pytest $T -x -q
B2RL_V6_EPS_SCALE=1e7 run eps1e7 600 python -m pytest tests/test_per_gpu.py tests/test_headline_shapes_gpu.py -x -q
run new 900 python -m pytest tests/test_conv_gpu.py tests/test_train_driver_gpu.py -q
run benchq 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 8 --warmup 2
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 \
  -k regex:"k_sample_exact_v6" -s 4 -c 1 -o gpurun_out/r02_v6e \
  python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 4 > gpurun_out/c9_ncu_v6.log 2>&1
ls -la gpurun_out | tail -3
