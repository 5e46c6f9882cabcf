#!/bin/bash
set -u
mkdir -p gpurun_out
T="tests/test_per_gpu.py tests/test_headline_shapes_gpu.py tests/test_fused_step_gpu.py tests/test_replay_buffers_gpu.py tests/test_sac_kernels_gpu.py"
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/c5.log
  timeout -s KILL $to "$@" > gpurun_out/c5_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/c5_${name}.log | tr '\n' ' ' | cut -c1-400)" | tee -a gpurun_out/c5.log; }
run tests 900 python -m pytest $T -x -q
run agents 900 python -m pytest tests/test_agents_gpu.py -x -q
run benchq 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 8 --warmup 2
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 \
  -k regex:"k_sample_exact_v6" -s 4 -c 2 -o gpurun_out/r02_v6 \
  python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 4 > gpurun_out/c5_ncu_v6.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_replay_step" -s 30 -c 4 -o gpurun_out/r02_step2 \
  python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 12 > gpurun_out/c5_ncu_step.log 2>&1
ls -la gpurun_out | tail -4
