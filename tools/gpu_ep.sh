#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_conv_tc_gpu.py -q > gpurun_out/gemm_test.log 2>&1
echo "gemm+conv tests rc=$?"; tail -2 gpurun_out/gemm_test.log
timeout 300 python tools/gemm_phases.py > gpurun_out/gemm_phases.log 2>&1; cat gpurun_out/gemm_phases.log | tail -8 | cut -c1-300
timeout 600 python tools/bench_gemm.py > gpurun_out/gemm_bench.log 2> gpurun_out/gemm_bench.err
cat gpurun_out/gemm_bench.log | cut -c1-175
B2RL_LINEAR=auto B2RL_CONV=auto timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 2 --warmup 3 --passes 8 > gpurun_out/rb_bench_auto_auto.log 2>&1
python - <<PY
import json
for line in open('gpurun_out/rb_bench_auto_auto.log'):
    if line.startswith('{'):
        d = json.loads(line)['rainbow']; print(d['env_steps_per_sec'], d['e2e_env_steps_per_sec'], d['ms_per_update_incl_acting'])
PY
