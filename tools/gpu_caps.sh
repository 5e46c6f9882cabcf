#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_conv_tc_gpu.py -q > gpurun_out/gemm_test.log 2>&1
echo "gemm+conv tests rc=$?"; tail -2 gpurun_out/gemm_test.log
B="python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 4"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_replay_step -s 2 -c 1 -f -o gpurun_out/r02_step_exact $B > gpurun_out/final_ncu_r02_step_exact.log 2>&1
echo "exact rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:k_replay_step -s 14 -c 14 -f -o gpurun_out/r02_step_tail $B > gpurun_out/final_ncu_r02_step_tail.log 2>&1
echo "tail rc=$?"; ls -la gpurun_out/r02_step_exact.ncu-rep gpurun_out/r02_step_tail.ncu-rep
timeout 900 python bench.py --no-cpu-baseline --no-secondary --steps 10 --warmup 3 --passes 16 > gpurun_out/rb_check.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/rb_check.log'):
    if line.startswith('{'):
        d = json.loads(line)['rainbow']; print(d['env_steps_per_sec'], d['e2e_env_steps_per_sec'], d['ms_per_update_incl_acting'], d['vector_steps'])
PY
