mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_fused_step_gpu.py tests/test_replay_buffers_gpu.py tests/test_headline_shapes_gpu.py -q -x > gpurun_out/e2e_tests.log 2>&1
echo "tests rc=$? $(tail -1 gpurun_out/e2e_tests.log)"
timeout -s KILL 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 20 --warmup 5 > gpurun_out/e2e_bench.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/e2e_bench.log'):
    if line.startswith('{'):
        d = json.loads(line); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_pass'])
PY
