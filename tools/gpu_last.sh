#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q > gpurun_out/final_tests.log 2>&1
echo "suite rc=$? $(tail -1 gpurun_out/final_tests.log)"
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$? $(tail -1 gpurun_out/final_smoke.log)"
timeout -s KILL 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench.log 2>&1
echo "bench rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/final_bench.log'):
    if line.startswith('{'):
        d = json.loads(line)
        print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'traffic', d['roofline']['traffic'], 'frac', round(d['roofline']['frac'], 4),
              'parallel', round(d['throughput_mode']['value']), d['throughput_mode']['roofline']['traffic'],
              'rainbow', d['rainbow']['env_steps_per_sec'], d['rainbow']['e2e_env_steps_per_sec'], d['rainbow']['ms_per_update_incl_acting'])
        print(json.dumps(d['secondary'].get('k10_tensor_cores'))[:600])
PY
