#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > gpurun_out/final_tests.log 2>&1
echo "suite rc=$? $(tail -1 gpurun_out/final_tests.log)"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B2RL_WB_FINE=1 timeout -s KILL 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 20 --warmup 5 > gpurun_out/wb_bench.log 2>&1
python - <<'PY'
import json
for line in open('gpurun_out/wb_bench.log'):
    if line.startswith('{'):
        d = json.loads(line)
        print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'parallel', round(d['throughput_mode']['value']), d['throughput_mode']['roofline']['frac'])
        print('exact', d['roofline']['phases_of_one_launch'])
        print('parallel', d['throughput_mode']['roofline']['phases_of_one_launch'])
PY
