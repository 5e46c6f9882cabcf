#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/scale_n1.log 2>&1
echo "n1 rc=$?"
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/scale_n2.log 2>&1
echo "n2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/scale_n1.log', 'gpurun_out/scale_n2.log'):
    for line in open(f):
        if line.startswith('{'):
            d = json.loads(line)
            print(f, 'value', round(d['value']), 'ms_per_step', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']),
                  'parallel', round(d['throughput_mode']['value']), 'rainbow', round(d['rainbow']['env_steps_per_sec']), round(d['rainbow']['e2e_env_steps_per_sec']), d['rainbow']['ms_per_update_incl_acting'])
PY
tail -3 gpurun_out/scale_n2.log | cut -c1-200
