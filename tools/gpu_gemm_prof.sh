#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32x3 -s 3 -c 3 \
  -o gpurun_out/r02_gemm python tools/bench_gemm.py > gpurun_out/gemm_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/gemm_ncu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_gemm" -c 300 --csv \
  --log-file gpurun_out/gemm_launches.csv python tools/bench_gemm.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/gemm_launches.csv')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
ix = {n: i for i, n in enumerate(rows[h])}
seq = [(r[ix['Kernel Name']][:40], r[ix['Grid Size']] if 'Grid Size' in ix else '', float(r[ix['Metric Value']])) for r in rows[h+1:] if len(r) > ix['Metric Value']]
# print the sequence compressed: consecutive identical (name, grid) groups with median time
out = []
for name, grid, t in seq:
    if out and out[-1][0] == (name, grid):
        out[-1][1].append(t)
    else:
        out.append([(name, grid), [t]])
for (name, grid), ts in out[:80]:
    ts.sort(); print(name, grid, len(ts), 'median us', ts[len(ts)//2] / 1000 if ts[0] > 1000 else ts[len(ts)//2])
PY
