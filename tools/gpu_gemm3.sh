#!/bin/bash
mkdir -p gpurun_out
true
true
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_conv_tc_gpu.py -q -s > gpurun_out/gemm_test.log 2>&1
echo "gemm+conv tests rc=$?"; tail -3 gpurun_out/gemm_test.log
timeout 600 python tools/bench_gemm.py > gpurun_out/gemm_bench.log 2> gpurun_out/gemm_bench.err
cat gpurun_out/gemm_bench.log | cut -c1-200
timeout 600 python tools/bench_conv.py > gpurun_out/conv_bench.log 2> gpurun_out/conv_bench.err; cat gpurun_out/conv_bench.log; tail -3 gpurun_out/conv_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32x3 -s 3 -c 1 \
  -o gpurun_out/r02_gemm python tools/bench_gemm.py > gpurun_out/gemm_ncu.log 2>&1
echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_gemm" -c 400 --csv \
  --log-file gpurun_out/conv_launches.csv python tools/bench_conv.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/conv_launches.csv')))
h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
ix = {n: i for i, n in enumerate(rows[h])}
seq = [(r[ix['Kernel Name']][:44], r[ix['Grid Size']], float(r[ix['Metric Value']])) for r in rows[h+1:] if len(r) > ix['Metric Value']]
out = []
for name, grid, t in seq:
    if out and out[-1][0] == (name, grid): out[-1][1].append(t)
    else: out.append([(name, grid), [t]])
for (name, grid), ts in out[:60]:
    ts.sort(); m = ts[len(ts)//2]; print(name, grid, len(ts), 'median us', m / 1000 if m > 1000 else m)
PY
