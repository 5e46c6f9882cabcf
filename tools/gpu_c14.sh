#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 \
  -k regex:"k_sample_exact_v6" -s 4 -c 1 -o gpurun_out/r02_v6f \
  python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 4 > gpurun_out/c14_ncu.log 2>&1
ls -la gpurun_out/r02_v6f.ncu-rep
