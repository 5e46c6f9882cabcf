#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
echo "probe rc=$? ok-lines $(grep -c 'mismatches 0' gpurun_out/gemm_probe.log) of 12"
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -s > gpurun_out/gemm_test.log 2>&1
echo "gemm tests rc=$?"; tail -3 gpurun_out/gemm_test.log
timeout 900 python -m pytest tests/test_conv_tc_gpu.py -q -s > gpurun_out/conv_test.log 2>&1
echo "conv tests rc=$?"; grep -E "^\[conv|dgrad rel|passed|failed|Error" gpurun_out/conv_test.log | head -40
timeout 600 python tools/bench_gemm.py > gpurun_out/gemm_bench.log 2> gpurun_out/gemm_bench.err
cat gpurun_out/gemm_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tf32x3 -s 3 -c 1 \
  -o gpurun_out/r02_gemm python tools/bench_gemm.py > gpurun_out/gemm_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 python tools/bench_conv.py > gpurun_out/conv_bench.log 2> gpurun_out/conv_bench.err; cat gpurun_out/conv_bench.log; tail -3 gpurun_out/conv_bench.err
