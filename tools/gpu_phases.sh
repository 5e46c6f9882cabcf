#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_conv_tc_gpu.py -q > gpurun_out/gemm_test.log 2>&1
echo "gemm+conv tests rc=$?"; tail -2 gpurun_out/gemm_test.log
timeout 300 python tools/gemm_phases.py > gpurun_out/gemm_phases.log 2>&1; cat gpurun_out/gemm_phases.log | tail -12
timeout 600 python tools/bench_gemm.py > gpurun_out/gemm_bench.log 2> gpurun_out/gemm_bench.err
cat gpurun_out/gemm_bench.log | cut -c1-175
timeout 600 python tools/bench_conv.py > gpurun_out/conv_bench.log 2> gpurun_out/conv_bench.err; cat gpurun_out/conv_bench.log; tail -3 gpurun_out/conv_bench.err
