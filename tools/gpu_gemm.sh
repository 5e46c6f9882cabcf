#!/bin/bash
# first runs of the tcgen05 product: parity vs fp64, then timing vs cuBLAS
mkdir -p gpurun_out
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
echo "probe rc=$?"; head -60 gpurun_out/gemm_probe.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -s -x > gpurun_out/gemm_test.log 2>&1
echo "gemm tests rc=$?"
tail -5 gpurun_out/gemm_test.log
grep "^\[gemm\|^\[linear" gpurun_out/gemm_test.log | head -40
timeout 600 python tools/bench_gemm.py > gpurun_out/gemm_bench.log 2> gpurun_out/gemm_bench.err
echo "bench rc=$?"
cat gpurun_out/gemm_bench.log
tail -3 gpurun_out/gemm_bench.err
timeout 600 python tools/rainbow_profile.py > gpurun_out/rainbow_profile_tc.log 2>&1
echo "rainbow profile (tcgen05 linear) rc=$?"; grep " ms " gpurun_out/rainbow_profile_tc.log | head -20
B2RL_LINEAR=cublas timeout 600 python tools/rainbow_profile.py > gpurun_out/rainbow_profile_cublas.log 2>&1
echo "rainbow profile (cuBLAS linear) rc=$?"; grep " ms " gpurun_out/rainbow_profile_cublas.log | head -20
