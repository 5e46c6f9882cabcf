#!/bin/bash
# First GPU call of the next round (one B200, ~12 min):
#   ./grun.sh 1500 'bash tools/round2_first_call.sh'
# 1. the parked validations of the CPU-only additions
# 2. the whole GPU suite
# 3. the default bench line, then the same with the Rainbow update graphed
# 4. secondary workloads (last)
# Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
B2RL_PENDING=1 python -m pytest tests/test_zz_pending_validation_gpu.py -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pending.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02_gpu_tests.log
python bench.py 2>gpurun_out/r02_bench.err | tail -1 | tee gpurun_out/r02_bench_default.json
python bench.py --rainbow-graph --no-cpu-baseline 2>>gpurun_out/r02_bench.err | tail -1 | tee gpurun_out/r02_bench_rainbow_graph.json
# 5. the opt-in FMA descent of the exact sampler: bit-identity tests, then the bench line
B2RL_SAMPLER_DESCENT=fma python -m pytest tests/test_per_gpu.py tests/test_replay_buffers_gpu.py -q -x 2>&1 | tail -5 | tee gpurun_out/r02_fma_tests.log
B2RL_SAMPLER_DESCENT=fma python bench.py --no-cpu-baseline --no-rainbow 2>>gpurun_out/r02_bench.err | tail -1 | tee gpurun_out/r02_bench_fma.json
python tools/bench_secondary.py 2>>gpurun_out/r02_bench.err | tee gpurun_out/r02_secondary.jsonl
