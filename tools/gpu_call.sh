#!/bin/bash
# usage: tools/gpu_call.sh <tag> -- runs the steps listed in $STEPS (space separated) on the GPU box
set -u
mkdir -p gpurun_out
tag=$1
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "== $name" | tee -a gpurun_out/${tag}.log
  timeout -s KILL $to "$@" > gpurun_out/${tag}_${name}.log 2>&1
  echo "rc=$? $(tail -3 gpurun_out/${tag}_${name}.log | tr '\n' ' ')" | tee -a gpurun_out/${tag}.log
}
for s in $STEPS; do
  case $s in
    fused)    run fused 600 python -m pytest tests/test_fused_step_gpu.py -x -q ;;
    per)      run per 600 python -m pytest tests/test_per_gpu.py tests/test_replay_buffers_gpu.py -x -q ;;
    headline) run headline 900 python -m pytest tests/test_headline_shapes_gpu.py -x -q ;;
    all)      run all 1500 python -m pytest tests -m gpu -x -q ;;
    smoke)    run smoke 300 python -c "import __graft_entry__ as g; g.smoke()" ;;
    bench)    run bench 900 python bench.py ;;
    benchq)   run benchq 600 python bench.py --no-cpu-baseline --no-rainbow --no-secondary ;;
  esac
done
