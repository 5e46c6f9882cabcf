#!/bin/bash
# round-end evidence: whole GPU suite, smoke, reference arm, default bench, launch list, ncu --set full captures
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/final.log
  timeout -s KILL $to "$@" > gpurun_out/final_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/final_${name}.log | tr '\n' ' ' | cut -c1-300)" | tee -a gpurun_out/final.log; }
: > gpurun_out/final.log
run tests 1500 python -m pytest tests -m gpu -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run ref 900 python bench.py --impl reference --steps 20 --warmup 5
run bench 1200 python bench.py --steps 20 --warmup 5
run k10 600 bash -c "python tools/bench_gemm.py; python tools/bench_conv.py; python tools/gemm_phases.py"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 1500 --csv \
  --log-file gpurun_out/r02_launches_ncu.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 --passes 6 > gpurun_out/final_ncu_list.log 2>&1
B="python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 4"
cap() { local out=$1 regex=$2 skip=$3; shift 3
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f \
    -o gpurun_out/$out "$@" > gpurun_out/final_ncu_$out.log 2>&1; echo "$out rc=$?" | tee -a gpurun_out/final.log; }
cap r02_step_exact 'k_replay_step<9' 4 $B
cap r02_step_parallel 'k_replay_step<0, 2>' 4 $B
cap r02_sampler_v6 'k_sample_exact_v6' 2 $B
cap r02_gather '^k_gather|k_gather\(' 2 $B
cap r02_update_multi 'k_update_multi' 4 $B
cap r02_gemm 'k_gemm_tf32x3' 6 python tools/bench_gemm.py
cap r02_gae 'k_gae' 1 python tools/bench_secondary.py
cap r02_ppo_loss 'k_ppo_loss' 3 python tools/bench_secondary.py
cap r02_polyak 'k_polyak' 3 python tools/bench_secondary.py
cap r02_sac_target 'k_sac_target' 3 python tools/bench_secondary.py
cap r02_conv1_u8 'k_conv_nature1' 3 python bench.py --no-cpu-baseline --no-secondary --steps 1 --warmup 1 --passes 2
ls -la gpurun_out/*.ncu-rep | tail -14
