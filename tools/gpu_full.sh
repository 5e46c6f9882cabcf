#!/bin/bash
# round-end style run: whole GPU suite, smoke, reference arm, default bench, launch list + full captures
set -u
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; echo "== $name" | tee -a gpurun_out/full.log
  timeout -s KILL $to "$@" > gpurun_out/full_${name}.log 2>&1
  echo "rc=$? $(tail -2 gpurun_out/full_${name}.log | tr '\n' ' ' | cut -c1-300)" | tee -a gpurun_out/full.log; }
run tests 1500 python -m pytest tests -m gpu -q
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run ref 900 python bench.py --impl reference --steps 20 --warmup 5
run bench 900 python bench.py --steps 20 --warmup 5
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 900 --csv \
  --log-file gpurun_out/r02_launches_ncu.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 --passes 6 > gpurun_out/full_ncu_list.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_replay_step" -s 20 -c 2 \
  -o gpurun_out/r02_step_exact python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 6 > gpurun_out/full_ncu_a.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_replay_stepILi0ELi2|k_replay_step<0, 2>" -s 4 -c 2 \
  -o gpurun_out/r02_step_parallel python bench.py --no-cpu-baseline --no-rainbow --no-secondary --steps 1 --warmup 1 --passes 6 > gpurun_out/full_ncu_b.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"k_update_multi|k_gae|k_ppo_loss|k_polyak|k_sac_target|k_conv_nature1" -c 12 \
  -o gpurun_out/r02_small python bench.py --no-cpu-baseline --steps 1 --warmup 1 --passes 4 > gpurun_out/full_ncu_c.log 2>&1
ls -la gpurun_out | tail -6
