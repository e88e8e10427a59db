#!/bin/bash
# round 2, call 33: evidence — ncu launch list of the bench command (eager step: one launch per kernel), ncu --set full of the three
# largest conv classes, torch.profiler of the final tree, SASS census
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 14000 --csv --log-file gpurun_out/g33_launches_full.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/g33_launches_bench.log 2>&1; tail -2 gpurun_out/g33_launches_bench.log | cut -c1-200; wc -l gpurun_out/g33_launches_full.csv
timeout 600 python scripts/profile_full_step.py > gpurun_out/g33_profile_full.txt 2>&1; head -4 gpurun_out/g33_profile_full.txt | cut -c1-160
for c in r1 r3 sh8; do
  NOBIAS=$([ $c = sh8 ] && echo "" || echo 1) REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/g33_$c -f python scripts/bench_conv.py $c > gpurun_out/g33_ncu_$c.log 2>&1; tail -1 gpurun_out/g33_ncu_$c.log
done
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad -s 1 -c 1 -o gpurun_out/g33_r3w -f python scripts/bench_conv.py r3w > gpurun_out/g33_ncu_r3w.log 2>&1; tail -1 gpurun_out/g33_ncu_r3w.log
