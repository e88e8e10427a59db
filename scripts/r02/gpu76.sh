#!/bin/bash
# round 2, call 76: ncu launch list of the bench command on the final tree (eager step: one launch per kernel)
mkdir -p gpurun_out
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 11000 --csv --log-file gpurun_out/g76_launches_full.csv python bench.py --steps 1 --warmup 1 --no-graphs --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/g76_launches_bench.log 2>&1; tail -2 gpurun_out/g76_launches_bench.log | cut -c1-200; wc -l gpurun_out/g76_launches_full.csv
