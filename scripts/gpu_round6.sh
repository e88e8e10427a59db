#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-1200
python scripts/profile_full_step.py > gpurun_out/profile_step.log 2>&1
grep -A 42 "host side" gpurun_out/profile_step.log | cut -c1-150; tail -2 gpurun_out/profile_step.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; head -c 330 gpurun_out/bench_full.json; echo
