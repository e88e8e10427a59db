#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-1500
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_full_under_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/profile_full_step.py > gpurun_out/profile_step.log 2>&1 || true
tail -40 gpurun_out/profile_step.log
