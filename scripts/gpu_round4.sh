#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log | cut -c1-1800
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; head -c 400 gpurun_out/bench_full.json; echo; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --workload painter --no-cpu-baseline > gpurun_out/bench_painter.json 2> gpurun_out/bench_painter.err
echo "bench rc=$?"; head -c 300 gpurun_out/bench_painter.json; echo
