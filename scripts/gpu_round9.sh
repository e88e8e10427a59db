#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-1200
timeout 600 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "bench rc=$?"; head -c 330 gpurun_out/bench_infer.json; echo; tail -3 gpurun_out/bench_infer.err
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; head -c 330 gpurun_out/bench_full.json; echo
