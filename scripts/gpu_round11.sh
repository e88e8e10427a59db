#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_infer_all.py -q -m gpu --tb=short > gpurun_out/pytest_gpu_sub.log 2>&1
tail -3 gpurun_out/pytest_gpu_sub.log | cut -c1-600
timeout 600 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "bench rc=$?"; head -c 200 gpurun_out/bench_infer.json; echo; tail -3 gpurun_out/bench_infer.err
CGB_TOPK=400 timeout 900 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/bench_full_top.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; head -c 200 gpurun_out/bench_full_top.json; echo
