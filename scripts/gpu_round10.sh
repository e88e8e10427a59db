#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_masker.py tests/test_gpu_infer_all.py -q -m gpu --tb=short > gpurun_out/pytest_gpu_sub.log 2>&1
tail -4 gpurun_out/pytest_gpu_sub.log | cut -c1-600
CGB_TOPK=40 timeout 600 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
echo "bench rc=$?"; head -c 330 gpurun_out/bench_infer.json; echo; tail -3 gpurun_out/bench_infer.err
