#!/bin/bash
# round 2, call 74: instance-norm backward with two pixels in flight per thread, statistics loop unrolled 4x
mkdir -p gpurun_out
timeout 300 python scripts/bench_hbm_kernels.py 2>&1 | grep -i "in_stats\|in_bwd" | tee gpurun_out/g74_in.txt
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_painter.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g74_full.err | tee gpurun_out/g74_full.json | cut -c1-200
