#!/bin/bash
# round 2, call 70: BatchNorm backward: one resident wave of CTAs (444 partial rows instead of 1184), 32 row groups in the fp64 fold
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_masker_ops.py tests/test_bn_dual.py tests/test_gpu_fused_stats.py tests/test_gpu_masker.py tests/test_gpu_full_step.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/g70_unit.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g70_full.err | tee gpurun_out/g70_full.json | cut -c1-200
timeout 300 python scripts/bench_hbm_kernels.py 2>&1 | grep -i "bn_" | head -12 | tee gpurun_out/g70_hbm_bn.txt
