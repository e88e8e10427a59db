#!/bin/bash
# round 2, call 31: C1 painter profile (all kernels) + all conv classes of the painter workload
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py > gpurun_out/g31_profile_painter.txt 2>&1; head -52 gpurun_out/g31_profile_painter.txt | cut -c1-160
timeout 600 python bench.py --workload painter --steps 6 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g31_bench_painter.json 2> gpurun_out/g31_bench_painter.err
