#!/bin/bash
# round 2, call 5: fp16 mode tests, HBM-kernel microbenchmark (flushed + warm L2), ncu of the BatchNorm passes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_half.py tests/test_gpu_fused_stats.py -q -m gpu --tb=short -s > gpurun_out/g5_half.log 2>&1; tail -30 gpurun_out/g5_half.log | cut -c1-250
timeout 600 python scripts/bench_hbm_kernels.py --json gpurun_out/g5_hbm_flushed.json > gpurun_out/g5_hbm_flushed.txt 2>&1; cat gpurun_out/g5_hbm_flushed.txt
timeout 600 python scripts/bench_hbm_kernels.py --warm --json gpurun_out/g5_hbm_warm.json > gpurun_out/g5_hbm_warm.txt 2>&1; cat gpurun_out/g5_hbm_warm.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bn_apply -c 6 -o gpurun_out/g5_ncu_bn python scripts/bench_hbm_kernels.py --only "8x80x80x256" --reps 1 > gpurun_out/g5_ncu_bn.log 2>&1; tail -3 gpurun_out/g5_ncu_bn.log
