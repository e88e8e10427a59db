#!/bin/bash
# round 2, call 47/48: per-pixel im2col kernel: tests + profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_first_layer_im2col.py tests/test_gpu_ops.py tests/test_gpu_painter.py tests/test_gpu_masker.py tests/test_gpu_discriminator.py tests/test_gpu_infer_all.py -q -m gpu --tb=short > gpurun_out/g47_unit.log 2>&1; tail -3 gpurun_out/g47_unit.log | cut -c1-300
timeout 600 python scripts/profile_full_step.py > gpurun_out/g47_profile_full.txt 2>&1; grep -n "total CUDA\|im2col\|col2im" gpurun_out/g47_profile_full.txt | cut -c1-150
