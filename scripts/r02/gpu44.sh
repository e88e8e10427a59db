#!/bin/bash
# round 2, call 44: tests of the dual BatchNorm backward on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bn_dual.py tests/test_conv_skip.py tests/test_gpu_masker.py tests/test_gpu_full_step.py tests/test_gpu_graphs.py tests/test_gpu_full_size.py tests/test_gpu_masker_ops.py tests/test_gpu_masker_v3.py -q -m gpu --tb=short > gpurun_out/g44_unit.log 2>&1; tail -5 gpurun_out/g44_unit.log | cut -c1-300
