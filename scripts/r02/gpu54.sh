#!/bin/bash
# round 2, call 54: compute-sanitizer memcheck over the conv unit tests (new epilogue, im2col / col2im, dual BatchNorm backward)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py tests/test_first_layer_im2col.py tests/test_conv_skip.py tests/test_bn_dual.py -q -m gpu --tb=line -x -k "not 640" > gpurun_out/g54_memcheck.log 2>&1; echo "rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/g54_memcheck.log; tail -6 gpurun_out/g54_memcheck.log | cut -c1-200
