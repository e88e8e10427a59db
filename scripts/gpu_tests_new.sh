#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_full_step.py -q -m gpu --tb=short > gpurun_out/pytest_new.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_new.log
tail -40 gpurun_out/pytest_new.log | cut -c1-1500
timeout 600 python scripts/diag_full_step.py > gpurun_out/diag_full_step.log 2>&1
tail -5 gpurun_out/diag_full_step.log
