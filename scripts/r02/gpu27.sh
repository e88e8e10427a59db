#!/bin/bash
# round 2, call 27: lean per-thread copy-out path (ReLU-masked dgrads)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_half.py -q -m gpu --tb=short -x > gpurun_out/g27_unit.log 2>&1; tail -3 gpurun_out/g27_unit.log | cut -c1-300
REPS=10 timeout 300 python scripts/bench_conv.py dg48 r3d vgg3d wg48 wgsh r3w r1w 2>&1 | tail -8
