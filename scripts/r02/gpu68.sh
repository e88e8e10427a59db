#!/bin/bash
# round 2, call 68: bias gradient out of the fused activation's backward pass (cgb_act_bwd_bias): tests, A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_painter.py tests/test_gpu_full_step.py tests/test_gpu_discriminator.py tests/test_gpu_masker.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/g68_unit.log
for m in 0 1; do
  CGB_ACT_BIAS_FUSED=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g68_full_$m.err | tee gpurun_out/g68_full_$m.json | cut -c1-200
done
CGB_ACT_BIAS_FUSED=1 timeout 600 python bench.py --workload masker --steps 5 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g68_masker.err | tee gpurun_out/g68_masker.json | cut -c1-200
