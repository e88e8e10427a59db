#!/bin/bash
# round 2, call 69: tiled max-pool backward: parity with F.max_pool2d's backward, masker tests, A/B on the full step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_masker_ops.py tests/test_gpu_masker.py tests/test_gpu_masker_v3.py tests/test_gpu_full_step.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/g69_unit.log
for m in 0 1; do
  CGB_MAXPOOL_TILED=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g69_full_$m.err | tee gpurun_out/g69_full_$m.json | cut -c1-200
done
