#!/bin/bash
# round 2, call 67: mlp_gamma / mlp_beta bias gradients out of the SPADE modulation backward pass (cgb_spade_modulate_bwd_bias): tests, A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_painter.py tests/test_gpu_full_step.py tests/test_gpu_masker.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/g67_unit.log
for m in 0 1; do
  CGB_SPADE_BIAS_FUSED=$m timeout 600 python bench.py --workload painter --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g67_painter_$m.err | tee gpurun_out/g67_painter_$m.json | cut -c1-200
  CGB_SPADE_BIAS_FUSED=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g67_full_$m.err | tee gpurun_out/g67_full_$m.json | cut -c1-200
done
