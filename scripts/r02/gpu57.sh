#!/bin/bash
# round 2, call 57: A/B of the next-tile mask prefetch into L2 (CGB_MASK_PF) and of the weight-stationary kernel for N = 128 slices
mkdir -p gpurun_out
{
for m in 0 3; do
  echo "== CGB_MASK_PF=$m (relu mask, per-thread copy-out)"
  CGB_MASK_PF=$m REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 vgg1d vgg2d vgg3d r3d 2>&1 | grep -v Warning
  echo "== CGB_MASK_PF=$m (lrelu mask, TMA store)"
  DACT=lrelu CGB_MASK_PF=$m REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 vgg2d 2>&1 | grep -v Warning
done
echo "== CGB_WS_MAX_CO=128 (weight-stationary slices for the 80-channel gamma||beta pair)"
CGB_WS_MAX_CO=128 REPS=20 timeout 300 python scripts/bench_conv.py gb80_640 dg80 2>&1 | grep -v Warning
echo "== default"
REPS=20 timeout 300 python scripts/bench_conv.py gb80_640 wg80 wg48 2>&1 | grep -v Warning
} | tee gpurun_out/g57_ab.txt
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_painter.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/g57_unit.log
for m in 0 3; do
  CGB_MASK_PF=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g57_full_$m.err | tee gpurun_out/g57_full_$m.json | cut -c1-300
done
