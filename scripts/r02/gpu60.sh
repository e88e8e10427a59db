#!/bin/bash
# round 2, call 60: second epilogue operand by TMA into the staging tile (CGB_AUX_TMA): parity at size, A/B, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/g60_unit.log
{
for m in 0 1; do
  echo "== CGB_AUX_TMA=$m relu / lrelu"
  CGB_AUX_TMA=$m REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 vgg1d vgg2d vgg3d r3d 2>&1 | grep -v Warning
  DACT=lrelu CGB_AUX_TMA=$m REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 2>&1 | grep -v Warning
done
CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg48 2>&1 | grep -v Warn | tail -8 | cut -c1-200
} | tee gpurun_out/g60_ab.txt
for m in 0 1; do
  CGB_AUX_TMA=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g60_full_$m.err | tee gpurun_out/g60_full_$m.json | cut -c1-300
done
