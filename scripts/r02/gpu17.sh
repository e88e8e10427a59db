#!/bin/bash
# round 2, call 17/19: lean epilogue phase 1 + rolling per-half stores: conv parity tests, timeline, single-kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_fused_stats.py tests/test_gpu_half.py -q -m gpu --tb=short -x > gpurun_out/g17_unit.log 2>&1; tail -5 gpurun_out/g17_unit.log | cut -c1-300
for c in r1 sh8; do
  NOBIAS=1 CGB_TC_TRACE=1 CGB_TC2=0 REPS=2 timeout 120 python scripts/bench_conv.py $c 2>&1 | tail -8 | cut -c1-220
done
REPS=20 timeout 300 python scripts/bench_conv.py r1 r1b r3 r3d sh8 gb48_8 aspp vgg3 vgg3d r4 d3 2>&1 | tail -12
