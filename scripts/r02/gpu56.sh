#!/bin/bash
# round 2, call 56: A/B of the unrolled 3x3 MMA issue in the weight-stationary kernel (CGB_WS_UNROLL), unit tests, painter bench
mkdir -p gpurun_out
for u in 0 1; do
  echo "== CGB_WS_UNROLL=$u"
  CGB_WS_UNROLL=$u REPS=20 timeout 300 python scripts/bench_conv.py gb48_8 gb48 gb80 gb160 sn24 dg48 2>&1 | grep -v Warning
done | tee gpurun_out/g56_ab.txt
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_painter.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/g56_unit.log
for u in 0 1; do
  CGB_WS_UNROLL=$u timeout 600 python bench.py --workload painter --steps 10 --warmup 3 2> gpurun_out/g56_painter_$u.err | tee gpurun_out/g56_painter_$u.json | cut -c1-300
done
