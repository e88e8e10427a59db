#!/bin/bash
# round 2, call 23: compile-time epilogue variants + rolling store: conv parity tests, timeline, single-kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_fused_stats.py tests/test_gpu_half.py -q -m gpu --tb=short -x > gpurun_out/g23_unit.log 2>&1; tail -5 gpurun_out/g23_unit.log | cut -c1-300
NOBIAS=1 CGB_TC_TRACE=1 CGB_TC2=0 REPS=2 timeout 120 python scripts/bench_conv.py r1 2>&1 | tail -6 | cut -c1-250
NOBIAS=1 REPS=20 timeout 300 python scripts/bench_conv.py r1 r1b r3 r3d 2>&1 | tail -4
REPS=20 timeout 300 python scripts/bench_conv.py sh8 gb48_8 aspp vgg3 vgg3d r4 d3 2>&1 | tail -7
