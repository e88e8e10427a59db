#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-1200
REPS=20 timeout 300 python scripts/bench_conv.py r3 r3d r3w r1 r1w r1b stem aspp gb48 gb80 > gpurun_out/bench_conv.log 2>&1; cat gpurun_out/bench_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 2 -o gpurun_out/prof_r3 python scripts/bench_conv.py r3 > gpurun_out/ncu_r3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 2 -o gpurun_out/prof_r1w python scripts/bench_conv.py r1w > gpurun_out/ncu_r1w.log 2>&1
ls -la gpurun_out/*.ncu-rep
