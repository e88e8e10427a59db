#!/bin/bash
mkdir -p gpurun_out
REPS=20 timeout 300 python scripts/bench_conv.py sh8 gb48_8 > gpurun_out/bench_conv2.log 2>&1; cat gpurun_out/bench_conv2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 2 -o gpurun_out/prof_sh8 python scripts/bench_conv.py sh8 > gpurun_out/ncu_sh8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 2 -o gpurun_out/prof_gb48_8 python scripts/bench_conv.py gb48_8 > gpurun_out/ncu_gb48_8.log 2>&1
ls -la gpurun_out/*.ncu-rep
