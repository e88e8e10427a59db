#!/bin/bash
# round 2, call 21: ncu --set full of the streaming kernel on the 256->1024 1x1 conv (rolling-store epilogue)
mkdir -p gpurun_out
NOBIAS=1 CGB_TC2=0 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/g21_r1 -f python scripts/bench_conv.py r1 > gpurun_out/g21_ncu.log 2>&1; tail -3 gpurun_out/g21_ncu.log
