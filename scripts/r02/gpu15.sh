#!/bin/bash
# round 2, call 15: per-role pipeline timeline (CGB_TC_TRACE) of the short-K 1x1 convs and the dilated 3x3
mkdir -p gpurun_out
for c in r1 r1b r3 sh8; do
  CGB_TC_TRACE=1 CGB_TC2=0 REPS=2 timeout 120 python scripts/bench_conv.py $c > gpurun_out/g15_trace_$c.txt 2>&1
  tail -40 gpurun_out/g15_trace_$c.txt | cut -c1-200
done
