#!/bin/bash
# round 2, call 16: finer epilogue timeline (stamps inside phase 1), without bias
mkdir -p gpurun_out
for c in r1 sh8; do
  NOBIAS=1 CGB_TC_TRACE=1 CGB_TC2=0 REPS=2 timeout 120 python scripts/bench_conv.py $c 2>&1 | tail -14 | cut -c1-220
done
