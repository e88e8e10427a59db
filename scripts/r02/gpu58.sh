#!/bin/bash
# round 2, call 58: per-role timelines (CGB_TC_TRACE) of the gamma||beta dgrads at 640^2 (weight-stationary kernel, K = 48) and of the 128->80 fprop
mkdir -p gpurun_out
{
for c in dg48 dg80 gb80_640; do CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py $c 2>&1 | grep -v Warn | tail -16 | cut -c1-200; done
DACT=lrelu CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg48 2>&1 | grep -v Warn | tail -16 | cut -c1-200
} | tee gpurun_out/g58_trace.txt
