#!/bin/bash
# round 2, call 55: timeline of the weight-stationary kernel: gamma||beta 128->48 @640^2 (n=8), 24->24, and the resident-weight 1x1
mkdir -p gpurun_out
for c in gb48_8 sn24; do CGB_TC_TRACE=1 REPS=2 timeout 120 python scripts/bench_conv.py $c 2>&1 | tail -14 | cut -c1-200; done
NOBIAS=1 CGB_WS_1X1=1 CGB_WS_1X1_MIN_STAGES=6 CGB_TC2=0 CGB_TC_TRACE=1 REPS=2 timeout 120 python scripts/bench_conv.py r1 2>&1 | tail -14 | cut -c1-200
