#!/bin/bash
# round 2, call 38: wgrad split count — one wave of CTAs vs two (CGB_WG_WAVES)
mkdir -p gpurun_out
for v in 2 1 3; do echo "CGB_WG_WAVES=$v"; ONLY=w_r3,w_r1,w_r1b,w_r4,w_l4,w_aspp,w_s2,w_128 CGB_WG_WAVES=$v timeout 300 python scripts/exp/tc2_check.py save 2>&1 | tail -8; done
for v in 2 1; do echo "CGB_WG_WAVES=$v"; CGB_WG_WAVES=$v REPS=10 timeout 300 python scripts/bench_conv.py wg48 wgsh 2>&1 | tail -2; done
