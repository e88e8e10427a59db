#!/bin/bash
# round 2, call 42: resident-weight 1x1 convs (CGB_WS_1X1=1) re-measured with the rebuilt epilogue
mkdir -p gpurun_out
for v in 0 1; do echo "CGB_WS_1X1=$v"; ONLY=r1,r1b,l4,l4b,r1d CGB_WS_1X1=$v timeout 300 python scripts/exp/tc2_check.py $([ $v = 0 ] && echo save || echo check) 2>&1 | tail -10; done
echo "CGB_WS_1X1=1 MIN_STAGES=6"; ONLY=r1,r1b,l4,l4b,r1d CGB_WS_1X1=1 CGB_WS_1X1_MIN_STAGES=6 timeout 300 python scripts/exp/tc2_check.py check 2>&1 | tail -10
