#!/bin/bash
# round 2, call 40: wgrad operand roles — M = output channels (vector reds) whenever co >= 128
mkdir -p gpurun_out
for v in 1 0; do echo "CGB_WG_XM=$v"; ONLY=w_r3,w_r1,w_r1b,w_r4,w_l4,w_aspp,w_s2,w_128 CGB_WG_XM=$v timeout 300 python scripts/exp/tc2_check.py $([ $v = 1 ] && echo save || echo check) 2>&1 | tail -16; done
