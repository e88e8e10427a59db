#!/bin/bash
# round 2, call 62: N = 128 tiles for the short-K 1x1 convs of the ResNet (five 32 KB stages, two staging tiles) against N = 256
mkdir -p gpurun_out
{
for b in 256 128; do
  echo "== CGB_BN_MAX_1X1=$b"
  NOBIAS=1 CGB_BN_MAX_1X1=$b CGB_BN_MAX_K=1024 REPS=50 timeout 300 python scripts/bench_conv.py r1 r1b sh8 2>&1 | grep -v Warning
done
NOBIAS=1 CGB_BN_MAX_1X1=128 CGB_BN_MAX_K=1024 CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py r1 2>&1 | grep -v Warn | tail -8 | cut -c1-200
NOBIAS=1 CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py r1 2>&1 | grep -v Warn | tail -8 | cut -c1-200
} | tee gpurun_out/g62_ab.txt
