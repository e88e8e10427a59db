#!/bin/bash
# round 2, call 61: L2 tensor prefetch two tiles ahead (single staging tile), weight-stationary slices of 64 channels, N = 128 slices for K = 80
mkdir -p gpurun_out
{
for l in 0 1; do
  echo "== CGB_AUX_L2=$l"
  CGB_AUX_L2=$l REPS=20 timeout 300 python scripts/bench_conv.py dg48 vgg1d 2>&1 | grep -v Warning
  DACT=lrelu CGB_AUX_L2=$l REPS=20 timeout 300 python scripts/bench_conv.py dg48 2>&1 | grep -v Warning
done
echo "== CGB_WS_MIN_NT=2 (two 64-channel slices, two staging tiles)"
CGB_WS_MIN_NT=2 REPS=20 timeout 300 python scripts/bench_conv.py dg48 vgg1d 2>&1 | grep -v Warning
echo "== CGB_WS_MAX_CO=128 (K = 80: two 64-channel slices instead of the streaming kernel)"
CGB_WS_MAX_CO=128 REPS=20 timeout 300 python scripts/bench_conv.py dg80 gb80_640 2>&1 | grep -v Warning
CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg48 2>&1 | grep -v Warn | tail -6 | cut -c1-200
CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg80 2>&1 | grep -v Warn | tail -6 | cut -c1-200
} | tee gpurun_out/g61_ab.txt
