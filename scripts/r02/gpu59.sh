#!/bin/bash
# round 2, call 59: bulk L2 prefetch of the next tile's mask (A/B), the same dgrad without a mask on both copy-out paths, timeline
mkdir -p gpurun_out
{
for m in 0 3; do
  echo "== CGB_MASK_PF=$m relu / lrelu"
  CGB_MASK_PF=$m REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 vgg1d 2>&1 | grep -v Warning
  DACT=lrelu CGB_MASK_PF=$m REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 2>&1 | grep -v Warning
done
echo "== no mask, per-thread copy-out (CGB_TMA_STORE=0)"
DACT=none CGB_TMA_STORE=0 REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 2>&1 | grep -v Warning
echo "== no mask, TMA store"
DACT=none REPS=20 timeout 300 python scripts/bench_conv.py dg48 dg80 2>&1 | grep -v Warning
CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg48 2>&1 | grep -v Warn | tail -8 | cut -c1-200
DACT=none CGB_TMA_STORE=0 CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg48 2>&1 | grep -v Warn | tail -8 | cut -c1-200
DACT=none CGB_TC_TRACE=1 REPS=1 timeout 120 python scripts/bench_conv.py dg48 2>&1 | grep -v Warn | tail -8 | cut -c1-200
} | tee gpurun_out/g59_ab.txt
