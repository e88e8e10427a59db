#!/bin/bash
# round 2, call 73: grid cap of the flat elementwise kernels (grid_for: 16 CTAs of 256 threads per SM = two waves) against one wave
mkdir -p gpurun_out
{
for k in 16 8 4; do
  echo "== CGB_FLAT_CTAS=$k"
  CGB_FLAT_CTAS=$k timeout 300 python scripts/bench_hbm_kernels.py 2>&1 | grep -vi "bn_\|spade\|in_stats\|in_bwd\|Warn"
done
} | tee gpurun_out/g73_flat_grid.txt
for k in 16 8; do
  CGB_FLAT_CTAS=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g73_full_$k.err | tee gpurun_out/g73_full_$k.json | cut -c1-200
done
