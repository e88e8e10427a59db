#!/bin/bash
# round 2, call 72: CTAs per SM of the chunked SPADE / instance-norm passes (pick_chunks), BatchNorm grid defaults re-measured, full step
mkdir -p gpurun_out
{
for k in 8 6 4 3; do
  echo "== CGB_CHUNK_CTAS=$k"
  CGB_CHUNK_CTAS=$k timeout 300 python scripts/bench_hbm_kernels.py 2>&1 | grep -i "spade\|in_\|instnorm"
done
echo "== BatchNorm passes, new defaults"
timeout 300 python scripts/bench_hbm_kernels.py --only bn_ 2>&1 | grep -i "bn_"
} | tee gpurun_out/g72_chunk_grid.txt
for k in 8 4; do
  CGB_CHUNK_CTAS=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g72_full_$k.err | tee gpurun_out/g72_full_$k.json | cut -c1-200
done
