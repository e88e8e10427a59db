#!/bin/bash
# round 2, call 71: grid size of the BatchNorm forward / finalize passes: CTAs per SM (8 = round-2 default; launch bounds allow 6 / 5 resident)
mkdir -p gpurun_out
{
for k in 8 6 4 3; do
  echo "== CGB_BN_FWD_CTAS=$k CGB_BN_FIN_CTAS=$k"
  CGB_BN_FWD_CTAS=$k CGB_BN_FIN_CTAS=$k timeout 300 python scripts/bench_hbm_kernels.py --only bn_ 2>&1 | grep -i "bn_apply_fwd\|bn_bwd_finalize"
done
} | tee gpurun_out/g71_bn_grid.txt
for k in 8 5; do
  CGB_BN_FWD_CTAS=$k CGB_BN_FIN_CTAS=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e 2> gpurun_out/g71_full_$k.err | tee gpurun_out/g71_full_$k.json | cut -c1-200
done
