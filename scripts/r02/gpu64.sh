#!/bin/bash
# round 2, call 64: is the bf16 graph-replay test's iteration-0 failure run-to-run noise?  three runs with and without the operand-by-TMA epilogue
mkdir -p gpurun_out
for m in 1 0; do for i in 1 2 3; do
  echo "== CGB_AUX_TMA=$m run $i"
  CGB_AUX_TMA=$m timeout 300 python -m pytest tests/test_gpu_graphs.py -q -m gpu --tb=line -k "dtype1" 2>&1 | grep -E "passed|failed|AssertionError" | cut -c1-250
done; done | tee gpurun_out/g64_graphs_flake.txt
