#!/bin/bash
mkdir -p gpurun_out
for v in 64 48 1000; do
echo "CGB_WS_MAX_CO=$v"
CGB_WS_MAX_CO=$v REPS=20 timeout 300 python scripts/bench_conv.py dg48 gb48 sn24 d3 2>&1 | grep -v Warn
CGB_WS_MAX_CO=$v timeout 600 python bench.py --workload painter --no-cpu-baseline --no-e2e 2>/dev/null | head -c 230; echo
done
timeout 900 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | head -c 230; echo
