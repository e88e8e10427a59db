#!/bin/bash
# second A/B: TMA store default-on, extended to single-N-tile convs and the weight-stationary kernel
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
CASES="sh8 gb48_8 gb48 gb80 gb160 sn24 r1 dg48"
REPS=20 CGB_TMA_STORE=0 timeout 120 python scripts/bench_conv.py $CASES > gpurun_out/tma_store2_conv_off.log 2>&1
REPS=20 timeout 120 python scripts/bench_conv.py $CASES > gpurun_out/tma_store2_conv_on.log 2>&1
paste -d'|' gpurun_out/tma_store2_conv_off.log gpurun_out/tma_store2_conv_on.log | cut -c1-230
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/tma_store2_bench_on.json 2> gpurun_out/tma_store2_bench_on.err
python - <<PY
import json
d = json.loads(open("gpurun_out/tma_store2_bench_on.json").read().strip().splitlines()[-1])
print("on", d["value"], d["ms_per_step"], d.get("e2e"), d["roofline"]["kernel"], d["roofline"]["frac"])
PY
