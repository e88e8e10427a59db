#!/bin/bash
# round 2, call 37: conv1 + skip gradient in one dgrad launch (ops.conv2d_skip, EPI_RES): tests + full / masker bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_skip.py tests/test_gpu_masker.py tests/test_gpu_full_step.py tests/test_gpu_graphs.py tests/test_gpu_full_size.py -q -m gpu --tb=short -x > gpurun_out/g37_unit.log 2>&1; tail -3 gpurun_out/g37_unit.log | cut -c1-300
for v in 0 1; do
CGB_CONV_SKIP=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g37_bench_full_skip$v.json 2> gpurun_out/g37_bench_full_skip$v.err
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.loads(open(f"gpurun_out/g37_bench_full_skip{v}.json").read().strip().splitlines()[-1])
    print("CONV_SKIP", v, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; launches/step", d.get("gpu_launches_per_step"))
PY
