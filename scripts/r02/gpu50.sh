#!/bin/bash
# round 2, call 50: wgrad accumulating straight into param.grad for leaf 1x1 weights: step tests + A/B bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_full_step.py tests/test_gpu_graphs.py tests/test_gpu_masker.py tests/test_gpu_masker_v3.py tests/test_gpu_trainer.py tests/test_gpu_sweep.py tests/test_gpu_full_size.py -q -m gpu --tb=short > gpurun_out/g50_unit.log 2>&1; tail -3 gpurun_out/g50_unit.log | cut -c1-300
for v in 0 1; do
CGB_WGRAD_INPLACE=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g50_bench_full_inplace$v.json 2> gpurun_out/g50_bench_full_inplace$v.err
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.loads(open(f"gpurun_out/g50_bench_full_inplace{v}.json").read().strip().splitlines()[-1])
    print("WGRAD_INPLACE", v, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; launches/step", d.get("gpu_launches_per_step"))
PY
