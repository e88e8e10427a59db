#!/bin/bash
# round 2, call 41: wgrad with M = output channels where co >= 128: parity tests + full / masker bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_masker.py tests/test_gpu_full_step.py -q -m gpu --tb=short -x > gpurun_out/g41_unit.log 2>&1; tail -3 gpurun_out/g41_unit.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g41_bench_full.json 2> gpurun_out/g41_bench_full.err
timeout 600 python bench.py --workload masker --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g41_bench_masker.json 2> gpurun_out/g41_bench_masker.err
python - <<'PY'
import json
for name in ("full", "masker"):
    d = json.loads(open(f"gpurun_out/g41_bench_{name}.json").read().strip().splitlines()[-1])
    print(name, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; step_frac", d["roofline"].get("step_frac"), "conv", d["roofline"].get("conv_aggregate"))
PY
