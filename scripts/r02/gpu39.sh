#!/bin/bash
# round 2, call 39: wgrad split count from a wave cost model: single kernels, wgrad parity tests, full-step bench
mkdir -p gpurun_out
ONLY=w_r3,w_r1,w_r1b,w_r4,w_l4,w_aspp,w_s2,w_128 timeout 300 python scripts/exp/tc2_check.py save 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_masker.py -q -m gpu --tb=short -x > gpurun_out/g39_unit.log 2>&1; tail -3 gpurun_out/g39_unit.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g39_bench_full.json 2> gpurun_out/g39_bench_full.err
timeout 600 python bench.py --workload masker --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g39_bench_masker.json 2> gpurun_out/g39_bench_masker.err
python - <<'PY'
import json
for name in ("full", "masker"):
    d = json.loads(open(f"gpurun_out/g39_bench_{name}.json").read().strip().splitlines()[-1])
    print(name, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; step_frac", d["roofline"].get("step_frac"), "conv", d["roofline"].get("conv_aggregate"))
PY
