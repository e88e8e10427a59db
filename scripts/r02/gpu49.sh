#!/bin/bash
# round 2, call 49: per-pixel im2col incl. the stem: masker tests + profile + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_masker.py tests/test_gpu_masker_v3.py tests/test_gpu_fused_stats.py tests/test_first_layer_im2col.py tests/test_gpu_full_step.py -q -m gpu --tb=short > gpurun_out/g49_unit.log 2>&1; tail -3 gpurun_out/g49_unit.log | cut -c1-300
timeout 600 python scripts/profile_full_step.py > gpurun_out/g49_profile_full.txt 2>&1; grep -n "total CUDA\|im2col\|col2im" gpurun_out/g49_profile_full.txt | cut -c1-150
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g49_bench_full.json 2> gpurun_out/g49_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/g49_bench_full.json").read().strip().splitlines()[-1])
print("full", round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; launches/step", d.get("gpu_launches_per_step"))
PY
