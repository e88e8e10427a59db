#!/bin/bash
# round 2, call 25: whole GPU suite + full-step bench on the new epilogue (compile-time variants, rolling store, pair from K=512)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g25_pytest.log 2>&1; tail -8 gpurun_out/g25_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g25_bench_full.json 2> gpurun_out/g25_bench_full.err; tail -c 300 gpurun_out/g25_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/g25_bench_full.json").read().strip().splitlines()[-1])
print("full:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
print("conv", d["roofline"]["conv_aggregate"], "step_frac", d["roofline"]["step_frac"])
PY
