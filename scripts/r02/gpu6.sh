#!/bin/bash
# round 2, call 6: low-register BatchNorm passes + few-channel layout kernels: unit tests, microbenchmark, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_stats.py tests/test_gpu_masker_ops.py tests/test_gpu_ops.py -q -m gpu --tb=short -x > gpurun_out/g6_unit.log 2>&1; tail -5 gpurun_out/g6_unit.log | cut -c1-250
timeout 600 python scripts/bench_hbm_kernels.py --json gpurun_out/g6_hbm_flushed.json > gpurun_out/g6_hbm_flushed.txt 2>&1; cat gpurun_out/g6_hbm_flushed.txt
timeout 900 python bench.py --steps 8 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g6_bench_full.json 2> gpurun_out/g6_bench_full.err; tail -c 800 gpurun_out/g6_bench_full.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/g6_bench_full.json").read().strip().splitlines()[-1])
    print("full:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "launches/step", d["gpu_launches_per_step"], "conv", d["roofline"]["conv_aggregate"])
except Exception as e:
    print("bench parse failed", e)
PY
