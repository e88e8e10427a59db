#!/bin/bash
# round 2, call 4: register-persistent epilogue statistics; profiler view of the step (where do the BatchNorm passes stand?)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_stats.py tests/test_gpu_graphs.py -q -m gpu --tb=line > gpurun_out/g4_new.log 2>&1; tail -8 gpurun_out/g4_new.log
timeout 600 python scripts/profile_full_step.py > gpurun_out/g4_profile_stats1.txt 2>&1; head -40 gpurun_out/g4_profile_stats1.txt | cut -c1-150
for v in 1 0; do
CGB_EPILOGUE_STATS=$v timeout 900 python bench.py --steps 8 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g4_bench_full_stats$v.json 2> gpurun_out/g4_bench_full_stats$v.err; tail -c 1500 gpurun_out/g4_bench_full_stats$v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/g4_bench_full_stats$v.json").read().strip().splitlines()[-1])
    print("full CGB_EPILOGUE_STATS=$v:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "launches/step", d["gpu_launches_per_step"], "conv", d["roofline"]["conv_aggregate"])
except Exception as e:
    print("bench parse failed", e)
PY
done
