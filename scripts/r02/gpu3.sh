#!/bin/bash
# round 2, call 3: epilogue statistics + flat BatchNorm passes + parity at the benchmarked shapes, then the whole suite and the bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_stats.py tests/test_gpu_graphs.py -q -m gpu --tb=short > gpurun_out/g3_new.log 2>&1; tail -25 gpurun_out/g3_new.log
timeout 900 python -m pytest tests/test_gpu_parity_at_size.py -q -m gpu --tb=short > gpurun_out/g3_at_size.log 2>&1; tail -25 gpurun_out/g3_at_size.log
timeout 900 python -m pytest tests -q -m gpu --tb=short -x --deselect tests/test_gpu_graphs.py --deselect tests/test_gpu_fused_stats.py --deselect tests/test_gpu_parity_at_size.py > gpurun_out/g3_pytest.log 2>&1; tail -6 gpurun_out/g3_pytest.log
for v in 1 0; do
CGB_EPILOGUE_STATS=$v timeout 900 python bench.py --steps 8 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g3_bench_full_stats$v.json 2> gpurun_out/g3_bench_full_stats$v.err; tail -c 1500 gpurun_out/g3_bench_full_stats$v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/g3_bench_full_stats$v.json").read().strip().splitlines()[-1])
    print("full CGB_EPILOGUE_STATS=$v:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "launches/step", d["gpu_launches_per_step"], "conv", d["roofline"]["conv_aggregate"])
except Exception as e:
    print("bench parse failed", e)
PY
done
