#!/bin/bash
# round 2, call 2: CUDA-graph step vs eager (tests), the whole GPU suite on the new tree, the new bench line (graphs, GPU-eager bar)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graphs.py -q -m gpu --tb=short -x > gpurun_out/g2_graphs.log 2>&1; tail -15 gpurun_out/g2_graphs.log
CGB_RUN_SWEEP=1 timeout 600 python -m pytest tests -q -m gpu --tb=short -x --deselect tests/test_gpu_graphs.py > gpurun_out/g2_pytest.log 2>&1; tail -4 gpurun_out/g2_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --topk 1000 > gpurun_out/g2_bench_full.json 2> gpurun_out/g2_bench_full.err; tail -c 1500 gpurun_out/g2_bench_full.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/g2_bench_full.json").read().strip().splitlines()[-1])
    print("full graphs:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "e2e", d["e2e"], "launches/step", d["gpu_launches_per_step"])
    print("gpu eager:", json.dumps(d["gpu_eager_baseline"])[:1200])
    print("cpu:", d["cpu_baseline"])
    print("roofline:", {k: v for k, v in d["roofline"].items() if k not in ("top_classes",)})
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --workload masker --steps 4 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g2_bench_masker.json 2> gpurun_out/g2_bench_masker.err; tail -c 800 gpurun_out/g2_bench_masker.err; head -c 700 gpurun_out/g2_bench_masker.json
