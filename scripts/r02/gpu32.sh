#!/bin/bash
# round 2, call 32: bias tap (mlp_shared bias gradient out of the wgrad), two staging tiles in the weight-stationary kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_fused_stats.py tests/test_gpu_painter.py tests/test_gpu_masker.py tests/test_gpu_infer_all.py -q -m gpu --tb=short -x > gpurun_out/g32_unit.log 2>&1; tail -3 gpurun_out/g32_unit.log | cut -c1-300
for v in 0 1; do echo "CGB_WS_STAGING2=$v"; CGB_WS_STAGING2=$v REPS=10 timeout 300 python scripts/bench_conv.py sn24 gb48 dg48 2>&1 | tail -3; done
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g32_bench_full.json 2> gpurun_out/g32_bench_full.err; tail -c 200 gpurun_out/g32_bench_full.err
timeout 600 python bench.py --workload painter --steps 8 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g32_bench_painter.json 2> gpurun_out/g32_bench_painter.err
timeout 600 python bench.py --workload infer --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g32_bench_infer.json 2> gpurun_out/g32_bench_infer.err
python - <<'PY'
import json
for name in ("full", "painter", "infer"):
    try:
        d = json.loads(open(f"gpurun_out/g32_bench_{name}.json").read().strip().splitlines()[-1])
        print(name, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None,
              "launches/step", d.get("gpu_launches_per_step"), "step_frac", d["roofline"].get("step_frac"), "conv", d["roofline"].get("conv_aggregate"))
    except Exception as e:
        print(name, "parse failed", e)
PY
