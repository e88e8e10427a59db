#!/bin/bash
# round 2, call 8: dual packing / memset-free statistics (tests + bench), resident-weight 1x1 A/B, infer bench bf16 vs fp16, painter bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_new_kernels.py tests/test_gpu_fused_stats.py tests/test_gpu_full_step.py -q -m gpu --tb=short -x > gpurun_out/g8_unit.log 2>&1; tail -5 gpurun_out/g8_unit.log | cut -c1-300
for v in 0 1; do echo "CGB_WS_1X1=$v"; CGB_WS_1X1=$v REPS=20 timeout 300 python scripts/bench_conv.py r1 r1b r3 2>&1 | tail -4; done
timeout 900 python bench.py --steps 8 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g8_bench_full.json 2> gpurun_out/g8_bench_full.err; tail -c 600 gpurun_out/g8_bench_full.err
CGB_WS_1X1=1 timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g8_bench_full_ws1x1.json 2> gpurun_out/g8_bench_full_ws1x1.err
timeout 600 python bench.py --workload infer --steps 8 --warmup 3 > gpurun_out/g8_bench_infer_bf16.json 2> gpurun_out/g8_bench_infer_bf16.err
timeout 600 python bench.py --workload infer --dtype fp16 --steps 8 --warmup 3 > gpurun_out/g8_bench_infer_fp16.json 2> gpurun_out/g8_bench_infer_fp16.err; tail -c 400 gpurun_out/g8_bench_infer_fp16.err
timeout 600 python bench.py --workload painter --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g8_bench_painter.json 2> gpurun_out/g8_bench_painter.err; tail -c 400 gpurun_out/g8_bench_painter.err
python - <<'PY'
import json
for name in ("full", "full_ws1x1", "infer_bf16", "infer_fp16", "painter"):
    try:
        d = json.loads(open(f"gpurun_out/g8_bench_{name}.json").read().strip().splitlines()[-1])
        print(name, round(d["value"], 2), "img/s", round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None,
              "launches/step", d.get("gpu_launches_per_step"), "eager", d.get("gpu_eager_baseline", {}) and d["gpu_eager_baseline"].get("value"))
    except Exception as e:
        print(name, "parse failed", e)
PY
