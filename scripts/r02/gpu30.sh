#!/bin/bash
# round 2, call 30: AdvEnt D first conv via im2col; masker/advent tests; full / painter / infer benches on the new epilogue
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_discriminator.py tests/test_gpu_full_step.py tests/test_gpu_masker.py tests/test_first_layer_im2col.py -q -m gpu --tb=short -x > gpurun_out/g30_unit.log 2>&1; tail -3 gpurun_out/g30_unit.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g30_bench_full.json 2> gpurun_out/g30_bench_full.err; tail -c 200 gpurun_out/g30_bench_full.err
timeout 600 python bench.py --workload painter --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g30_bench_painter.json 2> gpurun_out/g30_bench_painter.err
timeout 600 python bench.py --workload infer --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g30_bench_infer.json 2> gpurun_out/g30_bench_infer.err
timeout 600 python bench.py --workload masker --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g30_bench_masker.json 2> gpurun_out/g30_bench_masker.err; tail -c 300 gpurun_out/g30_bench_masker.err
python - <<'PY'
import json
for name in ("full", "painter", "infer", "masker"):
    try:
        d = json.loads(open(f"gpurun_out/g30_bench_{name}.json").read().strip().splitlines()[-1])
        print(name, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None,
              "launches/step", d.get("gpu_launches_per_step"), "step_frac", d["roofline"].get("step_frac"), "conv", d["roofline"].get("conv_aggregate"))
    except Exception as e:
        print(name, "parse failed", e)
PY
