#!/bin/bash
# round 2, call 46/51: whole GPU suite (twice: flakiness) + smoke + every bench workload on the final tree, 1 GPU
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g51_pytest.log 2>&1; tail -4 gpurun_out/g51_pytest.log | cut -c1-200
timeout 1800 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g51_pytest_second_run.log 2>&1; tail -2 gpurun_out/g51_pytest_second_run.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g51_smoke.log 2>&1; tail -1 gpurun_out/g51_smoke.log | cut -c1-200
timeout 1200 python bench.py --steps 10 --warmup 3 --topk 1000 > gpurun_out/g51_bench_full.json 2> gpurun_out/g51_bench_full.err
timeout 600 python bench.py --workload painter --steps 8 --warmup 3 --topk 1000 --no-cpu-baseline > gpurun_out/g51_bench_painter.json 2> gpurun_out/g51_bench_painter.err
timeout 600 python bench.py --workload masker --steps 6 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g51_bench_masker.json 2> gpurun_out/g51_bench_masker.err
timeout 600 python bench.py --workload infer --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g51_bench_infer.json 2> gpurun_out/g51_bench_infer.err
timeout 600 python bench.py --workload infer --dtype fp16 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/g51_bench_infer_fp16.json 2> gpurun_out/g51_bench_infer_fp16.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/g51_bench_reference_arm.json 2> gpurun_out/g51_bench_reference_arm.err; tail -c 600 gpurun_out/g51_bench_reference_arm.json
timeout 600 python scripts/profile_full_step.py > gpurun_out/g51_profile_full.txt 2>&1
python - <<'PY'
import json
for name in ("full", "painter", "masker", "infer", "infer_fp16"):
    try:
        d = json.loads(open(f"gpurun_out/g51_bench_{name}.json").read().strip().splitlines()[-1])
        print(name, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; e2e", round(d["e2e"]["value"], 2) if d.get("e2e") else None,
              "launches/step", d.get("gpu_launches_per_step"), "step_frac", round(d["roofline"].get("step_frac", 0), 4), "conv", d["roofline"].get("conv_aggregate", {}).get("frac"),
              "eager", (d.get("gpu_eager_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(name, "parse failed", e)
PY
