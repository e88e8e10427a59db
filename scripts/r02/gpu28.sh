#!/bin/bash
# round 2, call 28: batched mask loads in the per-thread copy-out; unit tests, masked dgrads, full-step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_half.py tests/test_gpu_painter.py -q -m gpu --tb=short -x > gpurun_out/g28_unit.log 2>&1; tail -3 gpurun_out/g28_unit.log | cut -c1-300
REPS=10 timeout 300 python scripts/bench_conv.py dg48 r3d vgg3d 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g28_bench_full.json 2> gpurun_out/g28_bench_full.err; tail -c 300 gpurun_out/g28_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/g28_bench_full.json").read().strip().splitlines()[-1])
print("full:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
print("conv", d["roofline"]["conv_aggregate"], "step_frac", d["roofline"]["step_frac"])
PY
