#!/bin/bash
# round 2, call 29: first-layer convs as im2col + GEMM (discriminator model0, VGG conv1_1): tests + full-step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_first_layer_im2col.py tests/test_gpu_discriminator.py tests/test_gpu_trainer.py tests/test_gpu_full_step.py tests/test_gpu_graphs.py -q -m gpu --tb=short -x > gpurun_out/g29_unit.log 2>&1; tail -3 gpurun_out/g29_unit.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g29_bench_full.json 2> gpurun_out/g29_bench_full.err; tail -c 300 gpurun_out/g29_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/g29_bench_full.json").read().strip().splitlines()[-1])
print("full:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
print("conv", d["roofline"]["conv_aggregate"], "step_frac", d["roofline"]["step_frac"])
PY
