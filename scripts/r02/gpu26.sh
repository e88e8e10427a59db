#!/bin/bash
# round 2, call 26: 8 epilogue warps + x32 accumulator loads + split variants: unit tests, single kernels, full-step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_at_size.py tests/test_gpu_fused_stats.py tests/test_gpu_half.py -q -m gpu --tb=short -x > gpurun_out/g26_unit.log 2>&1; tail -3 gpurun_out/g26_unit.log | cut -c1-300
NOBIAS=1 REPS=20 timeout 300 python scripts/bench_conv.py r1 r1b r3 r3d 2>&1 | tail -4
REPS=20 timeout 300 python scripts/bench_conv.py sh8 gb48_8 dg48 aspp vgg3 vgg3d r4 d3 2>&1 | tail -8
ONLY=r1,r1b,stats,l4,gb160 CGB_TC2=0 timeout 300 python scripts/exp/tc2_check.py save 2>&1 | tail -5
ONLY=r1,r1b,stats,l4,gb160 CGB_TC2=2 timeout 300 python scripts/exp/tc2_check.py check 2>&1 | tail -10
timeout 900 python bench.py --steps 10 --warmup 3 --topk 1000 --no-cpu-baseline --no-gpu-eager > gpurun_out/g26_bench_full.json 2> gpurun_out/g26_bench_full.err; tail -c 300 gpurun_out/g26_bench_full.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/g26_bench_full.json").read().strip().splitlines()[-1])
print("full:", d["value"], "img/s", d["ms_per_step"], "ms/step; eager", d["eager_ms_per_step"], "e2e", d["e2e"]["value"], "launches/step", d["gpu_launches_per_step"])
print("conv", d["roofline"]["conv_aggregate"], "step_frac", d["roofline"]["step_frac"])
PY
