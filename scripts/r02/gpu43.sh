#!/bin/bash
# round 2, call 43: BatchNorm backward summing its two consumers' gradients (cgb_bn_train_bwd2): tests + A/B bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bn_dual.py tests/test_conv_skip.py tests/test_gpu_masker.py tests/test_gpu_full_step.py tests/test_gpu_graphs.py tests/test_gpu_full_size.py tests/test_gpu_masker_ops.py -q -m gpu --tb=short -x > gpurun_out/g43_unit.log 2>&1; tail -3 gpurun_out/g43_unit.log | cut -c1-300
for v in 0 1; do
CGB_BN_DUAL=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g43_bench_full_dual$v.json 2> gpurun_out/g43_bench_full_dual$v.err
done
CGB_BN_DUAL=1 timeout 600 python bench.py --workload masker --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g43_bench_masker.json 2> gpurun_out/g43_bench_masker.err
python - <<'PY'
import json
for v in ("full_dual0", "full_dual1", "masker"):
    d = json.loads(open(f"gpurun_out/g43_bench_{v}.json").read().strip().splitlines()[-1])
    print(v, round(d["value"], 2), d["unit"], round(d["ms_per_step"], 2), "ms; launches/step", d.get("gpu_launches_per_step"))
PY
