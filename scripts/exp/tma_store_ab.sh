#!/bin/bash
# A/B of the TMA-store copy-out of the streaming conv kernel (CGB_TMA_STORE=1): parity tests with the flag on, then
# single-kernel timings and the full-step bench with the flag off / on.  usage: gpurun -- 'bash scripts/exp/tma_store_ab.sh'
mkdir -p gpurun_out
CGB_TMA_STORE=1 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_masker_ops.py tests/test_gpu_painter.py \
    tests/test_gpu_masker.py tests/test_gpu_discriminator.py -q -m gpu --tb=short -x > gpurun_out/tma_store_pytest.log 2>&1
echo "pytest(tma) rc=$?" >> gpurun_out/tma_store_pytest.log; tail -4 gpurun_out/tma_store_pytest.log
CASES="sh8 r1 r1b r3 r3d r2 r4 vgg3 vgg3d vgg2 aspp sn640"
REPS=20 CGB_TMA_STORE=0 timeout 120 python scripts/bench_conv.py $CASES > gpurun_out/tma_store_conv_off.log 2>&1
REPS=20 CGB_TMA_STORE=1 timeout 120 python scripts/bench_conv.py $CASES > gpurun_out/tma_store_conv_on.log 2>&1
paste -d'|' gpurun_out/tma_store_conv_off.log gpurun_out/tma_store_conv_on.log | cut -c1-230
CGB_TMA_STORE=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/tma_store_bench_off.json 2> gpurun_out/tma_store_bench_off.err
CGB_TMA_STORE=1 timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/tma_store_bench_on.json 2> gpurun_out/tma_store_bench_on.err
for f in off on; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/tma_store_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"])
except Exception as e:
    print("$f", "failed", e)
PY
done
