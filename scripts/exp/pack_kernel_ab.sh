#!/bin/bash
# First GPU call of the next round: A/B of the one-launch weight packing (CGB_PACK_KERNEL=1, written after round 1's GPU budget
# was spent).  Bit-exactness first (tests/test_gpu_zz_new_kernels.py::test_pack_weight_kernel), then the parity suite with the flag on,
# then the full-step bench off / on.   usage: gpurun --timeout 900 -- 'bash scripts/exp/pack_kernel_ab.sh'
# The unit log also covers the other entry points written after the budget was spent (reverse-Huber depth loss, argmax-confusion
# validation metrics, diff-augment): tests/test_gpu_zz_new_kernels.py holds all of them.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_zz_new_kernels.py -q -m gpu --tb=short > gpurun_out/pack_kernel_unit.log 2>&1; tail -3 gpurun_out/pack_kernel_unit.log
CGB_RUN_SWEEP=1 CGB_PACK_KERNEL=1 timeout 420 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pack_kernel_pytest.log 2>&1; tail -4 gpurun_out/pack_kernel_pytest.log
for v in 0 1; do
  CGB_PACK_KERNEL=$v timeout 200 python bench.py --no-cpu-baseline > gpurun_out/pack_kernel_bench_$v.json 2> gpurun_out/pack_kernel_bench_$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/pack_kernel_bench_$v.json").read().strip().splitlines()[-1])
print("CGB_PACK_KERNEL=$v", d["value"], "img/s", d["ms_per_step"], "ms/step, e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
PY
done
