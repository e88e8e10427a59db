#!/bin/bash
# round 2, first GPU call: baseline of the round-1 tree with the one-launch weight packing A/B, the option sweep on hardware,
# and the FULL per-class conv table (CGB_TOPK=1000) of the three workloads for offline roofline analysis.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/g1_smi.txt
CGB_RUN_SWEEP=1 CGB_PACK_KERNEL=1 timeout 600 python -m pytest tests -q -m gpu --tb=short -x > gpurun_out/g1_pytest.log 2>&1; tail -4 gpurun_out/g1_pytest.log
for v in 0 1; do
  CGB_TOPK=1000 CGB_PACK_KERNEL=$v timeout 300 python bench.py --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/g1_bench_full_pack$v.json 2> gpurun_out/g1_bench_full_pack$v.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/g1_bench_full_pack$v.json").read().strip().splitlines()[-1])
print("full CGB_PACK_KERNEL=$v", d["value"], "img/s", d["ms_per_step"], "ms/step, e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
PY
done
CGB_TOPK=1000 CGB_PACK_KERNEL=1 timeout 300 python bench.py --no-cpu-baseline --workload painter --steps 5 --warmup 3 > gpurun_out/g1_bench_painter.json 2> gpurun_out/g1_bench_painter.err; tail -c 600 gpurun_out/g1_bench_painter.json
CGB_TOPK=1000 CGB_PACK_KERNEL=1 timeout 300 python bench.py --no-cpu-baseline --workload infer --steps 5 --warmup 3 > gpurun_out/g1_bench_infer.json 2> gpurun_out/g1_bench_infer.err; tail -c 600 gpurun_out/g1_bench_infer.json
