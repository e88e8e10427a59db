#!/bin/bash
# Round-end measurement call: GPU tests, the default bench line (full step, with cpu_baseline + e2e), the reference arm,
# the painter / inference workloads and the single-kernel timings quoted in DESIGN.md.
#   gpurun --timeout 900 -- 'bash scripts/gpu_final.sh'
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
REPS=20 timeout 120 python scripts/bench_conv.py r1w r3w wgsh wg48 r1 r3 sh8 gb48_8 > gpurun_out/bench_conv_final.log 2>&1; cat gpurun_out/bench_conv_final.log
timeout 300 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; head -c 400 gpurun_out/bench_full.json; echo
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_full_ref.json 2>> gpurun_out/bench_full.err; head -c 200 gpurun_out/bench_full_ref.json; echo
timeout 200 python bench.py --workload painter --no-cpu-baseline > gpurun_out/bench_painter.json 2> gpurun_out/bench_painter.err; head -c 200 gpurun_out/bench_painter.json; echo
timeout 200 python bench.py --workload infer --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err; head -c 200 gpurun_out/bench_infer.json; echo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
