#!/bin/bash
# round 2, call 65: round-end measurements on the final tree
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g65_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g65_pytest_gpu.log; tail -3 gpurun_out/g65_pytest_gpu.log
REPS=20 timeout 200 python scripts/bench_conv.py r1w r3w wgsh wg48 r1 r3 sh8 gb48_8 gb48 gb80_640 dg48 dg80 vgg1d sn24 2>&1 | grep -v Warn > gpurun_out/g65_bench_conv.log; cat gpurun_out/g65_bench_conv.log
timeout 400 python bench.py > gpurun_out/g65_bench_full.json 2> gpurun_out/g65_bench_full.err; head -c 300 gpurun_out/g65_bench_full.json; echo
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/g65_bench_ref.json 2>> gpurun_out/g65_bench_full.err; head -c 300 gpurun_out/g65_bench_ref.json; echo
timeout 300 python bench.py --workload painter > gpurun_out/g65_bench_painter.json 2> gpurun_out/g65_bench_painter.err; head -c 200 gpurun_out/g65_bench_painter.json; echo
timeout 300 python bench.py --workload masker --no-cpu-baseline > gpurun_out/g65_bench_masker.json 2> gpurun_out/g65_bench_masker.err; head -c 200 gpurun_out/g65_bench_masker.json; echo
timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/g65_bench_infer.json 2> gpurun_out/g65_bench_infer.err; head -c 200 gpurun_out/g65_bench_infer.json; echo
timeout 300 python bench.py --workload infer --dtype fp16 --no-cpu-baseline --no-gpu-eager > gpurun_out/g65_bench_infer_fp16.json 2> gpurun_out/g65_bench_infer_fp16.err; head -c 200 gpurun_out/g65_bench_infer_fp16.json; echo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/g65_smoke.log 2>&1; tail -2 gpurun_out/g65_smoke.log | cut -c1-300
