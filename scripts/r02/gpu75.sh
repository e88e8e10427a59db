#!/bin/bash
# round 2, call 75: round-end measurements on the final tree (after the grid-size / max-pool / bias-gradient changes)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g75_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g75_pytest_gpu.log; tail -3 gpurun_out/g75_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/g75_bench_full.json 2> gpurun_out/g75_bench_full.err; head -c 300 gpurun_out/g75_bench_full.json; echo
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/g75_bench_ref.json 2>> gpurun_out/g75_bench_full.err; head -c 200 gpurun_out/g75_bench_ref.json; echo
timeout 300 python bench.py --workload painter > gpurun_out/g75_bench_painter.json 2> gpurun_out/g75_bench_painter.err; head -c 200 gpurun_out/g75_bench_painter.json; echo
timeout 300 python bench.py --workload masker --no-cpu-baseline > gpurun_out/g75_bench_masker.json 2> gpurun_out/g75_bench_masker.err; head -c 200 gpurun_out/g75_bench_masker.json; echo
timeout 300 python bench.py --workload infer --no-cpu-baseline > gpurun_out/g75_bench_infer.json 2> gpurun_out/g75_bench_infer.err; head -c 200 gpurun_out/g75_bench_infer.json; echo
timeout 300 python bench.py --workload infer --dtype fp16 --no-cpu-baseline --no-gpu-eager > gpurun_out/g75_bench_infer_fp16.json 2> gpurun_out/g75_bench_infer_fp16.err; head -c 200 gpurun_out/g75_bench_infer_fp16.json; echo
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/g75_smoke.log 2>&1; tail -1 gpurun_out/g75_smoke.log | cut -c1-100
timeout 600 python scripts/profile_full_step.py > gpurun_out/g75_profile_full.txt 2>&1; head -8 gpurun_out/g75_profile_full.txt | cut -c1-160
timeout 300 python scripts/bench_hbm_kernels.py > gpurun_out/g75_hbm_kernels_flushed.txt 2>&1; tail -3 gpurun_out/g75_hbm_kernels_flushed.txt
