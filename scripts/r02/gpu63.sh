#!/bin/bash
# round 2, call 63: evidence for the operand-by-TMA epilogue: ncu --set full of the 48->128 masked dgrad, memcheck over the conv unit
# tests, the whole GPU suite
mkdir -p gpurun_out
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_ws_kernel -s 1 -c 1 -o gpurun_out/g63_dg48 -f python scripts/bench_conv.py dg48 > gpurun_out/g63_ncu.log 2>&1; tail -2 gpurun_out/g63_ncu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_ops.py -q -m gpu --tb=line -x -k "conv" > gpurun_out/g63_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/g63_memcheck.log | cut -c1-200
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g63_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g63_pytest_gpu.log; tail -4 gpurun_out/g63_pytest_gpu.log
