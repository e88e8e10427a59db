#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench rc=$?"; head -c 330 gpurun_out/bench_full.json; echo; tail -2 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_full_ref.json 2>> gpurun_out/bench_full.err; head -c 300 gpurun_out/bench_full_ref.json; echo
timeout 600 python bench.py --workload painter > gpurun_out/bench_painter.json 2> gpurun_out/bench_painter.err; head -c 300 gpurun_out/bench_painter.json; echo
REPS=20 timeout 300 python scripts/bench_conv.py r3 r3d r3w r1 r1w r1b stem aspp vgg3 gb48 gb80 gb160 dg48 wg48 > gpurun_out/bench_conv.log 2>&1; cat gpurun_out/bench_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 2 -o gpurun_out/prof_r3 python scripts/bench_conv.py r3 > gpurun_out/ncu_r3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 2 -o gpurun_out/prof_r1 python scripts/bench_conv.py r1 > gpurun_out/ncu_r1.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log | cut -c1-600
ls -la gpurun_out/*.ncu-rep
