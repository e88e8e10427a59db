#!/bin/bash
# One gpurun call of the round: GPU tests, the three bench workloads + the CPU reference arm, single-kernel timings, and
# `ncu --set full` captures of the dominant kernels (summarise with scripts/ncu_summary.py into profiles/).
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh'            (add --gpus 2 and run scripts/gpu_round.sh ddp for N=2)
mkdir -p gpurun_out
if [ "$1" = "ddp" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_full_2gpu.json 2> gpurun_out/bench_full_2gpu.err
  head -c 400 gpurun_out/bench_full_2gpu.json; echo; exit 0
fi
timeout 1500 python -m pytest tests -q -m gpu --tb=short > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; head -c 300 gpurun_out/bench_full.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_full_ref.json 2>> gpurun_out/bench_full.err
timeout 600 python bench.py --workload painter > gpurun_out/bench_painter.json 2> gpurun_out/bench_painter.err
timeout 600 python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
REPS=20 timeout 300 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1
python scripts/profile_full_step.py > gpurun_out/profile_step.log 2>&1
for c in r3 r1 r1w sh8 gb48_8; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc' -s 1 -c 2 -o gpurun_out/prof_$c \
      python scripts/bench_conv.py $c > gpurun_out/ncu_$c.log 2>&1
done
# launch list of the default bench command (slow: ~14k launches per step under ncu; bounded by the timeout)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_full.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_full_under_ncu.log 2>&1
ls -la gpurun_out
