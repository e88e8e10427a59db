#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full_2gpu.json 2> gpurun_out/bench_full_2gpu.err
echo "rc=$?"; head -c 400 gpurun_out/bench_full_2gpu.json; echo; tail -5 gpurun_out/bench_full_2gpu.err | cut -c1-300
