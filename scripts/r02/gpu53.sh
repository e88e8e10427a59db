#!/bin/bash
# round 2, call 53 (8 GPUs): weak-scaling bench at N = 8 on the final tree (the driver's scaling run does 1/2/4/8)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g53_bench_8gpu.json 2> gpurun_out/g53_bench_8gpu.err; tail -1 gpurun_out/g53_bench_8gpu.json | cut -c1-300; tail -3 gpurun_out/g53_bench_8gpu.err | cut -c1-300
