#!/bin/bash
# round 2, call 66 (2 GPUs): data-parallel correctness and the 2-GPU weak-scaling bench on the final tree (new epilogue)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/g66_dp_check.json 2> gpurun_out/g66_dp_check.err; tail -2 gpurun_out/g66_dp_check.json | cut -c1-900
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g66_bench_2gpu.json 2> gpurun_out/g66_bench_2gpu.err; tail -1 gpurun_out/g66_bench_2gpu.json | cut -c1-400
