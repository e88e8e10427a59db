#!/bin/bash
# round 2, call 36/52 (2 GPUs): data-parallel correctness (scripts/dp_check.py), the multi-GPU test, 2-GPU weak-scaling bench + reference arm semantics
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/g52_dp_check.json 2> gpurun_out/g52_dp_check.err; tail -2 gpurun_out/g52_dp_check.json | cut -c1-900
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --tb=short > gpurun_out/g52_multi.log 2>&1; tail -2 gpurun_out/g52_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/g52_bench_2gpu.json 2> gpurun_out/g52_bench_2gpu.err; tail -1 gpurun_out/g52_bench_2gpu.json | cut -c1-400
