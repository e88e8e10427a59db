#!/bin/bash
# round 2, call 7 (2 GPUs): the whole GPU suite on the current tree, the data-parallel correctness check, the 2-GPU bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g7_smi.txt
timeout 1200 python -m pytest tests -q -m gpu --tb=short > gpurun_out/g7_pytest.log 2>&1; tail -12 gpurun_out/g7_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_check.py > gpurun_out/g7_dp_check.json 2> gpurun_out/g7_dp_check.err; tail -c 1200 gpurun_out/g7_dp_check.err; cat gpurun_out/g7_dp_check.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/g7_bench_2gpu.json 2> gpurun_out/g7_bench_2gpu.err; tail -c 600 gpurun_out/g7_bench_2gpu.err; head -c 500 gpurun_out/g7_bench_2gpu.json
