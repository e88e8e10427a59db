#!/bin/bash
mkdir -p gpurun_out
for cfg in wcache viewgrad aliasparam "wcache,viewgrad,aliasparam"; do
CGB_DEBUG_DISABLE=$cfg timeout 600 python -m pytest tests/test_gpu_full_step.py -q -m gpu --tb=short -k fp32 > gpurun_out/iso_$cfg.log 2>&1
echo "disable=$cfg"; grep -E "^E  +Assert|passed|failed" gpurun_out/iso_$cfg.log | cut -c1-300
done
