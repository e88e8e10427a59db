#!/bin/bash
# round 2, call 9: first run of the cta_group::2 streaming kernel (own timeouts: a hung cluster kernel must not eat the call)
mkdir -p gpurun_out
CGB_TC2=0 timeout 120 python scripts/exp/tc2_check.py save > gpurun_out/g9_tc2_ref.txt 2>&1; echo "save rc=$?"; cat gpurun_out/g9_tc2_ref.txt | tail -12
CGB_TC2=1 timeout 120 python scripts/exp/tc2_check.py check > gpurun_out/g9_tc2_check.txt 2>&1; echo "check rc=$?"; cat gpurun_out/g9_tc2_check.txt | tail -24
nvidia-smi --query-gpu=name,memory.used --format=csv
