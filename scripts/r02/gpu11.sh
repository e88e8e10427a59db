#!/bin/bash
mkdir -p gpurun_out
CGB_TC2=0 CGB_TC2_WGRAD=0 timeout 180 python scripts/exp/tc2_check.py save > gpurun_out/g11_ref.txt 2>&1; echo "save rc=$?"
CGB_TC2=1 CGB_TC2_WGRAD=1 timeout 180 python scripts/exp/tc2_check.py check > gpurun_out/g11_check.txt 2>&1; echo "check rc=$?"
paste <(grep " us" gpurun_out/g11_ref.txt) <(grep " us" gpurun_out/g11_check.txt | awk '{print $(NF-1), $NF}')
grep -E "FAIL" gpurun_out/g11_check.txt; grep -c OK gpurun_out/g11_check.txt; tail -3 gpurun_out/g11_check.txt
