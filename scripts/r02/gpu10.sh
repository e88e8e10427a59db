#!/bin/bash
mkdir -p gpurun_out
CGB_TC2=0 timeout 180 python scripts/exp/tc2_check.py save > gpurun_out/g10_tc2_ref.txt 2>&1; echo "save rc=$?"
CGB_TC2=2 timeout 180 python scripts/exp/tc2_check.py check > gpurun_out/g10_tc2_check.txt 2>&1; echo "check rc=$?"
paste <(grep " us" gpurun_out/g10_tc2_ref.txt) <(grep " us" gpurun_out/g10_tc2_check.txt | awk '{print $(NF-1), $NF}')
grep -E "OK|FAIL" gpurun_out/g10_tc2_check.txt | tr '\n' ';'
