#!/bin/bash
# round 2, call 14: fresh torch.profiler breakdown of the full step on the current tree (all rows)
mkdir -p gpurun_out
timeout 600 python scripts/profile_full_step.py > gpurun_out/g14_profile.txt 2>&1; head -5 gpurun_out/g14_profile.txt | cut -c1-200
