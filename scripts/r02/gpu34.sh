#!/bin/bash
# round 2, call 34: halo wgrad with an all-taps-resident N = 48 tile (CGB_WG_HALO_BN=48) vs the default forms
mkdir -p gpurun_out
ONLY=w_r3,w_r4,w_aspp,w_128 timeout 300 python scripts/exp/tc2_check.py save 2>&1 | tail -4
ONLY=w_r3,w_r4,w_aspp,w_128 CGB_WG_HALO_BN=48 timeout 300 python scripts/exp/tc2_check.py check 2>&1 | tail -8
ONLY=w_r3,w_r4,w_aspp,w_128 CGB_WG_HALO_BN=32 timeout 300 python scripts/exp/tc2_check.py check 2>&1 | tail -8
