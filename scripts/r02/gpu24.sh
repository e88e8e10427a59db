#!/bin/bash
# round 2, call 24: pair kernel with compile-time epilogue variants: one-CTA (CGB_TC2=0) vs pair on every shape (CGB_TC2=2), bitwise check
mkdir -p gpurun_out
ONLY=r1,r1b,r3,r3d,odd80,s2,aspp,l4,stats,vgg3,vgg3d,vgg4,r4,d3,gb160,l4b,r1d
ONLY=$ONLY CGB_TC2=0 timeout 300 python scripts/exp/tc2_check.py save > gpurun_out/g24_ref.txt 2>&1; cat gpurun_out/g24_ref.txt | tail -20
ONLY=$ONLY CGB_TC2=2 timeout 300 python scripts/exp/tc2_check.py check > gpurun_out/g24_check.txt 2>&1; cat gpurun_out/g24_check.txt | tail -40
