mkdir -p gpurun_out
for v in 0 1; do
echo "CGB_WS_1X1=$v"
CGB_WS_1X1=$v REPS=30 timeout 300 python scripts/bench_conv.py r1 r1b sh8 r3 2>&1 | grep -v Warn
CGB_WS_1X1=$v timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "conv_fwd_bwd" 2>&1 | tail -2
done
CGB_WS_1X1=1 timeout 900 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | head -c 230; echo
