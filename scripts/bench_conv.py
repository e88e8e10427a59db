"""Micro-benchmark of single conv launches (CUDA events), for ncu captures and tuning.
usage: python scripts/bench_conv.py [case ...]   cases: gb48 gb80 sh dg48 wg48 sn24"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climategan_b200 import _lib, ops
dev = torch.device("cuda:0")
CASES = {
    # name: (which, n, ci, co, h, w, k, pad)
    "gb48": ("fwd", 16, 128, 48, 640, 640, 3, 1),
    "gb80": ("fwd", 16, 128, 80, 320, 320, 3, 1),
    "gb160": ("fwd", 16, 128, 160, 160, 160, 3, 1),
    "sh": ("fwd", 16, 32, 128, 640, 640, 1, 0),
    "sn24": ("fwd", 16, 24, 24, 640, 640, 3, 1),
    "sn640": ("fwd", 16, 640, 640, 20, 20, 3, 1),
    "dg48": ("dgrad", 16, 128, 48, 640, 640, 3, 1),
    "wg48": ("wgrad", 16, 128, 48, 640, 640, 3, 1),
    "wgsh": ("wgrad", 16, 32, 128, 640, 640, 1, 0),
}
names = sys.argv[1:] or list(CASES)
reps = int(os.environ.get("REPS", "5"))
for name in names:
    which, n, ci, co, h, w, k, pad = CASES[name]
    x = torch.randn(n, h, w, ci, device=dev).bfloat16()
    wp = (torch.randn(co, k * k, ci, device=dev) * 0.05).bfloat16()
    bias = torch.zeros(co, device=dev)
    g = ops.ConvGeom(k, k, 1, 1, pad, _lib.PAD_ZERO, _lib.ACT_NONE, 0.2, _lib.ENGINE_AUTO)
    gy = torch.randn(n, h, w, co, device=dev).bfloat16()
    def run():
        if which == "fwd":
            return ops.conv_fwd_raw(x, wp, bias, None, g)
        if which == "dgrad":
            return ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g, _lib.ACT_RELU, x)
        return ops.conv_wgrad_raw(x, gy, g, False)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * n * h * w * ci * co * k * k
    byts = 2.0 * n * h * w * (ci + co)
    print(f"{name:6s} {which:5s} {ci}->{co} k{k} @{h}x{w} n={n}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.0f} GB/s (min HBM traffic)", flush=True)
