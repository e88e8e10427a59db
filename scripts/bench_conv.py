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
    # masker (full train step, 8 images per domain): (which, n, ci, co, h, w, k, pad, dil, stride)
    "r3": ("fwd", 8, 256, 256, 80, 80, 3, 2, 2, 1),       # ResNet layer3 conv2, dilation 2 — the step's dominant class
    "r3d": ("dgrad", 8, 256, 256, 80, 80, 3, 2, 2, 1),
    "r3w": ("wgrad", 8, 256, 256, 80, 80, 3, 2, 2, 1),
    "r1": ("fwd", 8, 256, 1024, 80, 80, 1, 0, 1, 1),      # layer3 conv3
    "r1w": ("wgrad", 8, 256, 1024, 80, 80, 1, 0, 1, 1),
    "r1b": ("fwd", 8, 1024, 256, 80, 80, 1, 0, 1, 1),     # layer3 conv1
    "stem": ("fwd", 8, 8, 64, 640, 640, 7, 3, 1, 2),      # conv1 7x7 s2 on the 3-channel image
    "aspp": ("fwd", 8, 2048, 256, 80, 80, 3, 12, 12, 1),  # ASPP atrous branch
    "vgg3": ("fwd", 8, 256, 256, 160, 160, 3, 1, 1, 1),   # VGG19 conv3_x
    "vgg3d": ("dgrad", 8, 256, 256, 160, 160, 3, 1, 1, 1),
    "vgg2": ("fwd", 8, 128, 128, 320, 320, 3, 1, 1, 1),   # VGG19 conv2_2
    "r2": ("fwd", 8, 128, 128, 80, 80, 3, 1, 1, 1),       # ResNet layer2 conv2
    "r4": ("fwd", 8, 512, 512, 80, 80, 3, 4, 4, 1),       # ResNet layer4 conv2 (dilation 4)
    "d3": ("fwd", 8, 512, 512, 39, 39, 4, 1, 1, 1),       # discriminator 512->512 4x4
    "sh8": ("fwd", 8, 32, 128, 640, 640, 1, 0, 1, 1),     # SPADE mlp_shared as a K=32 GEMM, painter batch of the full step
    "gb48_8": ("fwd", 8, 128, 48, 640, 640, 3, 1, 1, 1),  # gamma||beta, painter batch of the full step
    "gb80_640": ("fwd", 16, 128, 80, 640, 640, 3, 1, 1, 1),   # gamma||beta of the 40-channel block at 640^2 (C1 painter)
    "dg80": ("dgrad", 16, 128, 80, 640, 640, 3, 1, 1, 1),
    "wg80": ("wgrad", 16, 128, 80, 640, 640, 3, 1, 1, 1),
    "vgg1d": ("dgrad", 8, 64, 64, 640, 640, 3, 1, 1, 1),      # VGG19 conv1_2 dgrad through ReLU
    "vgg2d": ("dgrad", 8, 128, 128, 320, 320, 3, 1, 1, 1),
}
names = sys.argv[1:] or list(CASES)
reps = int(os.environ.get("REPS", "5"))
DACT = {"relu": _lib.ACT_RELU, "lrelu": _lib.ACT_LRELU, "none": _lib.ACT_NONE}[os.environ.get("DACT", "relu")]
for name in names:
    which, n, ci, co, h, w, k, pad, dil, stride = (tuple(CASES[name]) + (1, 1))[:10]
    ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x = torch.randn(n, h, w, ci, device=dev).bfloat16()
    wp = (torch.randn(co, k * k, ci, device=dev) * 0.05).bfloat16()
    bias = None if os.environ.get("NOBIAS") else torch.zeros(co, device=dev)
    g = ops.ConvGeom(k, k, stride, dil, pad, _lib.PAD_ZERO, _lib.ACT_NONE, 0.2, _lib.ENGINE_AUTO)
    gy = torch.randn(n, ho, wo, co, device=dev).bfloat16()
    def run():
        if which == "fwd":
            return ops.conv_fwd_raw(x, wp, bias, None, g)
        if which == "dgrad":
            return ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g, DACT, x if DACT != _lib.ACT_NONE else None)
        return ops.conv_wgrad_raw(x, gy, g, False)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * n * ho * wo * ci * co * k * k
    byts = 2.0 * n * (h * w * ci + ho * wo * co)
    print(f"{name:6s} {which:5s} {ci}->{co} k{k} d{dil} s{stride} @{h}x{w} n={n}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s  {byts/ms/1e6:.0f} GB/s (min HBM traffic)", flush=True)
