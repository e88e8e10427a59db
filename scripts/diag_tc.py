"""Diagnostic: tcgen05 engine vs SIMT engine on the same bf16 operands (prints, no asserts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from climategan_b200 import _lib, ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
CASES = [
    # n, ci, co, h, w, k, stride, dil, pad
    (1, 64, 16, 16, 8, 1, 1, 1, 0),
    (1, 64, 16, 16, 8, 3, 1, 1, 1),
    (2, 128, 48, 16, 16, 3, 1, 1, 1),
    (2, 128, 80, 20, 12, 3, 1, 1, 1),
    (2, 8, 128, 16, 16, 3, 1, 1, 1),
    (1, 24, 24, 12, 12, 3, 1, 1, 1),
    (2, 40, 24, 8, 8, 1, 1, 1, 0),
    (3, 640, 640, 5, 5, 3, 1, 1, 1),
    (2, 128, 1280, 5, 5, 3, 1, 1, 1),
    (1, 64, 32, 20, 20, 3, 1, 6, 6),
    (2, 8, 16, 16, 16, 4, 2, 1, 1),
    (1, 320, 160, 40, 40, 3, 1, 1, 1),
    (2, 128, 48, 64, 64, 3, 1, 1, 1),
    (1, 24, 24, 40, 48, 3, 1, 1, 1),
    (2, 128, 80, 33, 50, 3, 1, 1, 1),
    (1, 64, 32, 64, 64, 3, 1, 2, 2),
    (1, 128, 160, 48, 40, 3, 1, 1, 1),
    (1, 48, 128, 64, 32, 3, 1, 1, 1),
    (1, 32, 128, 64, 64, 1, 1, 1, 0),
    (2, 64, 128, 32, 32, 4, 2, 1, 1),
    (1, 16, 24, 23, 19, 3, 2, 1, 1),
    (1, 8, 64, 64, 64, 7, 2, 1, 3),
    (2, 128, 64, 17, 17, 1, 2, 1, 0),
]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for idx, (n, ci, co, h, w, k, s, dil, pad) in enumerate(CASES):
    if only is not None and idx != only:
        continue
    x = torch.randn(n, h, w, ci, device=dev).bfloat16()
    wp = (torch.randn(co, k * k, ci, device=dev) / (ci * k * k) ** 0.5).bfloat16()
    bias = torch.randn(co, device=dev) * 0.1
    res = {}
    for eng in (_lib.ENGINE_SIMT, _lib.ENGINE_TCGEN05):
        g = ops.ConvGeom(k, k, s, dil, pad, _lib.PAD_ZERO, _lib.ACT_LRELU, 0.2, eng)
        y = ops.conv_fwd_raw(x, wp, bias, None, g)
        torch.cuda.synchronize()
        res[eng] = y.float()
    a, b = res[_lib.ENGINE_SIMT], res[_lib.ENGINE_TCGEN05]
    err = float((a - b).abs().max() / a.abs().max())
    msg = f"case {idx} {(n,ci,co,h,w,k,s,dil,pad)} fwd relmax {err:.3e}"
    if True:
        gy = torch.randn_like(res[_lib.ENGINE_SIMT]).bfloat16()
        mask = torch.randn(n, h, w, ci, device=dev).bfloat16()
        r2 = {}
        for eng in (_lib.ENGINE_SIMT, _lib.ENGINE_TCGEN05):
            g = ops.ConvGeom(k, k, s, dil, pad, _lib.PAD_ZERO, _lib.ACT_NONE, 0.2, eng)
            gx = ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g, _lib.ACT_LRELU, mask)
            torch.cuda.synchronize()
            r2[eng] = gx.float()
        a, b = r2[_lib.ENGINE_SIMT], r2[_lib.ENGINE_TCGEN05]
        msg += f" dgrad relmax {float((a - b).abs().max() / a.abs().max()):.3e}"
    print(msg, flush=True)

    if True:
        gy = torch.randn_like(res[_lib.ENGINE_SIMT]).bfloat16()
        r3 = {}
        for eng in (_lib.ENGINE_SIMT, _lib.ENGINE_TCGEN05):
            g = ops.ConvGeom(k, k, s, dil, pad, _lib.PAD_ZERO, _lib.ACT_NONE, 0.2, eng)
            gw, gb = ops.conv_wgrad_raw(x, gy, g, True)
            torch.cuda.synchronize()
            r3[eng] = (gw, gb)
        (a, ab), (b, bb) = r3[_lib.ENGINE_SIMT], r3[_lib.ENGINE_TCGEN05]
        print(f"     wgrad relmax {float((a - b).abs().max() / a.abs().max()):.3e} bias relmax {float((ab - bb).abs().max() / ab.abs().max()):.3e}", flush=True)
