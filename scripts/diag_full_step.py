"""Per-parameter gradient agreement between the bf16 (tcgen05) and fp32 (SIMT) runs of the full G step on the GPU —
localises where the bf16 path departs.  usage: python scripts/diag_full_step.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_full_step import _build

dev = torch.device("cuda:0")
grads = {}
for dtype in (torch.float32, torch.bfloat16):
    meta, g, t, mdb = _build(dev, dtype)
    t.update_G(mdb)
    grads[dtype] = {k: p.grad.detach().float().clone() for k, p in t.G.named_parameters() if p.requires_grad and p.grad is not None}
    print(dtype, {k: round(v, 5) for k, v in __import__("tests.test_gpu_full_step", fromlist=["_flatten"])._flatten(t.losses_to_host()).items()})
rows = []
for k, a in grads[torch.float32].items():
    b = grads[torch.bfloat16][k]
    cos = float((a.flatten() @ b.flatten()) / (a.norm() * b.norm()).clamp_min(1e-30))
    rows.append((cos, k, float(a.norm()), float(b.norm())))
for cos, k, na, nb in rows:
    print(f"{cos:8.4f} {na:11.4e} {nb:11.4e} {k}")
