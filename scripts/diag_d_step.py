"""fp32 product vs golden, per-parameter gradient-norm errors for the D step (and G step) of the full-step fixture."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.test_gpu_full_step import _build

dev = torch.device("cuda:0")
meta, g, t, mdb = _build(dev, torch.float32)
t.update_G(mdb)
gn = np.array([float(p.grad.norm()) if p.grad is not None and p.requires_grad else -1.0 for _, p in t.G.named_parameters()])
for n, a, b in zip(meta["g_param_names"], gn, g["G.gradnorm"]):
    if b >= 0 and abs(a - b) > 5e-4 * b + 1e-9:
        print("G", n, a, b, abs(a - b) / max(b, 1e-30))
t.update_D(mdb)
dn = np.array([float(p.grad.norm()) if p.grad is not None and p.requires_grad else -1.0 for _, p in t.D.named_parameters()])
for n, a, b in zip(meta["d_param_names"], dn, g["D.gradnorm"]):
    if b >= 0 and abs(a - b) > 5e-4 * b + 1e-9:
        print("D", n, a, b, abs(a - b) / max(b, 1e-30))
print("done")
