"""Sensitivity of the full-step fixture (tests/golden/full_step.*): gradients of the CPU oracle under a 1e-7 / 1e-6 relative
perturbation of the generator weights.  usage: PYTHONPATH=. python scripts/sensitivity_full_step.py"""
import json, os, numpy as np, torch
from climategan_b200.utils import full_opts, synth_batch
from oracle import full_step_oracle as fo
from tests.golden.weights import fill_state_dict
from tests.helpers import GOLDEN
meta = json.load(open(os.path.join(GOLDEN, "full_step.json")))
size, batch = meta["size"], meta["batch"]
mk = lambda shapes, seed: fill_state_dict([(k, tuple(s)) for k, s in shapes], seed)
def run(eps):
    gsd, dsd, vsd = mk(meta["g_shapes"], 21), mk(meta["d_shapes"], 22), mk(meta["v_shapes"], 23)
    g_names = meta["g_param_names"]
    frozen = {k for k in g_names if ".bn" in k or "downsample.1" in k}
    torch.manual_seed(0)
    for k in g_names:
        if eps and gsd[k].dtype.is_floating_point and not k.endswith(("_u","_v")):
            gsd[k] = gsd[k] * (1 + eps * torch.randn_like(gsd[k]))
        gsd[k].requires_grad_(not k.endswith(("weight_u", "weight_v")) and not (k.startswith("encoder.") and k in frozen))
    opts = full_opts(size=size)
    mdb = synth_batch(opts, batch, size, 7)
    loss, terms = fo.full_g_loss(gsd, dsd, vsd, mdb, size // 16)
    loss.backward()
    return {k: gsd[k].grad.clone() for k in meta["full_g"]}, float(loss)
g0, l0 = run(0.0)
for eps in (1e-7, 1e-6):
    g1, l1 = run(eps)
    print("eps", eps, "loss", l0, l1)
    for k in g0:
        print("   ", k, float((g0[k]-g1[k]).abs().max()/g0[k].abs().max()))
