"""Sensitivity of a train-step fixture (default tests/golden/masker_step_spade.*): gradients of the REFERENCE Trainer's first
update_G / update_D under a 1e-7 / 1e-6 / 1e-3 relative perturbation of the generator weights (build container only; needs
/root/reference).  The numbers set the tolerances of tests/test_gpu_full_step.py::test_spade_masker_step_* and
test_base_depth_classify_step_*.
usage: PYTHONPATH=. python scripts/sensitivity_spade_step.py [fixture name]"""
import json
import os
import sys

import numpy as np
import torch

from oracle import ref_trainer as rt
from oracle import refshim
from tests.golden.weights import fill_state_dict
from tests.helpers import GOLDEN

CASE = sys.argv[1] if len(sys.argv) > 1 else "masker_step_spade"
meta = json.load(open(os.path.join(GOLDEN, CASE + ".json")))
size, batch = meta["size"], meta["batch"]
refshim.load("blocks").SPADEResnetBlock.cuda = lambda self, *a, **k: self
if (meta.get("overrides") or {}).get("gen.encoder.architecture") == "deeplabv3":   # the fixture's shallow ResNet (make_golden.py)
    _dl, _rn = refshim.load("deeplab", "deeplab.resnet101_v3")
    _nb = list(meta["overrides"]["gen.deeplabv3.nblocks"])
    _dl.ResNet101 = lambda output_stride=8, BatchNorm=None, verbose=0, no_init=False: _rn.ResNet(
        _rn.Bottleneck, _nb, output_stride, BatchNorm, verbose=verbose, no_init=no_init)


def run(eps):
    opts = rt.full_opts(size=size, tasks=tuple(meta["tasks"]), use_spade=meta.get("use_spade", False), overrides=meta.get("overrides"))
    t = rt.build_reference_trainer(opts, size)
    rt.load_weights(t)
    torch.manual_seed(0)
    with torch.no_grad():
        for k, p in t.G.named_parameters():
            if eps and not k.endswith(("_u", "_v")):
                p.mul_(1 + eps * torch.randn_like(p))
    mdb = rt.synth_batch(opts, batch, size, seed=7)
    for p in t.D.parameters():
        p.requires_grad = False
    t.update_G(mdb)
    gp = dict(t.G.named_parameters())
    norms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for p in gp.values()])
    grads = {"G.grad::" + k: gp[k].grad.clone() for k in meta["full_g"]}
    for p in t.D.parameters():
        p.requires_grad = True
    t.update_D(mdb)
    dp = dict(t.D.named_parameters())
    dnorms = np.array([float(p.grad.norm()) if p.grad is not None else -1.0 for p in dp.values()])
    grads.update({"D.grad::" + k: dp[k].grad.clone() for k in meta["full_d"]})
    return norms, dnorms, grads, list(gp), list(dp)


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


n0, d0, g0, names, dnames = run(0.0)
# 1e-7 / 1e-6: one fp32 ulp (the fp32-storage test's noise floor); 1e-3: the size of bf16 rounding (2^-9 relative)
for eps in (1e-7, 1e-6, 1e-3):
    n1, d1, g1, _, _ = run(eps)
    for tag, a, b, nm in (("G", n0, n1, names), ("D", d0, d1, dnames)):
        rel = np.where(a > 1e-4, np.abs(b - a) / np.maximum(a, 1e-30), 0.0)
        order = np.argsort(-rel)[:5]
        print("eps", eps, tag, "worst gradient-norm changes (norm > 1e-4):", [(nm[i], round(float(rel[i]), 5)) for i in order])
    for k in g0:
        print("   ", k, "max-rel", float((g0[k] - g1[k]).abs().max() / g0[k].abs().max()), "cos", round(cos(g0[k], g1[k]), 4))
