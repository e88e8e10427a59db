"""Shared test helpers: golden loading, state_dict construction, error metrics."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from tests.golden.weights import fill_state_dict, synth_inputs

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name="painter_small"):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    arrs = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    shapes = [(k, tuple(s)) for k, s in meta["shapes"]]
    sd = fill_state_dict(shapes, meta["weight_seed"])
    if "input_seed" not in meta:
        return meta, arrs, sd, None
    x, m, target = synth_inputs(meta["batch"], meta["size"], meta["input_seed"])
    return meta, arrs, sd, (x, m, target)


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b|  (the tolerance form SURVEY.md §8d states)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b) -> float:
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b) -> float:
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def torch_diff_aug(x, params, cut_h, cut_w):
    """Plain PyTorch fp32 statement of ops.diff_aug (autograd-differentiable), from the contract in include/cgb200.h: colour
    jitter, then an integer translation with zero fill, then the cutout box.  Pinned to the reference's DiffTransforms in
    tests/test_diff_aug.py; the GPU suite compares the kernels with it."""
    n, c, h, w = x.shape
    b, cf, sf = (params[:, k].view(n, 1, 1, 1).to(x) for k in range(3))
    v = x + b
    m = v.mean((1, 2, 3), keepdim=True)
    v = (v - m) * cf + m
    mc = v.mean(1, keepdim=True)
    v = (v - mc) * sf + mc
    rows = []
    for k in range(n):
        tx, ty, ox, oy = (int(params[k, q]) for q in (3, 4, 5, 6))
        i0, i1, j0, j1 = max(0, -tx), min(h, h - tx), max(0, -ty), min(w, w - ty)
        moved = torch.nn.functional.pad(v[k, :, i0 + tx:i1 + tx, j0 + ty:j1 + ty], (j0, w - j1, i0, h - i1))
        keep = torch.ones(h, w, dtype=x.dtype, device=x.device)
        if cut_h > 0:
            a, bb = ox - cut_h // 2, oy - cut_w // 2
            keep[max(a, 0):min(a + cut_h - 1, h - 1) + 1, max(bb, 0):min(bb + cut_w - 1, w - 1) + 1] = 0
        rows.append(moved * keep)
    return torch.stack(rows)
