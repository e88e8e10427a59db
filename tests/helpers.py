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
