"""Deterministic, version-stable synthetic weights for the golden fixtures.

``np.random.RandomState`` streams are frozen by NumPy's compatibility policy, so the same
(key order, shapes, seed) always yields the same state_dict — the fixtures therefore store only
inputs and reference outputs, not megabytes of weights.
"""
from __future__ import annotations

import numpy as np
import torch


def fill_state_dict(shapes, seed: int):
    """shapes: ordered list of (key, shape).  Returns {key: fp32 tensor}.

    conv weights ~ N(0, 0.7^2/fan_in) ; biases ~ N(0, 0.1^2) ; spectral-norm u/v: unit-norm gaussians
    (as the reference initialises them, climategan/norms.py:129-133) ; BatchNorm: weight 1+0.1n, running_mean 0.1n,
    running_var 0.5+0.5|n|, num_batches_tracked 0.
    """
    rs = np.random.RandomState(seed)
    out = {}
    for key, shape in shapes:
        shape = tuple(shape)
        a = rs.standard_normal(size=shape).astype(np.float32)
        if key.endswith("num_batches_tracked"):
            out[key] = torch.zeros(shape, dtype=torch.int64)
            continue
        if key.endswith("running_var"):
            a = np.abs(a) * 0.5 + 0.5
        elif key.endswith("running_mean"):
            a = a * 0.1
        elif len(shape) == 1 and key.endswith("weight"):  # BatchNorm scale
            a = 1.0 + 0.1 * a
        elif key.endswith("weight_u") or key.endswith("weight_v"):
            a = a / (np.linalg.norm(a) + 1e-12)
        elif key.endswith("bias"):
            a = a * 0.1
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
            a = a * (0.7 / np.sqrt(fan_in))
        out[key] = torch.from_numpy(a.astype(np.float32))
    return out


def synth_inputs(n: int, s: int, seed: int):
    """x ~ U(-1,1) [n,3,s,s] ; m ~ Bernoulli(.5) blocks [n,1,s,s] ; target ~ U(-1,1)."""
    rs = np.random.RandomState(seed)
    x = (rs.random_sample((n, 3, s, s)) * 2 - 1).astype(np.float32)
    m = (rs.random_sample((n, 1, s, s)) > 0.5).astype(np.float32)
    t = (rs.random_sample((n, 3, s, s)) * 2 - 1).astype(np.float32)
    return torch.from_numpy(x), torch.from_numpy(m), torch.from_numpy(t)
