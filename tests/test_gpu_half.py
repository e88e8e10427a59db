"""fp16 storage — the reference's inference-only ``--half`` mode (apply_events.py:467-468, trainer.py:263-264; BASELINE.json
configs[4]) — on the GPU: tcgen05 ``kind::f16`` on fp16 operands against fp64 on the same fp16-rounded operands, and
``Trainer.infer_all(half=True)`` against the reference Trainer's own outputs (tests/golden/infer_all.*)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.helpers import GOLDEN, rel_max

pytestmark = pytest.mark.gpu

CASES = [
    # n, ci, co, h, w, k, stride, dil, pad, act
    (2, 128, 40, 48, 40, 3, 1, 1, 1, _lib.ACT_NONE),     # SPADE gamma/beta (weight-stationary halo kernel on a big enough map)
    (2, 256, 256, 20, 20, 3, 1, 2, 2, _lib.ACT_RELU),    # ResNet conv2 d2 with the folded BN's ReLU
    (2, 256, 1024, 20, 20, 1, 1, 1, 0, _lib.ACT_NONE),   # 1x1, four N tiles
    (2, 8, 64, 64, 64, 4, 2, 1, 1, _lib.ACT_LRELU),      # stride 2
    (1, 2048, 256, 16, 16, 3, 1, 12, 12, _lib.ACT_NONE), # ASPP d12, K = 18432
]


@pytest.mark.parametrize("case", CASES)
def test_fp16_conv_fwd_and_dgrad_on_tcgen05(cuda, case):
    n, ci, co, h, w, k, stride, dil, pad, act = case
    torch.manual_seed(CASES.index(case) + 40)
    q = lambda t: t.half().float()  # noqa: E731
    x = q(torch.randn(n, ci, h, w))
    wt = q(torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5)
    b = torch.randn(co) * 0.1
    xr = x.double().requires_grad_(True)
    yr = F.conv2d(xr, wt.double(), b.double(), stride=stride, padding=pad, dilation=dil)
    yr = {_lib.ACT_NONE: lambda t: t, _lib.ACT_RELU: F.relu, _lib.ACT_LRELU: lambda t: F.leaky_relu(t, 0.2)}[act](yr)
    xs = ops.to_storage(x.to(cuda), torch.float16)
    assert xs.dtype == torch.float16
    g = ops.ConvGeom(k, k, stride, dil, pad, _lib.PAD_ZERO, act, 0.2, _lib.ENGINE_TCGEN05)
    wp = ops.pack_weight(wt.to(cuda), torch.float16, cis=xs.shape[-1])
    y = ops.conv_fwd_raw(xs, wp, ops.pad_bias(b.to(cuda), wp.shape[0]), None, g)
    assert y.dtype == torch.float16
    # fp16 output rounding: 2^-11 relative per element (8x tighter than bf16's 2^-8)
    assert rel_max(ops.from_storage(y, co), yr) < 1.5e-3
    if act == _lib.ACT_NONE:
        gy = q(torch.randn_like(yr).float())
        yr.backward(gy.double())
        gx = ops.conv_dgrad_raw(ops.to_storage(gy.to(cuda), torch.float16), wp, tuple(xs.shape), g)
        assert rel_max(ops.from_storage(gx, ci), xr.grad) < 1.5e-3


def test_infer_all_half_matches_reference_trainer(cuda):
    """Trainer.infer_all(half=True): fp16 storage for the generator, against the reference Trainer's own (fp32) uint8 events:
    within 2 LSB on >= 99 % of the pixels, and closer than the bf16 mode on the same fixture, whose stated tolerance is 8 LSB on
    97 % (tests/test_gpu_infer_all.py).  The printout is the per-mode row of DESIGN.md's parity table."""
    import random

    from tests.test_gpu_infer_all import _trainer

    errs = {}
    for mode in ("bf16", "fp16", "fp32"):
        meta, g, t, x = _trainer(cuda, torch.float32 if mode == "fp32" else torch.bfloat16)
        random.seed(meta["seeds"]["random"])
        out = t.infer_all(x.clone(), numpy=True, bin_value=0.5, return_masks=True, half=(mode == "fp16"))
        assert t.G.storage_dtype == (torch.float32 if mode == "fp32" else torch.bfloat16)   # half=True is per call
        errs[mode] = {}
        for ev in ("flood", "wildfire", "smog"):
            d = np.abs(out[ev][:, ::2, ::2].astype(np.int32) - g[ev].astype(np.int32))
            errs[mode][ev] = (round(float((d <= 1).mean()), 4), round(float((d <= 2).mean()), 4), round(float(d.mean()), 4), int(d.max()))
        errs[mode]["mask_mismatch"] = round(float((out["mask"][:, :, ::2, ::2] != g["mask"]).mean()), 5)
    print("\ninfer_all uint8 agreement with the reference Trainer (frac <= 1 LSB, frac <= 2 LSB, mean |d| LSB, max LSB):")
    for mode, e in errs.items():
        print("  ", mode, e)
    for ev in ("flood", "wildfire", "smog"):
        assert errs["fp16"][ev][1] >= 0.99, (ev, errs)
        assert errs["fp16"][ev][2] <= errs["bf16"][ev][2] + 1e-3, (ev, errs)
    assert errs["fp16"]["mask_mismatch"] <= 5e-3
