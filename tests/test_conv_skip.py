"""ops.conv2d_skip (the ResNet bottleneck's conv1 sharing its input with the identity branch, resnetmulti_v2.py:40-56): the skip
gradient is added in the dgrad launch's epilogue instead of a separate pass.  Value, data gradient and weight gradient against
the plain composition conv2d(x) + x on the same operands — CPU through the emulated ABI, GPU through the kernels (where the
fused form rounds the sum once: tolerance = one bf16 rounding)."""
import pytest
import torch

from climategan_b200 import _lib, ops
from tests.emulib import emulated_library


def _run(dev, n=2, h=24, w=40, ci=64, co=32):
    g = torch.Generator().manual_seed(5)
    x = ops.to_storage(torch.randn(n, ci, h, w, generator=g).to(dev), torch.bfloat16)
    wt = (torch.randn(co, ci, 1, 1, generator=g) * 0.1).to(dev)
    gy = torch.randn(n, h, w, co, generator=g).to(dev).bfloat16()
    gs = torch.randn(n, h, w, ci, generator=g).to(dev).bfloat16()
    res = {}
    for fused in (False, True):
        xs, wd = x.clone().requires_grad_(), wt.clone().requires_grad_()
        if fused:
            y, partial, skip = ops.conv2d_skip(xs, wd, want_stats=False)
            assert partial is None and skip.data_ptr() == xs.data_ptr()
        else:
            y, skip = ops.conv2d(xs, wd), xs
        (y.float() * gy.float()).sum().backward(retain_graph=True) if False else torch.autograd.backward([y, skip], [gy, gs])
        res[fused] = (y.detach().float().cpu(), xs.grad.float().cpu(), wd.grad.cpu())
    (y0, gx0, gw0), (y1, gx1, gw1) = res[False], res[True]
    assert torch.equal(y0, y1)
    assert float((gw0 - gw1).abs().max()) <= 1e-4 * float(gw0.abs().max())   # same launch twice: atomics-ordered fp32 reduction
    # unfused: bf16(dgrad) + bf16 skip, rounded again; fused: one rounding of the fp32 sum -> within one bf16 ulp of each other
    assert float((gx0 - gx1).abs().max()) <= 2 ** -7 * float(gx0.abs().max())
    return gx0, gx1


def test_conv_skip_emulated():
    with emulated_library():
        _run(torch.device("cpu"))


@pytest.mark.gpu
def test_conv_skip_gpu(cuda):
    gx0, gx1 = _run(cuda, n=2, h=80, w=80, ci=1024, co=256)     # the layer-3 shape of the benchmarked step
    assert float((gx0 - gx1).abs().mean()) < 2e-3 * float(gx0.abs().mean())
