"""First-layer convs on images (discriminator model0: 4x4 stride 2 on 3 or 4 channels, discriminator.py:122; VGG conv1_1) routed
as im2col + one K = round8(k*k*c) GEMM (ops.conv2d_first_layer, cgb_im2col_strided / cgb_col2im_strided): forward, data gradient
(the col2im adjoint), weight and bias gradients against torch.nn.functional.conv2d — on CPU through the emulated ABI, on the GPU
through the kernels."""
import pytest
import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.emulib import emulated_library

CASES = [(3, 16, 4, 2, 1, 64, 66), (4, 24, 4, 2, 1, 66, 64), (3, 16, 3, 1, 1, 64, 64)]   # cin, cout, k, stride, pad, h, w


def _run(cin, cout, k, stride, pad, h, w, dev):
    g = torch.Generator().manual_seed(cin * 100 + k)
    n = 2
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) * 0.2).bfloat16().float()
    b = torch.randn(cout, generator=g)
    xr, wr, br = x.clone().requires_grad_(), wt.clone().requires_grad_(), b.clone().requires_grad_()
    yr = F.leaky_relu(F.conv2d(xr, wr, br, stride=stride, padding=pad), 0.2)
    gy = torch.randn(yr.shape, generator=g).bfloat16().float()
    yr.backward(gy)
    xs = ops.to_storage(x.to(dev), torch.bfloat16).requires_grad_()
    wd, bd = wt.to(dev).requires_grad_(), b.to(dev).requires_grad_()
    calls = {}
    y = ops.conv2d_first_layer(xs, wd, bd, stride=stride, pad=pad, act=_lib.ACT_LRELU, slope=0.2)
    yo = ops.from_storage(y, cout)
    yo.backward(gy.to(dev))
    gx = ops.from_storage(xs.grad, cin)
    assert float(xs.grad[..., cin:].abs().max()) == 0.0          # the storage pad channels of the data gradient stay zero
    tol = 2e-2                                                     # bf16 storage of y, gx and of the patches' gradient
    assert float((yo.cpu() - yr).abs().max()) < tol * float(yr.abs().max())
    assert float((gx.cpu() - xr.grad).abs().max()) < tol * float(xr.grad.abs().max())
    assert float((wd.grad.cpu() - wr.grad).abs().max()) < tol * float(wr.grad.abs().max())
    assert float((bd.grad.cpu() - br.grad).abs().max()) < tol * float(br.grad.abs().max())
    return calls


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_first_layer_im2col_emulated(case):
    with emulated_library():
        _run(*case, torch.device("cpu"))


def test_route_is_taken(monkeypatch):
    """The route is the im2col one (a regression to the k*k-tap conv would still pass the numerics)."""
    seen = []
    with emulated_library():
        real = ops.im2col_strided
        monkeypatch.setattr(ops, "im2col_strided", lambda *a, **k: (seen.append(a[1:]), real(*a, **k))[1])
        x = ops.to_storage(torch.randn(1, 3, 64, 64), torch.bfloat16)
        ops.conv2d_first_layer(x, torch.randn(8, 3, 4, 4), None, stride=2, pad=1)
        assert seen == [(3, 4, 1, 1, 2)]
        seen.clear()
        ops.conv2d_first_layer(x, torch.randn(8, 3, 1, 1), None)                      # 1x1: direct
        ops.conv2d_first_layer(ops.to_storage(torch.randn(1, 3, 32, 32), torch.bfloat16), torch.randn(8, 3, 4, 4), None, stride=2, pad=1)
        assert seen == []


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_first_layer_im2col_gpu(cuda, case):
    _run(*case, cuda)
