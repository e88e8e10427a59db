"""Conv2dBlock with every option of the reference (climategan/blocks.py:49-147: pad zero / reflect / replicate; norm none /
spectral / batch / instance / layer / adain; activation relu / lrelu / prelu / selu / tanh / sigmoid / none) and the two MUNIT
norms (norms.py:8-81), against the REFERENCE's own modules on the same weights — forward, input gradient and every parameter
gradient — on CPU through the plain-PyTorch emulation of the C ABI (tests/emulib.py); the kernels themselves are compared with
the same emulation on the GPU (tests/test_gpu_zz_new_kernels.py)."""
import pytest
import torch

from climategan_b200 import ops
from climategan_b200.blocks import Conv2dBlock
from oracle import refshim
from tests.emulib import emulated_library

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")]

COMBOS = [
    # norm, activation, pad_type, kernel, padding, stride
    ("instance", "relu", "reflect", 3, 1, 1),
    ("instance", "tanh", "zero", 3, 1, 2),
    ("layer", "lrelu", "replicate", 3, 1, 1),
    ("layer", "selu", "zero", 1, 0, 1),
    ("adain", "relu", "zero", 3, 1, 1),
    ("adain", "prelu", "replicate", 3, 2, 1),
    ("batch", "prelu", "zero", 3, 1, 1),
    ("none", "selu", "replicate", 3, 1, 1),
    ("none", "prelu", "reflect", 3, 1, 1),
    ("spectral", "sigmoid", "replicate", 3, 1, 1),
    ("batch", "selu", "reflect", 3, 1, 1),
]


@pytest.mark.parametrize("combo", COMBOS, ids=lambda c: "-".join(map(str, c)))
def test_conv2dblock_option_matches_reference_module(combo):
    norm, act, pad_type, k, padding, stride = combo
    blocks_mod = refshim.load("blocks")
    torch.manual_seed(COMBOS.index(combo))
    cin, cout, n, h, w = 6, 12, 3, 11, 9          # 12 logical channels -> 16 storage channels: the pad channels must stay zero
    ref = blocks_mod.Conv2dBlock(cin, cout, k, stride, padding, norm=norm, activation=act, pad_type=pad_type).train()
    mine = Conv2dBlock(cin, cout, k, stride, padding, norm=norm, activation=act, pad_type=pad_type).train()
    assert list(mine.state_dict()) == list(ref.state_dict())
    mine.load_state_dict(ref.state_dict(), strict=True)
    if norm == "adain":
        wgt, b = torch.rand(n * cout) + 0.5, torch.randn(n * cout)
        ref.norm.weight, ref.norm.bias = wgt.clone().requires_grad_(True), b.clone().requires_grad_(True)
        mine.norm.weight, mine.norm.bias = wgt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    x = torch.randn(n, cin, h, w)
    gy = None
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gy = torch.randn_like(yr)
    yr.backward(gy)
    with emulated_library():
        xs = ops.to_storage(x, torch.float32).requires_grad_(True)
        ys = mine(xs)
        if act != "sigmoid":   # (sigmoid(0) = 0.5 lands in the pad channels of the conv epilogue; consumers never read them:
            assert float(ys.detach()[..., cout:].abs().max()) == 0.0   # packed weights are zero there, from_storage drops them)
        y = ops.from_storage(ys, cout)
        y.backward(gy)
        gx = ops.from_storage(xs.grad, cin)
    assert torch.allclose(y, yr, atol=2e-5, rtol=1e-4), float((y - yr).abs().max())
    assert torch.allclose(gx, xr.grad, atol=5e-5, rtol=1e-3), float((gx - xr.grad).abs().max())
    rp, mp = dict(ref.named_parameters()), dict(mine.named_parameters())
    for name, p in rp.items():
        if p.grad is None:
            assert mp[name].grad is None or float(mp[name].grad.abs().max()) == 0.0, name
            continue
        assert mp[name].grad is not None, name
        assert torch.allclose(mp[name].grad, p.grad, atol=1e-4, rtol=2e-3), (name, float((mp[name].grad - p.grad).abs().max()))
    if norm == "adain":
        assert torch.allclose(mine.norm.weight.grad, ref.norm.weight.grad, atol=1e-4, rtol=2e-3)
        assert torch.allclose(mine.norm.bias.grad, ref.norm.bias.grad, atol=1e-4, rtol=2e-3)
    if norm == "batch":
        assert torch.allclose(mine.norm.running_var, ref.norm.running_var, atol=1e-5)


def test_conv2dblock_rejects_unknown_options():
    with pytest.raises(ValueError):
        Conv2dBlock(4, 8, 3, norm="group")
    with pytest.raises(ValueError):
        Conv2dBlock(4, 8, 3, activation="gelu")
    with pytest.raises(AssertionError):
        Conv2dBlock(4, 8, 3, pad_type="circular")
