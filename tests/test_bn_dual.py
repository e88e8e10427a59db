"""BatchNorm's backward summing the gradients of its two consumers in its own first pass (ops.batchnorm_act(dual=True),
cgb_bn_train_bwd2): the ResNet bottleneck's output feeds the next block's conv1 and its identity branch
(resnetmulti_v2.py:40-56); autograd would add the two gradients with a separate pass over the tensor.  The encoder's gradients
with the dual hand-over against the plain composition — CPU through the emulated ABI, GPU through the kernels."""
import pytest
import torch

from climategan_b200 import ops
from climategan_b200.deeplab.resnetmulti_v2 import ResNetMulti
from tests.emulib import emulated_library


def _grads(dev, dual, size):
    torch.manual_seed(11)
    net = ResNetMulti([2, 1, 2, 1], n_res=0).to(dev).train()
    for p in net.parameters():
        p.requires_grad_(p.dim() == 4)           # conv weights (the BatchNorm affines are frozen in the reference, :16-18)
    x = ops.to_storage(torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(3)).to(dev), torch.bfloat16)
    old = ops._BN_DUAL
    ops._BN_DUAL = dual
    try:
        z = net.forward_storage(x)
        gz = torch.randn(z.shape, generator=torch.Generator().manual_seed(4)).to(dev).to(z.dtype)
        z.backward(gz)
    finally:
        ops._BN_DUAL = old
    return z.detach().float().cpu(), {k: p.grad.float().cpu() for k, p in net.named_parameters() if p.grad is not None}


def _compare(dev, size):
    z0, g0 = _grads(dev, False, size)
    zn, gn = _grads(dev, False, size)      # the plain composition a second time: the run-to-run noise of the GPU path (the conv
    z1, g1 = _grads(dev, True, size)       # epilogue's statistics and the wgrad reductions are atomics-ordered; 0 on the emulation)
    rel = lambda a, b: float((a - b).norm() / (a.norm() + 1e-30))
    assert rel(z0, z1) <= max(3 * rel(z0, zn), 1e-6)               # the forward is the same launches
    assert g0.keys() == g1.keys() and len(g0) > 20
    for k in g0:
        # the fused form adds the two bf16 gradients in fp32 (one rounding less per block): rounding-sized differences
        assert rel(g0[k], g1[k]) <= max(3 * rel(g0[k], gn[k]), 8e-2), (k, rel(g0[k], g1[k]), rel(g0[k], gn[k]))


def test_bn_dual_emulated():
    with emulated_library():
        _compare(torch.device("cpu"), 64)


@pytest.mark.gpu
def test_bn_dual_gpu(cuda):
    _compare(cuda, 128)
