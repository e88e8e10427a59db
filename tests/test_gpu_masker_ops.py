"""GPU parity of the masker training-path operators (csrc/masker_ops.cu) against plain PyTorch fp32 references of the
same op on the same seeded inputs: train-mode BatchNorm (+act, +residual) fwd/bwd, max-pool / bilinear / reflect-pad /
channel-mean / global-mean adjoints, dropout statistics, and every masker loss with its gradient."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.helpers import rel_max

pytestmark = pytest.mark.gpu


def _st(x, dtype=torch.float32):
    return ops.to_storage(x, dtype)


def _nchw(x, c):
    return ops.from_storage(x, c)


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.mark.parametrize("act", [_lib.ACT_NONE, _lib.ACT_RELU, _lib.ACT_LRELU])
@pytest.mark.parametrize("with_res", [False, True])
@pytest.mark.parametrize("c,n,h,w", [(16, 3, 9, 7), (64, 2, 12, 12), (20, 2, 5, 6)])
def test_batchnorm_act_train(cuda, act, with_res, c, n, h, w):
    x = _rand(n, c, h, w, seed=1) * 2 + 0.5
    r = _rand(n, c, h, w, seed=2)
    gy = _rand(n, c, h, w, seed=3)
    bn = torch.nn.BatchNorm2d(c)
    bn.weight.data = 1 + 0.3 * _rand(c, seed=4)
    bn.bias.data = 0.2 * _rand(c, seed=5)
    bn_ref = torch.nn.BatchNorm2d(c)
    bn_ref.load_state_dict(bn.state_dict())
    # reference
    xr, rr = x.clone().requires_grad_(), r.clone().requires_grad_()
    y = bn_ref(xr) + (rr if with_res else 0)
    y = F.relu(y) if act == _lib.ACT_RELU else (F.leaky_relu(y, 0.2) if act == _lib.ACT_LRELU else y)
    y.backward(gy)
    # ours
    bn = bn.to(cuda)
    xd, rd = x.to(cuda).requires_grad_(), r.to(cuda).requires_grad_()
    yo = _nchw(ops.batchnorm_act(_st(xd), bn, _st(rd) if with_res else None, act, 0.2), c)
    yo.backward(gy.to(cuda))
    assert rel_max(yo, y) < 2e-5
    assert rel_max(xd.grad, xr.grad) < 2e-4
    if with_res:
        assert rel_max(rd.grad, rr.grad) < 1e-5
    assert rel_max(bn.weight.grad, bn_ref.weight.grad) < 2e-4
    assert rel_max(bn.bias.grad, bn_ref.bias.grad) < 2e-4
    assert rel_max(bn.running_mean, bn_ref.running_mean) < 1e-5
    assert rel_max(bn.running_var, bn_ref.running_var) < 1e-5
    assert int(bn.num_batches_tracked) == 1


def test_batchnorm_eval_and_frozen_affine(cuda):
    c = 24
    x = _rand(2, c, 6, 6, seed=1)
    bn = torch.nn.BatchNorm2d(c)
    bn.running_mean.data = 0.1 * _rand(c, seed=2)
    bn.running_var.data = 0.5 + _rand(c, seed=3).abs()
    for p in bn.parameters():
        p.requires_grad = False
    bn.eval()
    ref = bn(x)
    xd = x.to(cuda).requires_grad_()
    yo = _nchw(ops.batchnorm_act(_st(xd), bn.to(cuda), None, _lib.ACT_NONE), c)
    yo.sum().backward()
    assert rel_max(yo, ref) < 1e-5
    assert bn.weight.grad is None
    xr = x.clone().requires_grad_()
    bn.cpu()(xr).sum().backward()
    assert rel_max(xd.grad, xr.grad) < 1e-5


def test_batchnorm_act_train_bf16(cuda):
    """bf16 storage: same op against the fp32 reference evaluated on the bf16-rounded operands (tolerance 1.5e-2 of full scale,
    the bf16 epsilon being 3.9e-3)."""
    n, c, h, w = 4, 64, 24, 24
    q = lambda t: t.bfloat16().float()  # noqa: E731
    x, r, gy = q(_rand(n, c, h, w, seed=1) * 2 + 0.5), q(_rand(n, c, h, w, seed=2)), q(_rand(n, c, h, w, seed=3))
    bn_ref = torch.nn.BatchNorm2d(c)
    bn_ref.weight.data = 1 + 0.3 * _rand(c, seed=4)
    bn = torch.nn.BatchNorm2d(c)
    bn.load_state_dict(bn_ref.state_dict())
    xr, rr = x.clone().requires_grad_(), r.clone().requires_grad_()
    y = F.relu(bn_ref(xr) + rr)
    y.backward(gy)
    bn = bn.to(cuda)
    xd, rd = x.to(cuda).requires_grad_(), r.to(cuda).requires_grad_()
    yo = _nchw(ops.batchnorm_act(_st(xd, torch.bfloat16), bn, _st(rd, torch.bfloat16), _lib.ACT_RELU, 0.2), c)
    yo.backward(gy.to(cuda))
    assert rel_max(yo, y) < 1.5e-2
    assert rel_max(xd.grad, xr.grad) < 1.5e-2
    assert rel_max(bn.weight.grad, bn_ref.weight.grad) < 1.5e-2
    assert rel_max(bn.running_var, bn_ref.running_var) < 1e-4


@pytest.mark.parametrize("h,w", [(9, 9), (10, 13), (16, 16)])
def test_maxpool3s2_ceil_fwd_bwd(cuda, h, w):
    x = torch.relu(_rand(2, 8, h, w, seed=1)).round()  # many ties (post-ReLU zeros / repeated integers)
    xr = x.clone().requires_grad_()
    y = F.max_pool2d(xr, 3, 2, 0, ceil_mode=True)
    gy = _rand(*y.shape, seed=2)
    y.backward(gy)
    xd = x.to(cuda).requires_grad_()
    yo = _nchw(ops.maxpool3s2_ceil(_st(xd)), 8)
    yo.backward(gy.to(cuda))
    assert rel_max(yo, y) == 0
    assert rel_max(xd.grad, xr.grad) < 1e-6


@pytest.mark.parametrize("pad", [0, 1])
@pytest.mark.parametrize("n,c,h,w", [(2, 64, 80, 80), (1, 40, 37, 45), (3, 8, 33, 64)])
def test_maxpool3s2_bwd_tiled_matches_torch(cuda, pad, n, c, h, w):
    """The tiled backward (16-bit storage, maps >= 32 x 32: masker_ops.cu maxpool3s2_bwd_tiled_kernel) against F.max_pool2d's own
    backward: small-integer inputs (many ties -> the first-maximum rule matters) and integer gradients, so bf16 sums are exact."""
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, 6, (n, c, h, w), generator=g).float()
    xr = x.clone().requires_grad_()
    y = F.max_pool2d(xr, 3, 2, pad, ceil_mode=(pad == 0))
    gy = torch.randint(-8, 9, tuple(y.shape), generator=g).float()
    y.backward(gy)
    xd = x.to(cuda).requires_grad_()
    pool = ops.maxpool3s2_ceil if pad == 0 else ops.maxpool3s2_pad1
    yo = _nchw(pool(_st(xd, torch.bfloat16)), c)
    yo.backward(gy.to(cuda))
    assert torch.equal(yo.detach().float().cpu(), y.detach())
    assert torch.equal(xd.grad.float().cpu(), xr.grad)


@pytest.mark.parametrize("ac", [True, False])
@pytest.mark.parametrize("hi,wi,ho,wo", [(5, 7, 10, 14), (8, 8, 16, 16), (6, 5, 13, 9), (12, 12, 5, 7)])
def test_resize_bilinear_fwd_bwd(cuda, ac, hi, wi, ho, wo):
    x = _rand(2, 11, hi, wi, seed=1)
    xr = x.clone().requires_grad_()
    y = F.interpolate(xr, size=(ho, wo), mode="bilinear", align_corners=ac)
    gy = _rand(*y.shape, seed=2)
    y.backward(gy)
    xd = x.to(cuda).requires_grad_()
    yo = _nchw(ops.resize_bilinear(_st(xd), ho, wo, align_corners=ac), 11)
    yo.backward(gy.to(cuda))
    assert rel_max(yo, y) < 1e-5
    assert rel_max(xd.grad, xr.grad) < 1e-5


@pytest.mark.parametrize("pad", [1, 2, 3])
def test_reflect_pad_fwd_bwd(cuda, pad):
    x = _rand(2, 8, 6, 5, seed=1)
    xr = x.clone().requires_grad_()
    y = F.pad(xr, (pad,) * 4, mode="reflect")
    gy = _rand(*y.shape, seed=2)
    y.backward(gy)
    xd = x.to(cuda).requires_grad_()
    yo = _nchw(ops.reflect_pad(_st(xd), pad), 8)
    yo.backward(gy.to(cuda))
    assert rel_max(yo, y) == 0
    assert rel_max(xd.grad, xr.grad) < 1e-6


def test_channel_mean_global_mean_broadcast_mul(cuda):
    x = _rand(2, 20, 6, 7, seed=1)
    b = _rand(2, 20, 6, 7, seed=5)
    xr = x.clone().requires_grad_()
    br = b.clone().requires_grad_()
    cm = (xr * br).mean(1, keepdim=True)
    gm = F.adaptive_avg_pool2d(xr, 1)
    up = F.interpolate(gm, size=(6, 7), mode="bilinear", align_corners=True)
    g1, g2 = _rand(*cm.shape, seed=2), _rand(*up.shape, seed=3)
    (cm * g1).sum().backward(retain_graph=True)
    (up * g2).sum().backward()
    xd = x.to(cuda).requires_grad_()
    bd = b.to(cuda).requires_grad_()
    xs = _st(xd)
    cmo = _nchw(ops.channel_mean(ops.mul(xs, _st(bd)), 20), 1)
    upo = _nchw(ops.broadcast_hw(ops.global_mean(xs), 6, 7), 20)
    ((cmo * g1.to(cuda)).sum() + (upo * g2.to(cuda)).sum()).backward()
    assert rel_max(cmo, cm) < 1e-5 and rel_max(upo, up) < 1e-5
    assert rel_max(xd.grad, xr.grad) < 1e-5
    assert rel_max(bd.grad, br.grad) < 1e-5


def test_dropout_statistics_and_backward(cuda):
    torch.manual_seed(0)
    x = torch.ones(4, 64, 32, 32, device=cuda, requires_grad=True)
    y = _nchw(ops.dropout(_st(x), 0.5, True), 64)
    keep = float((y > 0).float().mean())
    assert abs(keep - 0.5) < 0.01
    assert set(torch.unique(y).tolist()) == {0.0, 2.0}
    y.sum().backward()
    assert torch.equal(x.grad, y.detach())            # same mask, same 1/(1-p) scale
    assert ops.dropout(_st(x), 0.5, False) is not None and torch.equal(_nchw(ops.dropout(_st(x), 0.5, False), 64), x)


def test_softmax_cross_entropy_entropy(cuda):
    n, c, h, w = 2, 11, 9, 8
    logits = _rand(n, c, h, w, seed=1) * 2
    tgt = torch.randint(0, c, (n, h, w), generator=torch.Generator().manual_seed(2))
    depth = _rand(n, 1, h, w, seed=3).abs()
    lr = logits.clone().requires_grad_()
    ce = F.cross_entropy(lr, tgt)
    p = torch.softmax(lr, 1)
    ent = -p * torch.log2(p + 1e-30) / np.log2(c) * depth
    gent = _rand(*ent.shape, seed=4)
    (ce + (ent * gent).sum()).backward()
    ld = logits.to(cuda).requires_grad_()
    ceo = ops.cross_entropy_nchw(ld, tgt.to(cuda))
    po = ops.softmax_nchw(ld)
    ento = ops.prob_2_entropy(po, depth.to(cuda))
    (ceo + (ento * gent.to(cuda)).sum()).backward()
    assert abs(float(ceo) - float(ce)) < 1e-5 * abs(float(ce))
    assert rel_max(po, p) < 1e-5 and rel_max(ento, ent) < 1e-5
    assert rel_max(ld.grad, lr.grad) < 1e-4


@pytest.mark.parametrize("version", [1, 2])
def test_minent_loss(cuda, version):
    n, c, h, w = 2, 2, 12, 10
    p = torch.softmax(_rand(n, c, h, w, seed=1), 1)
    pr = p.clone().requires_grad_()
    e = -pr * torch.log2(pr + 1e-30) / np.log2(c)
    if version == 1:
        ref = e.sum() / (n * h * w)
    else:
        dm = e - e.sum() / (n * h * w)
        ref = (e + 0.1 * dm * dm).sum() / (n * h * w)
    ref.backward()
    pd = p.to(cuda).requires_grad_()
    out = ops.minent_loss(pd, version, 0.1)
    out.backward()
    assert abs(float(out) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_max(pd.grad, pr.grad) < 1e-4


def test_mask_head_losses(cuda):
    n, h, w = 2, 16, 12
    logits = _rand(n, 1, h, w, seed=1)
    target = (_rand(n, 1, h, w, seed=2) > 0).float()
    lr = logits.clone().requires_grad_()
    pp = torch.sigmoid(lr)
    prob = torch.cat([pp, 1 - pp], 1)
    h_tv = ((pp[:, :, 1:] - pp[:, :, :-1]) ** 2).sum()
    w_tv = ((pp[:, :, :, 1:] - pp[:, :, :, :-1]) ** 2).sum()
    tv = 2 * (h_tv / ((h - 1) * w) + w_tv / (h * (w - 1))) / n
    bce = F.binary_cross_entropy_with_logits(lr, target)
    gi = torch.mean(1.0 * ((target - pp) > 0.5))
    gprob = _rand(*prob.shape, seed=3)
    (tv + bce + (prob * gprob).sum()).backward()
    ld = logits.to(cuda).requires_grad_()
    probo = ops.sigmoid_pair(ld)
    tvo = ops.tv_loss(probo[:, :1])
    bceo = ops.bce_logits_loss(ld, target.to(cuda))
    gio = ops.ground_intersection_loss(probo[:, :1], target.to(cuda))
    (tvo + bceo + (probo * gprob.to(cuda)).sum()).backward()
    assert rel_max(probo, prob) < 1e-6
    assert abs(float(tvo) - float(tv)) < 1e-5 * abs(float(tv))
    assert abs(float(bceo) - float(bce)) < 1e-5 * abs(float(bce))
    assert abs(float(gio) - float(gi)) < 1e-6
    assert rel_max(ld.grad, lr.grad) < 1e-4


def _sigm_ref(prediction, target, gmweight=0.5, scale=4):
    """SIGMLoss.__call__ (losses.py:250-278) restated for the test."""
    t_pred, t_targ = torch.median(prediction), torch.median(target)
    s_pred, s_targ = torch.mean(torch.abs(prediction - t_pred)), torch.mean(torch.abs(target - t_targ))
    R = (prediction - t_pred) / s_pred - (target - t_targ) / s_targ
    bs, num_pix = prediction.shape[0], prediction.shape[-1] * prediction.shape[-2]
    sx = torch.tensor([[1., 0, -1], [2, 0, -2], [1, 0, -1]]).expand(bs, 1, 3, 3)
    sy = torch.tensor([[1., 2, 1], [0, 0, 0], [-1, -2, -1]]).expand(bs, 1, 3, 3)
    gm = 0
    for k in range(scale):
        R_ = F.interpolate(R, scale_factor=1 / 2 ** k)
        gm = gm + torch.sum(torch.abs(F.conv2d(R_, sx)) + torch.abs(F.conv2d(R_, sy)))
    return 0.5 / num_pix * torch.sum(torch.abs(R)) + gmweight / num_pix * gm


@pytest.mark.parametrize("n,h,w", [(1, 32, 32), (2, 40, 24), (3, 32, 48)])
def test_sigm_loss(cuda, n, h, w):
    # the reference expands its Sobel kernels to (batch, 1, 3, 3): F.conv2d then yields `batch` identical output maps per
    # image, i.e. the gradient-matching term carries a factor `batch` (losses.py:262-270) — kept bug-compatible.
    pred = _rand(n, 1, h, w, seed=1)
    targ = _rand(n, 1, h, w, seed=2).abs()
    pr = pred.clone().requires_grad_()
    ref = _sigm_ref(pr, targ)
    ref.backward()
    pd = pred.to(cuda).requires_grad_()
    out = ops.sigm_loss(pd, targ.to(cuda), 0.5, 4)
    out.backward()
    assert abs(float(out) - float(ref)) < 2e-5 * abs(float(ref)), (float(out), float(ref))
    assert rel_max(pd.grad, pr.grad) < 2e-4
