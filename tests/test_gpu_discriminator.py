"""GPU parity of the painter discriminator path (OmniDiscriminator['p'] on cat(real, fake), GANLoss, FeatMatchLoss,
HingeLoss) through libcgb200 against goldens produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from climategan_b200 import ops
from climategan_b200.discriminator import OmniDiscriminator, fc_discriminator_forward, get_fc_discriminator
from climategan_b200.losses import FeatMatchLoss, GANLoss, HingeLoss
from climategan_b200.utils import Dict
from tests.helpers import cosine, load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _divide(out):
    return ([[t[: t.size(0) // 2] for t in p] for p in out], [[t[t.size(0) // 2:] for t in p] for p in out])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_discriminator_matches_reference_golden(cuda, dtype):
    meta, g, sd, _ = load_golden("disc_small")
    opts = Dict(tasks=["p"], dis=dict(p=dict(ndf=meta["ndf"], n_layers=meta["n_layers"], norm="instance",
                                             use_sigmoid=False, num_D=meta["num_D"], get_intermediate_features=True,
                                             use_local_discriminator=False, init_type="xavier", init_gain=0.02)))
    D = OmniDiscriminator(opts, storage_dtype=dtype)
    assert [(k, tuple(v.shape)) for k, v in D.state_dict().items()] == [(k, tuple(s)) for k, s in meta["shapes"]]
    D.load_state_dict(sd, strict=True)
    D = D.to(cuda)
    real = torch.from_numpy(g["real"]).to(cuda)
    fake = torch.from_numpy(g["fake"]).to(cuda).requires_grad_(True)
    out = D["p"](torch.cat([real, fake], 0))
    assert len(out) == meta["num_D"] and len(out[0]) == meta["n_layers"] + 2
    pred_real, pred_fake = _divide(out)
    g_gan = GANLoss(use_lsgan=False)(pred_fake, True)
    g_feat = FeatMatchLoss()(pred_real, pred_fake)
    d_hinge = HingeLoss()(pred_fake, False, True) + HingeLoss()(pred_real, True, True)
    (g_gan + 10.0 * g_feat).backward()
    f32 = dtype == torch.float32
    # stated tolerances: fp32 storage 1e-4 / bf16 storage 3e-2 of full scale forward, losses 1e-5 / 2e-2 relative
    ftol, ltol = (1e-4, 1e-5) if f32 else (3e-2, 2e-2)
    for i in range(meta["num_D"]):
        assert rel_max(out[i][-1], torch.from_numpy(g[f"pred_{i}"])) < ftol
        assert rel_max(out[i][1], torch.from_numpy(g[f"feat_{i}_1"])) < ftol
    for name, val in (("g_gan", g_gan), ("g_feat", g_feat), ("d_hinge", d_hinge)):
        assert abs(float(val) - float(g[name])) / abs(float(g[name])) < ltol, name
    gr = torch.from_numpy(g["fake_grad"])
    if f32:
        assert rel_max(fake.grad, gr) < 1e-3
    else:
        assert cosine(fake.grad, gr) > 0.97 and rel_l2(fake.grad, gr) < 0.3
    params = dict(D.named_parameters())
    for k, v in g.items():
        if k.startswith("grad::"):
            gm, gref = params[k[6:]].grad, torch.from_numpy(v)
            if f32:
                assert rel_max(gm, gref) < 1e-3, k
            else:
                assert cosine(gm, gref) > 0.97, (k, cosine(gm, gref))
    if f32:
        norms_ref = dict(zip(meta["grad_keys"], g["grad_norms"]))
        for k in meta["grad_keys"]:
            if norms_ref[k] > 1e-6:
                assert abs(float(params[k].grad.norm()) - norms_ref[k]) / norms_ref[k] < 2e-3, k


def test_avgpool_and_instnorm_act(cuda):
    import torch.nn.functional as F

    torch.manual_seed(3)
    x = torch.randn(2, 4, 13, 10)
    xs = ops.to_storage(x.to(cuda), torch.float32).requires_grad_(True)
    y = ops.from_storage(ops.avgpool3s2(xs), 4)
    xr = x.clone().requires_grad_(True)
    yr = F.avg_pool2d(xr, 3, stride=2, padding=[1, 1], count_include_pad=False)
    assert rel_max(y, yr) < 1e-6
    gy = torch.randn_like(yr)
    yr.backward(gy)
    y.backward(gy.to(cuda))
    assert rel_max(ops.from_storage(xs.grad, 4), xr.grad) < 1e-6
    x = torch.randn(2, 20, 9, 7) * 2 + 1
    xs = ops.to_storage(x.to(cuda), torch.float32).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    yr = F.leaky_relu(F.instance_norm(xr), 0.2)
    y = ops.from_storage(ops.instnorm_act(xs, 2, 0.2), 20)
    assert rel_max(y, yr) < 1e-5
    gy = torch.randn_like(yr)
    yr.backward(gy)
    y.backward(gy.to(cuda))
    assert rel_max(ops.from_storage(xs.grad, 20), xr.grad) < 1e-4


def test_fc_discriminator(cuda):
    """AdvEnt discriminator of the masker heads (discriminator.py:327-361) vs the same convs in torch."""
    import torch.nn.functional as F

    torch.manual_seed(0)
    net = get_fc_discriminator(num_classes=2, ndf=8, use_norm=False)
    x = torch.randn(2, 2, 64, 64)
    ref = x
    for m in net:
        ref = F.conv2d(ref, m.weight, m.bias, stride=2, padding=1) if isinstance(m, torch.nn.Conv2d) else F.leaky_relu(ref, 0.2)
    out = fc_discriminator_forward(net.to(cuda), x.to(cuda), torch.float32)
    assert out.shape == (2, 1, 2, 2) and rel_max(out, ref) < 1e-5


@pytest.mark.parametrize("train_uv", [False, True])
def test_fc_discriminator_two_calls_before_backward(cuda, train_uv):
    """The D step runs each AdvEnt discriminator on the r and then the s batch before ONE backward.  In the reference the
    spectral-norm u / v Parameters are saved by reference and their .data swapped on every forward, so both backward passes
    use the last forward's u, v in d(sigma)/dW (and, once run_epoch has flipped requires_grad on them, u and v get gradients
    themselves).  Product (fp32) vs the oracle, which reproduces this through the same aliasing."""
    from climategan_b200 import ops
    from climategan_b200.discriminator import fc_discriminator_forward, get_fc_discriminator
    from oracle import full_step_oracle as fo
    from oracle.painter_oracle import SNState

    torch.manual_seed(0)
    net = get_fc_discriminator(num_classes=11, use_norm=True)
    sd = {"D." + k: v.detach().clone() for k, v in net.state_dict().items()}
    pa, pb = torch.softmax(torch.randn(2, 11, 32, 32), 1), torch.softmax(torch.randn(2, 11, 32, 32), 1)
    osd = {k: v.clone().requires_grad_(train_uv or not k.endswith(("_u", "_v"))) for k, v in sd.items()}
    sn = SNState(osd)
    lo = fo.advent(pa, 1, osd, sn, "D", None, wgan=False) + fo.advent(pb, 0, osd, sn, "D", None, wgan=False)
    lo.backward()
    net = net.to(cuda)
    for n, p in net.named_parameters():
        p.requires_grad_(train_uv or not n.endswith(("_u", "_v")))
    D = lambda t: fc_discriminator_forward(net, t, torch.float32)  # noqa: E731
    lp = (ops.const_target_loss(D(ops.prob_2_entropy(pa.to(cuda))), ops.LOSS_BCE_LOGITS, 1.0)
          + ops.const_target_loss(D(ops.prob_2_entropy(pb.to(cuda))), ops.LOSS_BCE_LOGITS, 0.0))
    lp.backward()
    assert abs(float(lp) - float(lo)) < 1e-5
    for n, p in net.named_parameters():
        go = osd["D." + n].grad
        assert (go is None) == (p.grad is None), n
        if go is not None:
            assert rel_max(p.grad, go) < 2e-4, n
        if n.endswith(("_u", "_v")):
            assert rel_max(p, osd["D." + n]) < 1e-5, n
