"""GPU parity tests of the individual libcgb200 kernels (through the C ABI via climategan_b200.ops)
against plain PyTorch fp32/fp64 CPU references of the same op."""
import pytest
import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.helpers import cosine, rel_l2, rel_max

pytestmark = pytest.mark.gpu

ENGINES = [_lib.ENGINE_SIMT, _lib.ENGINE_AUTO]


def _st(x, dtype, dev):
    return ops.to_storage(x.to(dev), dtype)


def _tol(dtype):
    return 2e-5 if dtype == torch.float32 else 1.5e-2


def _q(x, dtype):
    """round to the storage dtype (the reference is run on storage-rounded operands)"""
    return x.to(dtype).float()


CONV_CASES = [
    # n, ci, co, h, w, k, stride, dil, pad, pad_mode, act
    (2, 3, 128, 16, 16, 3, 1, 1, 1, "zero", _lib.ACT_RELU),     # SPADE mlp_shared
    (2, 128, 40, 16, 20, 3, 1, 1, 1, "zero", _lib.ACT_NONE),    # SPADE gamma/beta (N=40)
    (1, 40, 20, 12, 12, 3, 1, 1, 1, "zero", _lib.ACT_NONE),     # SN conv, odd channels
    (2, 40, 20, 8, 8, 1, 1, 1, 0, "zero", _lib.ACT_NONE),       # conv_s 1x1
    (2, 20, 3, 16, 16, 3, 1, 1, 1, "zero", _lib.ACT_TANH),      # conv_img
    (1, 16, 24, 17, 13, 3, 1, 1, 1, "reflect", _lib.ACT_LRELU), # reflect pad, ragged size
    (2, 8, 16, 16, 16, 4, 2, 1, 1, "zero", _lib.ACT_LRELU),     # discriminator 4x4 s2
    (1, 64, 32, 20, 20, 3, 1, 6, 6, "zero", _lib.ACT_NONE),     # ASPP atrous d6
    (1, 8, 16, 23, 23, 7, 2, 1, 3, "zero", _lib.ACT_RELU),      # stem 7x7 s2
    (3, 128, 256, 5, 5, 3, 1, 1, 1, "zero", _lib.ACT_NONE),     # tiny spatial, larger channels
    # masker training path (ResNet-101 encoder / ASPP / decoders)
    (2, 3, 64, 64, 64, 7, 2, 1, 3, "zero", _lib.ACT_NONE),      # stem 7x7 s2, 3 real input channels
    (2, 64, 128, 32, 32, 1, 2, 1, 0, "zero", _lib.ACT_NONE),    # layer2 1x1 stride 2 (conv1 / downsample)
    (2, 128, 128, 16, 16, 3, 1, 2, 2, "zero", _lib.ACT_NONE),   # layer3 3x3 dilation 2
    (2, 64, 64, 16, 16, 3, 1, 4, 4, "zero", _lib.ACT_NONE),     # layer4 3x3 dilation 4
    (2, 256, 64, 16, 16, 3, 1, 12, 12, "zero", _lib.ACT_NONE),  # ASPP d12: most taps fall in the padding
    (2, 256, 64, 16, 16, 3, 1, 18, 18, "zero", _lib.ACT_NONE),  # ASPP d18 >= map size: only the centre tap is live
    (2, 512, 128, 1, 1, 1, 1, 1, 0, "zero", _lib.ACT_RELU),     # ASPP image-pool branch: 1x1 conv on a 1x1 map
    (2, 64, 64, 18, 18, 3, 1, 1, 0, "zero", _lib.ACT_LRELU),    # pad-0 3x3 after an explicit reflect pad
    (2, 2048, 64, 8, 8, 1, 1, 1, 0, "zero", _lib.ACT_LRELU),    # mask decoder proj_conv, K = 2048
    (1, 256, 1024, 40, 40, 1, 1, 1, 0, "zero", _lib.ACT_NONE),  # ResNet conv3 1x1 on a map large enough for the resident-weight path
    # ResNet conv1 1x1 (the ReLU follows the BatchNorm, not the conv; with a fused gate 4e5 outputs at K=1024 always hold one
    # pre-activation within fp32 accumulation error of 0, whose flipped gate is an O(gy*w) dgrad outlier against fp64)
    (1, 1024, 256, 40, 40, 1, 1, 1, 0, "zero", _lib.ACT_NONE),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("engine", ENGINES)
def test_conv_fwd_bwd(cuda, case, dtype, engine):
    n, ci, co, h, w, k, stride, dil, pad, pad_mode, act = case
    pm = _lib.PAD_REFLECT if pad_mode == "reflect" else _lib.PAD_ZERO
    slope = 0.2
    for attempt in range(16):
        # deterministic (hash() of a tuple with str is salted per process)
        torch.manual_seed(CONV_CASES.index(case) * 7 + 1 + 1000 * attempt)
        x = _q(torch.randn(n, ci, h, w), dtype)
        wt = _q(torch.randn(co, ci, k, k) / (ci * k * k) ** 0.5, dtype)
        b = torch.randn(co) * 0.1

        # reference (fp64 CPU)
        xr = x.double().requires_grad_(True)
        wr = wt.double().requires_grad_(True)
        br = b.double().requires_grad_(True)
        xp = F.pad(xr, (pad,) * 4, mode="reflect") if pad_mode == "reflect" else F.pad(xr, (pad,) * 4)
        yr = F.conv2d(xp, wr, br, stride=stride, dilation=dil)
        # a pre-activation within fp32 accumulation error of 0 flips the (l)relu gate against the fp64 reference and
        # shows up as an O(gy*w) outlier in dgrad: redraw until there is none (the fused-gate cases have <= 3e4 outputs)
        if act not in (_lib.ACT_RELU, _lib.ACT_LRELU) or float(yr.detach().abs().min()) > 1e-5:
            break
    else:
        pytest.fail("no seed with a gate margin")
    yr = {_lib.ACT_NONE: lambda t: t, _lib.ACT_RELU: F.relu, _lib.ACT_LRELU: lambda t: F.leaky_relu(t, slope),
          _lib.ACT_TANH: torch.tanh}[act](yr)
    gy = _q(torch.randn_like(yr).float(), dtype)
    yr.backward(gy.double())

    xs = _st(x, dtype, cuda).requires_grad_(True)
    wg = wt.to(cuda).requires_grad_(True)
    bg = b.to(cuda).requires_grad_(True)
    y = ops.conv2d(xs, wg, bg, stride=stride, dil=dil, pad=pad, pad_mode=pm, act=act, slope=slope, engine=engine)
    y_nchw = ops.from_storage(y, co)
    tol = _tol(dtype)
    if dtype == torch.float32 and ci * k * k >= 1024:
        tol *= 4   # fp32 accumulation over K >= 1024 terms against the fp64 reference
    assert rel_max(y_nchw, yr) < tol
    if y.shape[-1] > co:  # pad channels stay exactly zero
        assert float(y[..., co:].abs().max()) == 0.0
    if pad_mode == "reflect":
        # dgrad with reflect padding is not built: weight/bias grads only
        xs2 = xs.detach()
        y2 = ops.from_storage(ops.conv2d(xs2, wg, bg, stride=stride, dil=dil, pad=pad, pad_mode=pm, act=act,
                                         slope=slope, engine=engine), co)
        y2.backward(gy.to(cuda))
    else:
        y_nchw.backward(gy.to(cuda))
        gx = ops.from_storage(xs.grad, ci)
        assert rel_max(gx, xr.grad) < tol, "dgrad"
    assert rel_max(wg.grad, wr.grad) < tol, "wgrad"
    assert rel_max(bg.grad, br.grad) < tol, "bias grad"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_residual(cuda, dtype):
    torch.manual_seed(5)
    x = _q(torch.randn(2, 16, 9, 9), dtype)
    r = _q(torch.randn(2, 24, 9, 9), dtype)
    wt = _q(torch.randn(24, 16, 3, 3) * 0.1, dtype)
    ref = F.conv2d(x, wt, None, padding=1) + r
    y = ops.conv2d(_st(x, dtype, cuda), wt.to(cuda), None, _st(r, dtype, cuda), pad=1)
    assert rel_max(ops.from_storage(y, 24), ref) < _tol(dtype)


@pytest.mark.parametrize("c,h,w", [(20, 16, 16), (40, 7, 9), (640, 5, 5), (8, 64, 64)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_instnorm_stats(cuda, c, h, w, dtype):
    torch.manual_seed(c)
    x = _q(torch.randn(3, c, h, w) * 2 + 1.5, dtype)
    mean, rstd = ops.instnorm_stats(_st(x, dtype, cuda))
    mr = x.double().mean((2, 3))
    vr = x.double().var((2, 3), unbiased=False)
    assert rel_max(mean[:, :c], mr) < 1e-5
    assert rel_max(rstd[:, :c], 1 / torch.sqrt(vr + 1e-5)) < 1e-5
    assert float(mean[:, c:].abs().max() if mean.shape[1] > c else 0) == 0.0


@pytest.mark.parametrize("c,h,w", [(20, 12, 12), (40, 6, 10), (128, 4, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [_lib.ACT_NONE, _lib.ACT_LRELU])
@pytest.mark.parametrize("col", [False, True])
def test_spade_layer(cuda, c, h, w, dtype, act, col):
    """One whole SPADE layer (norms.py:174-186) + lrelu, forward and every gradient."""
    from oracle import painter_oracle as po

    torch.manual_seed(7 * c + h)
    n = 2
    x = _q(torch.randn(n, c, h, w) * 1.5 + 0.3, dtype)
    seg = _q(torch.rand(n, 3, h, w) * 2 - 1, dtype)
    sd = {
        "p.mlp_shared.0.weight": torch.randn(128, 3, 3, 3) * 0.3, "p.mlp_shared.0.bias": torch.randn(128) * 0.1,
        "p.mlp_gamma.weight": torch.randn(c, 128, 3, 3) * 0.03, "p.mlp_gamma.bias": torch.randn(c) * 0.1,
        "p.mlp_beta.weight": torch.randn(c, 128, 3, 3) * 0.03, "p.mlp_beta.bias": torch.randn(c) * 0.1,
    }
    sd = {k: _q(v, dtype) if k.endswith("weight") else v for k, v in sd.items()}
    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xr = x.double().requires_grad_(True)
    out_r = po.spade(sdr, "p", xr, seg.double())
    if act == _lib.ACT_LRELU:
        out_r = F.leaky_relu(out_r, 0.2)
    go = _q(torch.randn_like(out_r).float(), dtype)
    out_r.backward(go.double())

    sdg = {k: v.to(cuda).requires_grad_(True) for k, v in sd.items()}
    xs = _st(x, dtype, cuda).requires_grad_(True)
    segs = _st(seg, dtype, cuda)
    mean, rstd = ops.instnorm_stats(xs)
    if col:  # mlp_shared as one K=32 GEMM over im2col patches of the conditioning
        segs = ops.im2col(segs, 3, 3, 1)
        assert segs.shape[-1] == 32
    out = ops.spade(xs, mean, rstd, segs, sdg["p.mlp_shared.0.weight"], sdg["p.mlp_shared.0.bias"],
                    sdg["p.mlp_gamma.weight"], sdg["p.mlp_gamma.bias"], sdg["p.mlp_beta.weight"],
                    sdg["p.mlp_beta.bias"], act, 0.2, seg_is_col=col)
    o = ops.from_storage(out, c)
    # Stated tolerances.  fp32 storage: 5e-5 of full scale everywhere.  bf16 storage: forward 1e-2 of full
    # scale; gradients are compared by cosine / relative L2 because a leaky-relu whose pre-activation flips
    # sign under bf16 rounding changes that single element's gradient by 5x (measured: cos 0.9996, L2 3e-2).
    o.backward(go.to(cuda))
    gx = ops.from_storage(xs.grad, c)
    if dtype == torch.float32:
        assert rel_max(o, out_r) < 5e-5
        assert rel_max(gx, xr.grad) < 5e-5, "gx"
    else:
        assert rel_max(o, out_r) < 1e-2
        assert cosine(gx, xr.grad) > 0.998 and rel_l2(gx, xr.grad) < 6e-2, "gx"
    for k in sd:
        g, gr = sdg[k].grad, sdr[k].grad
        if dtype == torch.float32:
            assert rel_max(g, gr) < 2e-4, k
        else:
            assert cosine(g, gr) > 0.998 and rel_l2(g, gr) < 6e-2, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_spade_layer_batch_stats_and_cond_grad(cuda, dtype):
    """SPADE with the BatchNorm param-free norm in train mode (norms.py:154-155, the masker's MaskSpadeDecoder) on a
    15-channel differentiable conditioning tensor: forward, gradient into x (through the batch statistics), into the
    conditioning (dgrad of mlp_shared) and into every weight, against plain PyTorch fp64; running statistics as
    F.batch_norm updates them."""
    torch.manual_seed(23)
    n, c, h, w, cn = 3, 24, 12, 10, 15
    x = _q(torch.randn(n, c, h, w) * 1.5 + 0.3, dtype)
    seg = _q(torch.rand(n, cn, h, w) * 2 - 1, dtype)
    sd = {"sh.w": torch.randn(128, cn, 3, 3) * 0.1, "sh.b": torch.randn(128) * 0.1,
          "g.w": torch.randn(c, 128, 3, 3) * 0.03, "g.b": torch.randn(c) * 0.1,
          "b.w": torch.randn(c, 128, 3, 3) * 0.03, "b.b": torch.randn(c) * 0.1}
    sd = {k: _q(v, dtype) if k.endswith("w") else v for k, v in sd.items()}
    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xr, segr = x.double().requires_grad_(True), seg.double().requires_grad_(True)
    rm, rv = torch.zeros(c, dtype=torch.float64), torch.ones(c, dtype=torch.float64)
    xn = F.batch_norm(xr, rm, rv, None, None, True, 0.1, 1e-5)
    actv = F.relu(F.conv2d(segr, sdr["sh.w"], sdr["sh.b"], padding=1))
    out_r = F.leaky_relu(xn * (1 + F.conv2d(actv, sdr["g.w"], sdr["g.b"], padding=1))
                         + F.conv2d(actv, sdr["b.w"], sdr["b.b"], padding=1), 0.2)
    go = _q(torch.randn_like(out_r).float(), dtype)
    out_r.backward(go.double())

    bn = torch.nn.BatchNorm2d(c, affine=False).to(cuda).train()
    sdg = {k: v.to(cuda).requires_grad_(True) for k, v in sd.items()}
    xs = _st(x, dtype, cuda).requires_grad_(True)
    segs = _st(seg, dtype, cuda).requires_grad_(True)
    mean, rstd = ops.batchnorm_stats_update(xs.detach(), bn)
    assert mean.shape == (1, xs.shape[-1])
    out = ops.spade(xs, mean, rstd, segs, sdg["sh.w"], sdg["sh.b"], sdg["g.w"], sdg["g.b"], sdg["b.w"], sdg["b.b"],
                    _lib.ACT_LRELU, 0.2, batch_stats=True)
    o = ops.from_storage(out, c)
    o.backward(go.to(cuda))
    gx, gseg = ops.from_storage(xs.grad, c), ops.from_storage(segs.grad, cn)
    assert rel_max(bn.running_mean, rm) < (1e-5 if dtype == torch.float32 else 1e-5)
    assert rel_max(bn.running_var, rv) < 1e-5
    assert int(bn.num_batches_tracked) == 1
    if dtype == torch.float32:
        assert rel_max(o, out_r) < 5e-5
        assert rel_max(gx, xr.grad) < 5e-5, "gx"
        assert rel_max(gseg, segr.grad) < 5e-5, "gseg"
    else:
        assert rel_max(o, out_r) < 1e-2
        assert cosine(gx, xr.grad) > 0.998 and rel_l2(gx, xr.grad) < 6e-2, "gx"
        assert cosine(gseg, segr.grad) > 0.998 and rel_l2(gseg, segr.grad) < 6e-2, "gseg"
    for k in sd:
        g, gr = sdg[k].grad, sdr[k].grad
        if dtype == torch.float32:
            assert rel_max(g, gr) < 2e-4, k
        else:
            assert cosine(g, gr) > 0.998 and rel_l2(g, gr) < 6e-2, k


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_make_m_cond_fwd_bwd(cuda, dtype):
    """OmniGenerator.make_m_cond (generator.py:196-230) as a differentiable op: cat[normalize(d), softmax(s), x] and its
    adjoint w.r.t. d (through the per-sample min / max, tutils.py:566-577) and s, against plain PyTorch fp64."""
    torch.manual_seed(31)
    n, ns, h, w = 3, 11, 9, 13
    d = _q(torch.randn(n, 1, h, w), dtype)
    s = _q(torch.randn(n, ns, h, w) * 2, dtype)
    xr_ = _q(torch.rand(n, 3, h, w) * 2 - 1, dtype)
    dr, sr = d.double().requires_grad_(True), s.double().requires_grad_(True)
    mn = dr.reshape(n, -1).min(1)[0].reshape(n, 1, 1, 1)
    t = dr - mn
    t = t / t.reshape(n, -1).max(1)[0].reshape(n, 1, 1, 1)
    ref = torch.cat([t, torch.softmax(sr, 1), xr_.double()], 1)
    go = _q(torch.randn_like(ref).float(), dtype)
    ref.backward(go.double())

    ds = _st(d, dtype, cuda).requires_grad_(True)
    ss = _st(s, dtype, cuda).requires_grad_(True)
    out = ops.make_m_cond(ds, ss, _st(xr_, dtype, cuda), ns)
    assert out.shape == (n, h, w, 16)
    o = ops.from_storage(out, 15)
    o.backward(go.to(cuda))
    gd, gs = ops.from_storage(ds.grad, 1), ops.from_storage(ss.grad, ns)
    if dtype == torch.float32:
        assert rel_max(o, ref) < 1e-5
        assert rel_max(gd, dr.grad) < 2e-5, "gd"
        assert rel_max(gs, sr.grad) < 2e-5, "gs"
    else:
        assert rel_max(o, ref) < 1e-2
        assert cosine(gd, dr.grad) > 0.999 and rel_l2(gd, dr.grad) < 3e-2, "gd"
        assert cosine(gs, sr.grad) > 0.999 and rel_l2(gs, sr.grad) < 3e-2, "gs"
    assert float(ds.grad[..., 1:].abs().max()) == 0.0 and float(ss.grad[..., ns:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resize_and_layout(cuda, dtype):
    torch.manual_seed(0)
    x = _q(torch.randn(2, 20, 10, 10), dtype)
    xs = _st(x, dtype, cuda).requires_grad_(True)
    assert xs.shape == (2, 10, 10, 24)
    assert torch.equal(ops.from_storage(xs, 20).cpu(), x)
    up = ops.upsample2x(xs)
    assert torch.equal(ops.from_storage(up, 20).cpu(), F.interpolate(x, scale_factor=2, mode="nearest"))
    g = _q(torch.randn(2, 20, 20, 20), dtype)
    ops.from_storage(up, 20).backward(g.to(cuda))
    gr = F.avg_pool2d(g, 2) * 4
    assert rel_max(ops.from_storage(xs.grad, 20), gr) < (1e-6 if dtype == torch.float32 else 1e-2)
    # arbitrary nearest down-size, as SPADE does to the conditioning (norms.py:179) and painter.py:152
    for size in [(5, 5), (7, 3), (10, 10), (4, 9)]:
        y = ops.resize_nearest(xs.detach(), *size)
        assert torch.equal(ops.from_storage(y, 20).cpu(), F.interpolate(x, size=size, mode="nearest")), size
    # ... and its adjoint for any ratio (down-sizing, non-integer up-sizing), against autograd of F.interpolate
    for size in [(5, 5), (7, 3), (4, 9), (13, 17), (15, 10), (30, 25)]:
        xa = xs.detach().clone().requires_grad_(True)
        gg = _q(torch.randn(2, 20, *size), dtype)
        ops.from_storage(ops.resize_nearest(xa, *size), 20).backward(gg.to(cuda))
        xr = x.double().requires_grad_(True)
        F.interpolate(xr, size=size, mode="nearest").backward(gg.double())
        assert rel_max(ops.from_storage(xa.grad, 20), xr.grad) < (1e-6 if dtype == torch.float32 else 1e-2), size


def test_spectral_power_iter(cuda):
    from oracle import painter_oracle as po

    torch.manual_seed(11)
    for shape in [(24, 16, 3, 3), (640, 64, 3, 3), (20, 40, 1, 1)]:
        w = torch.randn(*shape)
        u = po.l2normalize(torch.randn(shape[0]))
        v = po.l2normalize(torch.randn(w[0].numel()))
        wr = w.clone().requires_grad_(True)
        w_ref, u_ref, v_ref = po.spectral_norm_weight(wr, u, v)
        g = torch.randn_like(w)
        w_ref.backward(g)
        wg = w.to(cuda).requires_grad_(True)
        ug, vg = u.to(cuda), v.to(cuda)
        w_eff = ops.spectral_weight(wg, ug, vg)
        w_eff.backward(g.to(cuda))
        assert rel_max(w_eff, w_ref) < 1e-5
        assert rel_max(ug, u_ref) < 1e-5 and rel_max(vg, v_ref) < 1e-5  # mutated in place
        assert rel_max(wg.grad, wr.grad) < 1e-4


def test_paste_and_l1(cuda):
    torch.manual_seed(2)
    x = torch.rand(2, 3, 16, 16) * 2 - 1
    m = (torch.rand(2, 1, 16, 16) > 0.5).float()
    f = (torch.rand(2, 3, 16, 16) * 2 - 1).requires_grad_(True)
    t = torch.rand(2, 3, 16, 16)
    ref = x * (1 - m) + f * m
    lr = F.l1_loss(ref, t)
    lr.backward()
    fg = f.detach().to(cuda).requires_grad_(True)
    out = ops.paste(x.to(cuda), m.to(cuda), fg)
    loss = ops.l1_loss(out, t.to(cuda))
    loss.backward()
    assert rel_max(out, ref) < 1e-6 and abs(float(loss) - float(lr)) < 1e-6
    assert rel_max(fg.grad, f.grad) < 1e-6
    cond = ops.mask_cond(x.to(cuda), m.to(cuda), torch.float32)
    assert rel_max(ops.from_storage(cond, 3), x * (1 - m)) < 1e-7


def test_bad_arguments_raise(cuda):
    x = torch.zeros(1, 8, 8, 12, device=cuda)  # 12 channels: not a storage tensor
    with pytest.raises(ValueError):
        ops.instnorm_stats(x)
    xs = torch.zeros(1, 8, 8, 8, device=cuda)
    with pytest.raises(_lib.CgbError):  # residual + activation is rejected by the C ABI
        ops.conv2d(xs, torch.zeros(8, 8, 3, 3, device=cuda), None, torch.zeros(1, 8, 8, 8, device=cuda), pad=1,
                   act=_lib.ACT_RELU)
    with pytest.raises(_lib.CgbError):  # reflect pad >= size
        ops.conv2d(torch.zeros(1, 2, 2, 8, device=cuda), torch.zeros(8, 8, 5, 5, device=cuda), pad=2,
                   pad_mode=_lib.PAD_REFLECT)
