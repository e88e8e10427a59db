"""GPU tests of the smaller entry points off the default train path (reverse-Huber depth loss, one-launch weight packing,
argmax-confusion metrics, diff-augment: green on B200 since round 1's driver run; per-(sample, channel) affine, moments,
replicate padding: round 2), each against plain PyTorch.  The file sorts last on purpose: the suite is run with -x, and a
first-run failure of a new kernel here must not hide the validated tests before it.  Each kernel is also pinned on CPU through
the emulated ABI (tests/test_emulated.py, tests/test_conv2dblock_options.py)."""
import pytest
import torch

from climategan_b200 import _lib, ops
from tests.helpers import rel_max

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.mark.parametrize("n,h,w", [(2, 16, 12), (3, 33, 47)])
def test_dada_depth_loss(cuda, n, h, w):
    """DADADepthLoss (losses.py:596-620; gen.d.loss = "dada"): reverse Huber with the batch-wide threshold 0.2 * max|pred - label|
    (a constant of the graph), value and gradient against plain PyTorch."""
    pred = _rand(n, 1, h, w, seed=1)
    targ = _rand(n, 1, h, w, seed=2).abs()
    pr = pred.clone().requires_grad_()
    adiff = torch.abs(pr - targ)
    c = 0.2 * float(adiff.max())
    ref = ((adiff * (adiff <= c).float()).sum() + ((adiff * adiff + c * c) / (2 * c) * (adiff > c).float()).sum()) / adiff.numel()
    ref.backward()
    pd = pred.to(cuda).requires_grad_()
    out = ops.dada_depth_loss(pd, targ.to(cuda))
    out.backward()
    assert abs(float(out) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_max(pd.grad, pr.grad) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pack_weight_kernel(cuda, dtype):
    """cgb_pack_weight (one launch) against the torch packing (zeros + permute + slice copy): bit-exact, channel padding zero."""
    torch.manual_seed(3)
    for (o, i, k, cis, cos) in [(20, 40, 3, None, None), (128, 3, 3, 8, None), (1, 16, 3, None, 8), (64, 4, 4, 8, None),
                                (1024, 256, 1, None, None), (11, 256, 1, 256, 16), (48, 128, 3, 128, 96)]:
        w = torch.randn(o, i, k, k, device=cuda)
        ref = ops.pack_weight(w, dtype, cis=cis, cos=cos, kernel=False)
        got = ops.pack_weight(w, dtype, cis=cis, cos=cos, kernel=True)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        assert torch.equal(got, ref), (o, i, k, cis, cos)
    wt = torch.randn(8, 24, 3, 3, device=cuda).transpose(0, 1)   # non-contiguous input
    assert torch.equal(ops.pack_weight(wt, dtype, kernel=True), ops.pack_weight(wt, dtype, kernel=False))


def test_eval_metrics_match_reference_fixture_and_oracle(cuda):
    """cgb_argmax_confusion -> accuracy / mIOU (eval_metrics.py:68-124; Trainer.eval_images): the fixture generated from the
    reference's own functions (ties, absent classes, an ignore index, two-class masks, a ragged size), then the oracle at the full
    640 x 640 / 11 classes / batch 8 with NaN logits, and the size-independent property sum(conf) = pixels."""
    import numpy as np

    from climategan_b200 import eval_metrics as em
    from oracle import eval_metrics_oracle as o
    from tests.golden.eval_cases import cases
    from tests.test_eval_metrics import check_case

    for name, (pred, label, kind) in cases().items():
        check_case(name, pred.to(cuda), label.to(cuda), kind, em.accuracy, em.mIOU)
    rs = np.random.RandomState(11)
    pred = rs.standard_normal((8, 11, 640, 640)).astype(np.float32)
    pred[rs.random_sample(pred.shape) < 1e-4] = np.nan          # a NaN is the maximum (torch.argmax / np.argmax)
    label = rs.randint(0, 12, size=(8, 1, 640, 640)).astype(np.int64)    # 11 = out of range
    conf, lmax = em.confusion(torch.from_numpy(pred).to(cuda), torch.from_numpy(label).to(cuda))
    p = np.argmax(pred, axis=1).reshape(-1)
    want = np.bincount(p * 12 + label.reshape(-1), minlength=11 * 12).reshape(11, 12)
    assert conf.sum() == label.size and lmax == 11
    assert np.array_equal(conf, want)
    assert em.mIOU(torch.from_numpy(pred).to(cuda), torch.from_numpy(label).to(cuda)) == pytest.approx(o.miou(pred, label), rel=1e-14)


@pytest.mark.parametrize("shape,cut,shift", [((3, 3, 24, 40), (12, 20), (3, 5)), ((2, 3, 17, 23), (5, 7), (5, 7)),
                                             ((2, 3, 640, 640), (320, 320), (80, 80)), ((2, 4, 16, 16), (0, 0), (8, 8))])
def test_diff_aug_kernels(cuda, shape, cut, shift):
    """cgb_diff_aug_sum / _fwd / _bwd (DiffTransforms, transforms.py:493-626) against the plain-PyTorch statement of the same op
    (tests/helpers.torch_diff_aug, itself pinned to the reference's DiffTransforms in tests/test_diff_aug.py) on the same draws:
    value and gradient, translations and cutout boxes that leave the image included."""
    from tests.helpers import torch_diff_aug

    n, c, h, w = shape
    g = torch.Generator().manual_seed(h * w + n)
    x = torch.randn(*shape, generator=g)
    wgt = torch.randn(*shape, generator=g)
    p = torch.zeros(n, 8)
    p[:, 0] = torch.rand(n, generator=g) - 0.5
    p[:, 1] = torch.rand(n, generator=g) + 0.5
    p[:, 2] = torch.rand(n, generator=g) * 2
    p[:, 3] = torch.randint(-shift[0], shift[0] + 1, (n,), generator=g).float()
    p[:, 4] = torch.randint(-shift[1], shift[1] + 1, (n,), generator=g).float()
    p[:, 5] = torch.randint(0, h + 1, (n,), generator=g).float()
    p[:, 6] = torch.randint(0, w + 1, (n,), generator=g).float()
    p[0, 3:7] = torch.tensor([shift[0], -shift[1], 0, w])        # extremes: largest shifts, box clamped at two corners
    xr = x.clone().requires_grad_()
    want = torch_diff_aug(xr, p, *cut)
    (want * wgt).sum().backward()
    xd = x.to(cuda).requires_grad_()
    got = ops.diff_aug(xd, p.to(cuda), *cut)
    (got * wgt.to(cuda)).sum().backward()
    assert rel_max(got, want) < 5e-6     # fp32 rounding only (fp64 sample mean here, ATen's fp32 cascade sum in the reference)
    assert rel_max(xd.grad, xr.grad) < 2e-5


def test_trainer_eval_images(cuda):
    """Trainer.eval_images (trainer.py:1706-1799) on the device against the oracle metrics of the trainer's own predictions."""
    from tests.test_eval_metrics import check_trainer_eval_images

    check_trainer_eval_images(cuda)


# ---- round 2: per-(sample, channel) affine + activation, replicate padding, SELU (Conv2dBlock options off the default path)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [_lib.ACT_NONE, _lib.ACT_RELU, _lib.ACT_LRELU, _lib.ACT_TANH, _lib.ACT_SELU])
def test_affine_nc_fwd_bwd(cuda, dtype, act):
    import torch.nn.functional as F

    torch.manual_seed(act + 1)
    n, c, h, w = 3, 24, 19, 23
    q = lambda t: t.to(dtype).float()  # noqa: E731
    x = q(torch.randn(n, c, h, w))
    sc, sh = torch.randn(n, c) * 0.5 + 1, torch.randn(n, c)
    gy = q(torch.randn(n, c, h, w))
    fn = {_lib.ACT_NONE: lambda t: t, _lib.ACT_RELU: F.relu, _lib.ACT_LRELU: lambda t: F.leaky_relu(t, 0.2), _lib.ACT_TANH: torch.tanh,
          _lib.ACT_SELU: F.selu}[act]
    xr, scr, shr = x.double().requires_grad_(True), sc.double().requires_grad_(True), sh.double().requires_grad_(True)
    yr = fn(xr * scr.view(n, c, 1, 1) + shr.view(n, c, 1, 1))
    yr.backward(gy.double())
    xs = ops.to_storage(x.to(cuda), dtype).requires_grad_(True)
    scg, shg = sc.to(cuda).requires_grad_(True), sh.to(cuda).requires_grad_(True)
    y = ops.from_storage(ops.affine_nc(xs, scg, shg, act, 0.2), c)
    y.backward(gy.to(cuda))
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    assert rel_max(y, yr) < tol
    assert rel_max(ops.from_storage(xs.grad, c), xr.grad) < tol
    assert rel_max(scg.grad, scr.grad) < tol and rel_max(shg.grad, shr.grad) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_moments_and_replicate_pad(cuda, dtype):
    import torch.nn.functional as F

    torch.manual_seed(9)
    n, c, h, w, pad = 2, 16, 13, 10, 2
    x = torch.randn(n, c, h, w).to(dtype).float()
    xr = x.double().requires_grad_(True)
    m1r, m2r = xr.mean((2, 3)), (xr * xr).mean((2, 3))
    g1, g2 = torch.randn(n, c), torch.randn(n, c)
    (m1r * g1.double()).sum().backward(retain_graph=True)
    (m2r * g2.double()).sum().backward()
    xs = ops.to_storage(x.to(cuda), dtype).requires_grad_(True)
    m1, m2 = ops.moments(xs)
    ((m1 * g1.to(cuda)).sum() + (m2 * g2.to(cuda)).sum()).backward()
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    assert rel_max(m1, m1r) < 1e-5 and rel_max(m2, m2r) < 1e-5
    assert rel_max(ops.from_storage(xs.grad, c), xr.grad) < tol
    # replicate pad forward / adjoint
    xr2 = x.double().requires_grad_(True)
    yr = F.pad(xr2, (pad,) * 4, mode="replicate")
    gy = torch.randn_like(yr).to(dtype).double()
    yr.backward(gy)
    xs2 = ops.to_storage(x.to(cuda), dtype).requires_grad_(True)
    y = ops.from_storage(ops.replicate_pad(xs2, pad), c)
    assert rel_max(y, yr) == 0.0
    y.backward(gy.float().to(cuda))
    assert rel_max(ops.from_storage(xs2.grad, c), xr2.grad) < tol


@pytest.mark.parametrize("shape,target", [((480, 853), 128), ((1200, 800), 256), ((300, 300), 128), ((97, 160), 128)])
def test_input_edge_resize_and_crop(cuda, shape, target):
    """events.InputEdge (cgb_resize_crop_u8): uint8 HWC photo -> anti-aliased bilinear resize (short side = target) -> centre crop
    -> [-1, 1], against F.interpolate(mode="bilinear", antialias=True) + the reference's crop arithmetic (apply_events.py:223-241).
    Up- and down-scaling, both orientations."""
    import torch.nn.functional as F

    from climategan_b200.events import InputEdge, resize_and_crop_size

    h, w = shape
    g = torch.Generator().manual_seed(h + w)
    img = torch.randint(0, 256, (h, w, 3), generator=g, dtype=torch.uint8)
    rh, rw, top, left = resize_and_crop_size(h, w, target)
    ref = F.interpolate(img.permute(2, 0, 1)[None].float(), size=(rh, rw), mode="bilinear", antialias=True, align_corners=False)
    ref = ref[:, :, top:top + target, left:left + target]
    out = InputEdge(cuda, quantize=False)([img.numpy(), img.numpy()], target)
    assert tuple(out.shape) == (2, 3, target, target) and torch.equal(out[0], out[1])
    assert float((out[0].cpu() - (ref[0] / 255 - 0.5) * 2).abs().max()) < 2e-4
    outq = InputEdge(cuda, quantize=True)([img.numpy()], target)       # the reference truncates the resized image to uint8 (:231)
    want = (torch.floor(ref[0].clamp(0, 255)) / 255 - 0.5) * 2
    d = (outq[0].cpu() - want).abs()
    assert float((d > 1e-4).float().mean()) < 2e-3 and float(d.max()) <= 2.0 / 255 + 1e-4   # (floor flips where fp32 sums straddle an integer)


def test_device_prefetcher_overlaps_and_orders_copies(cuda):
    """climategan_b200.data.DevicePrefetcher on real streams: pinned H2D on a side stream, batches arrive intact and in order,
    the consumer only waits on the copy's event."""
    from climategan_b200.data import DevicePrefetcher

    host = [{"data": {"x": torch.full((4, 3, 64, 64), float(i)).pin_memory(), "s": torch.full((4, 1, 16, 16), i, dtype=torch.int64)},
             "domain": ["r"] * 4} for i in range(6)]
    seen = []
    for b in DevicePrefetcher(((h,) for h in host), cuda, depth=2):
        x = b[0]["data"]["x"]
        assert x.is_cuda and b[0]["data"]["s"].dtype == torch.int64 and b[0]["domain"] == ["r"] * 4
        seen.append(float((x * 2).mean()) / 2)      # a kernel on the compute stream consuming the prefetched tensor
    assert seen == [float(i) for i in range(6)]


def test_pack_weight_dual_matches_the_two_separate_packings(cuda):
    """cgb_pack_weight_dual: the forward packing and the dgrad packing in one launch, bit-exact against cgb_pack_weight +
    cgb_conv2d_pack_dgrad_weight, with channel padding on both sides."""
    torch.manual_seed(2)
    for (o, i, k) in [(20, 40, 3), (256, 64, 1), (11, 3, 7)]:
        w = torch.randn(o, i, k, k, device=cuda)
        for dt in (torch.bfloat16, torch.float16):
            a = ops.pack_weight(w, dt, with_dgrad=True)
            b = ops.pack_weight(w, dt)
            assert torch.equal(a, b)
            g = ops.ConvGeom(k, k, 1, 1, k // 2)
            gy = torch.zeros(1, 8, 8, b.shape[0], dtype=dt, device=cuda)
            ops.conv_dgrad_raw(gy, b, (1, 8, 8, b.shape[2]), g)       # builds b._cgb_wt the old way
            assert torch.equal(a._cgb_wt, b._cgb_wt)
