"""Norm statistics fused into the tcgen05 conv epilogue (cgb_conv2d_fwd_stats -> cgb_bn_train_fwd_partials) and the flat
train-mode BatchNorm passes, on the GPU, against (i) the separate statistics pass over the stored tensor and (ii) plain PyTorch
fp64 — including the BENCHMARKED shapes (8 x 80 x 80 ResNet maps, 256 / 1024 channels), which round 1 never exercised."""
import pytest
import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.helpers import rel_max

pytestmark = pytest.mark.gpu

# n, ci, co, h, w, k, stride, dil, pad   — every epilogue variant: TMA-store staging (N a multiple of 64, or one N tile), padded
# staging (strided dgrad-free cases with N % 64 != 0 and several N tiles), the weight-stationary halo kernel (co <= 64 on a
# large map), several N tiles (co > 256), ragged tile edges, tiles that span images
STATS_CASES = [
    (8, 256, 256, 80, 80, 3, 1, 2, 2),     # BENCH class: ResNet layer3 conv2 (one N tile of 256, TMA store)
    (8, 256, 1024, 80, 80, 1, 1, 1, 0),    # BENCH class: conv3 1x1, four N tiles
    (8, 1024, 256, 80, 80, 1, 1, 1, 0),    # BENCH class: conv1 1x1
    (2, 152, 64, 320, 320, 1, 1, 1, 0),    # the stem as an im2col GEMM (K = 152), weight-stationary-eligible map
    (2, 64, 64, 160, 160, 3, 1, 1, 1),     # layer1 conv2 on a large map: weight-stationary halo kernel
    (3, 128, 320, 37, 29, 3, 1, 1, 1),     # ragged edges, two N tiles of 160 (padded staging)
    (5, 64, 72, 9, 7, 1, 1, 1, 0),         # tiles span several images, batch tail, N = 72
    (2, 64, 128, 32, 32, 1, 2, 1, 0),      # stride 2
    (2, 512, 512, 40, 40, 3, 1, 4, 4),     # dilation 4
]


@pytest.mark.parametrize("case", STATS_CASES)
def test_conv_epilogue_stats_match_a_statistics_pass(cuda, case):
    n, ci, co, h, w, k, stride, dil, pad = case
    torch.manual_seed(STATS_CASES.index(case) + 3)
    x = ops.to_storage((torch.randn(n, ci, h, w, device=cuda) * 1.5 + 0.3), torch.bfloat16)
    wt = (torch.randn(co, ci, k, k, device=cuda) / (ci * k * k) ** 0.5)
    y_plain = ops.conv2d(x, wt, None, stride=stride, dil=dil, pad=pad)
    y, partial = ops.conv2d(x, wt, None, stride=stride, dil=dil, pad=pad, want_stats=True)
    assert partial is not None and partial.shape == (_lib.lib().cgb_conv2d_stats_rows(), 2, y.shape[-1])
    assert torch.equal(y, y_plain)                       # the statistics do not disturb the output
    npix = y.shape[0] * y.shape[1] * y.shape[2]
    tot = partial.double().sum(0)                        # [2, co]
    yd = y.double().reshape(npix, -1)
    # sums of the STORED values: exact up to fp32 accumulation inside a CTA (fp64 across CTAs)
    assert rel_max(tot[0], yd.sum(0)) < 2e-5 * max(1.0, float(npix) ** 0.5 / 30)
    assert rel_max(tot[1], (yd * yd).sum(0)) < 2e-5
    mean_ref, rstd_ref = ops.instnorm_stats(y.view(1, y.shape[0] * y.shape[1], y.shape[2], y.shape[3]))
    mean = (tot[0] / npix).float()
    var = (tot[1] / npix - (tot[0] / npix) ** 2).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + 1e-5)).float()
    assert rel_max(mean, mean_ref[0]) < 1e-4 and rel_max(rstd, rstd_ref[0]) < 1e-4


@pytest.mark.parametrize("shape", [(8, 80, 80, 256), (8, 80, 80, 1024), (2, 160, 160, 64), (3, 17, 13, 24), (2, 5, 5, 2048)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("residual", [False, True])
def test_batchnorm_train_fwd_bwd_at_size(cuda, shape, dtype, residual):
    """The flat BatchNorm passes (apply forward, backward part 1 + 2) against F.batch_norm in fp64 — at the benchmarked map sizes
    and on shapes whose vector count is not a multiple of the grid (tails, c/8 not a power of two)."""
    n, h, w, c = shape
    torch.manual_seed(c + n)
    q = lambda t: t.to(dtype).float()  # noqa: E731
    x = q(torch.randn(n, c, h, w) * 2 + 0.5)
    r = q(torch.randn(n, c, h, w)) if residual else None
    gy = q(torch.randn(n, c, h, w))
    bn = torch.nn.BatchNorm2d(c).to(cuda).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    xr = x.double().requires_grad_(True)
    rr = r.double().requires_grad_(True) if residual else None
    wr, br = bn.weight.detach().double().cpu().requires_grad_(True), bn.bias.detach().double().cpu().requires_grad_(True)
    pre = F.batch_norm(xr, None, None, wr, br, True, 0.1, bn.eps)
    pre = pre + rr if residual else pre
    yr = F.relu(pre)
    yr.backward(gy.double())
    # a pre-activation within rounding error of 0 flips the ReLU gate against the fp64 reference — an O(gy) difference in that
    # one element of the gradients (among 10^7 elements a few always sit there): compare away from the gate
    safe = (pre.detach().abs() > 1e-3)

    xs = ops.to_storage(x.to(cuda), dtype).requires_grad_(True)
    rs = ops.to_storage(r.to(cuda), dtype).requires_grad_(True) if residual else None
    y = ops.batchnorm_act(xs, bn, rs, _lib.ACT_RELU)
    yn = ops.from_storage(y, c)
    tol = 2e-5 if dtype == torch.float32 else 1.5e-2
    assert rel_max(yn, yr) < tol
    yn.backward(gy.to(cuda))
    assert float(safe.double().mean()) > 0.99
    assert rel_max(ops.from_storage(xs.grad, c).cpu() * safe, xr.grad * safe) < tol * 2, "gx"
    if residual:
        assert rel_max(ops.from_storage(rs.grad, c).cpu() * safe, rr.grad * safe) < tol * 2, "gresidual"
    # (a flipped gate moves one term of a 51 200-term channel sum: 3e-3 of the largest sum at most)
    assert rel_max(bn.weight.grad, wr.grad) < max(tol * 2, 3e-3) and rel_max(bn.bias.grad, br.grad) < max(tol * 2, 3e-3)


def test_conv_bn_act_chain_equals_unfused_chain(cuda):
    """ops.conv_bn_act (epilogue statistics) against conv -> statistics pass -> apply (CGB_EPILOGUE_STATS=0 path), forward,
    backward and running statistics, on a ResNet-bottleneck-sized problem in bf16."""
    torch.manual_seed(0)
    n, ci, co, hw = 8, 256, 256, 80
    x0 = ops.to_storage(torch.randn(n, ci, hw, hw, device=cuda), torch.bfloat16)
    wt = (torch.randn(co, ci, 3, 3, device=cuda) / (ci * 9) ** 0.5)
    gy = torch.randn(n, hw, hw, co, device=cuda).to(torch.bfloat16)
    outs = []
    for fused in (True, False):
        ops._EPI_STATS = fused
        bn = torch.nn.BatchNorm2d(co).to(cuda).train()
        x = x0.clone().requires_grad_(True)
        w = wt.clone().requires_grad_(True)
        y = ops.conv_bn_act(x, w, bn, dil=2, pad=2, act=_lib.ACT_RELU)
        y.backward(gy)
        outs.append((y.detach().float(), x.grad.float(), w.grad.clone(), bn.running_mean.clone(), bn.running_var.clone(),
                     bn.weight.grad.clone(), int(bn.num_batches_tracked)))
    ops._EPI_STATS = True
    a, b = outs
    assert a[6] == b[6] == 1
    assert rel_max(a[0], b[0]) < 1e-2            # bf16 outputs: statistics agree to ~1e-6, one bf16 ulp where rounding flips
    assert float((a[0] - b[0]).abs().mean()) < 1e-5
    assert rel_max(a[3], b[3]) < 1e-5 and rel_max(a[4], b[4]) < 1e-5
    assert rel_max(a[1], b[1]) < 2e-2 and rel_max(a[2], b[2]) < 2e-3 and rel_max(a[5], b[5]) < 2e-3
