"""Parity AT THE BENCHMARKED CONFIGURATION ("nothing is parity-tested at the benchmarked configuration"): the conv
classes that carry the full train step in BENCH_r01 / profiles/r02_*.json, at their exact shapes (8 images, 80x80 and 640x640
maps, bf16 storage) — persistent tcgen05 kernels with 148 CTAs, 200+ KB of shared memory and multi-hundred-MB tensors — against
(i) the CUDA-core SIMT engine on the SAME bf16 operands (an independent implementation: different tiling, no TMA, no tensor
cores; fp32 accumulation on both sides, so the two differ by output rounding only) over the WHOLE tensor, and (ii) an fp64
evaluation on the host of sampled outputs."""
import pytest
import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.helpers import rel_max

pytestmark = pytest.mark.gpu

# op, n, ci, co, h, w, k, stride, dil, pad     (channel counts are STORAGE channels, as the bench's class table lists them)
CLASSES = [
    ("fwd", 8, 256, 256, 80, 80, 3, 1, 2, 2),       # ResNet layer3 conv2
    ("fwd", 8, 256, 1024, 80, 80, 1, 1, 1, 0),      # ResNet layer3 conv3
    ("fwd", 8, 1024, 256, 80, 80, 1, 1, 1, 0),      # ResNet layer3 conv1
    ("fwd", 8, 32, 128, 640, 640, 1, 1, 1, 0),      # SPADE mlp_shared on im2col patches (K = 32)
    ("fwd", 8, 128, 48, 640, 640, 3, 1, 1, 1),      # SPADE gamma||beta, C = 20 (weight-stationary halo kernel)
    ("fwd", 8, 128, 80, 640, 640, 3, 1, 1, 1),      # SPADE gamma||beta, C = 40
    ("fwd", 8, 2048, 256, 80, 80, 3, 1, 12, 12),    # ASPP atrous d12
    ("fwd", 16, 8, 64, 640, 640, 4, 2, 1, 1),       # discriminator first layer, 4x4 stride 2, 16 = real + fake
    ("dgrad", 8, 256, 256, 80, 80, 3, 1, 2, 2),
    ("dgrad", 8, 1024, 256, 80, 80, 1, 1, 1, 0),
    ("dgrad", 8, 128, 80, 640, 640, 3, 1, 1, 1),    # the painter's heaviest dgrad (1.3 ms)
    ("dgrad", 8, 64, 128, 320, 320, 4, 2, 1, 1),    # stride-2 dgrad as parity-class sub-convolutions
    # second epilogue operand (derivative mask / residual) delivered by TMA into the staging tile (conv_tc.cu, struct EpiOperand)
    ("dgrad_relu", 8, 128, 48, 640, 640, 3, 1, 1, 1),   # weight-stationary kernel, one staging tile, two halves
    ("dgrad_relu", 8, 128, 80, 640, 640, 3, 1, 1, 1),   # streaming kernel, two staging tiles
    ("dgrad_lrelu", 4, 128, 80, 321, 323, 3, 1, 1, 1),  # ragged borders: tile tails clipped by the tensor maps
    ("dgrad_relu", 8, 256, 256, 160, 160, 3, 1, 1, 1),  # VGG conv3_x: N = 256, four halves, CTA-pair kernel
    ("dgrad_relu", 8, 64, 64, 640, 640, 3, 1, 1, 1),    # VGG conv1_2: one half
    ("dgrad_lrelu", 8, 320, 192, 80, 80, 3, 1, 1, 1),   # two N tiles of 160 channels: per-thread copy-out with the mask in phase 1
    ("dgrad_relu", 8, 512, 64, 80, 80, 3, 1, 1, 1),     # two N tiles of 256 channels, TMA store, four halves each
    ("fwd_res", 8, 256, 256, 80, 80, 3, 1, 1, 1),       # residual add in the epilogue
    ("wgrad", 8, 256, 256, 80, 80, 3, 1, 2, 2),
    ("wgrad", 8, 1024, 256, 80, 80, 1, 1, 1, 0),
    ("wgrad", 8, 256, 1024, 80, 80, 1, 1, 1, 0),
    ("wgrad", 8, 128, 80, 640, 640, 3, 1, 1, 1),    # halo wgrad on a 640x640 map
]


def _geom(k, stride, dil, pad, engine):
    return ops.ConvGeom(k, k, stride, dil, pad, _lib.PAD_ZERO, _lib.ACT_NONE, 0.2, engine)


@pytest.mark.parametrize("case", CLASSES, ids=lambda c: f"{c[0]}-{c[2]}to{c[3]}-k{c[6]}s{c[7]}d{c[8]}-{c[4]}x{c[5]}-n{c[1]}")
def test_bench_class_tcgen05_matches_simt_and_fp64(cuda, case):
    op, n, ci, co, h, w, k, stride, dil, pad = case
    torch.manual_seed(CLASSES.index(case) + 11)
    dt = torch.bfloat16
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x = torch.randn(n, h, w, ci, device=cuda).to(dt)
    wp = (torch.randn(co, k * k, ci, device=cuda) / (ci * k * k) ** 0.5).to(dt)
    gy = torch.randn(n, ho, wo, co, device=cuda).to(dt)
    g_tc, g_simt = _geom(k, stride, dil, pad, _lib.ENGINE_TCGEN05), _geom(k, stride, dil, pad, _lib.ENGINE_SIMT)
    w_oihw = wp.float().view(co, k, k, ci).permute(0, 3, 1, 2).double().cpu()
    gen = torch.Generator().manual_seed(5)
    if op == "fwd":
        a = ops.conv_fwd_raw(x, wp, None, None, g_tc)
        b = ops.conv_fwd_raw(x, wp, None, None, g_simt)
        # fp64 on sampled output pixels: the receptive field of each, gathered on the host
        for _ in range(24):
            i = int(torch.randint(0, n, (1,), generator=gen))
            oy = int(torch.randint(0, ho, (1,), generator=gen))
            ox = int(torch.randint(0, wo, (1,), generator=gen))
            patch = torch.zeros(ci, k, k, dtype=torch.float64)
            for dy in range(k):
                for dx in range(k):
                    iy, ix = oy * stride - pad + dy * dil, ox * stride - pad + dx * dil
                    if 0 <= iy < h and 0 <= ix < w:
                        patch[:, dy, dx] = x[i, iy, ix].double().cpu()
            ref = (w_oihw * patch.unsqueeze(0)).sum((1, 2, 3))
            got = a[i, oy, ox].double().cpu()
            assert float((got - ref).abs().max()) < 1e-2 * max(float(ref.abs().max()), 0.5), (i, oy, ox)
    elif op == "dgrad":
        a = ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g_tc)
        b = ops.conv_dgrad_raw(gy, wp.clone(), (n, h, w, ci), g_simt)
    elif op in ("dgrad_relu", "dgrad_lrelu"):
        dact = _lib.ACT_RELU if op == "dgrad_relu" else _lib.ACT_LRELU
        a = ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g_tc, dact, x)
        b = ops.conv_dgrad_raw(gy, wp.clone(), (n, h, w, ci), g_simt, dact, x)
        # the mask really acted: about half of the elements are zeroed (relu) / scaled by the slope (lrelu)
        plain = ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g_tc)
        neg = x.float() <= 0
        if op == "dgrad_relu":
            assert float(a[neg].abs().max()) == 0.0
            assert torch.equal(a[~neg], plain[~neg])
        else:
            assert rel_max(a[neg].float(), 0.2 * plain[neg].float()) < 1.0 / 64
            assert torch.equal(a[~neg], plain[~neg])
    elif op == "fwd_res":
        res = torch.randn(n, ho, wo, co, device=cuda).to(dt)
        a = ops.conv_fwd_raw(x, wp, None, res, g_tc)
        b = ops.conv_fwd_raw(x, wp, None, res, g_simt)
    else:
        a, _ = ops.conv_wgrad_raw(x, gy, g_tc, False)
        b, _ = ops.conv_wgrad_raw(x, gy, g_simt, False)
        # fp64 on sampled weight entries needs the full pixel reduction: use a 2-image slice through F.conv2d's adjoint instead
        xs, gs = x[:2].float().permute(0, 3, 1, 2).double().cpu(), gy[:2].float().permute(0, 3, 1, 2).double().cpu()
        if xs.numel() * k * k < 2e8:
            ref = torch.nn.grad.conv2d_weight(xs, (co, ci, k, k), gs, stride=stride, padding=pad, dilation=dil)
            a2, _ = ops.conv_wgrad_raw(x[:2].contiguous(), gy[:2].contiguous(), g_tc, False)
            got = a2.view(co, k, k, ci).permute(0, 3, 1, 2).double().cpu()
            assert rel_max(got, ref) < 2e-3
    torch.cuda.synchronize()
    assert a.shape == b.shape
    # whole tensor, tcgen05 vs SIMT: identical operands and fp32 accumulation, so at most one bf16 ulp (2^-8 relative per element)
    # on outputs (fwd / dgrad); the fp32 weight gradient differs by summation order only
    if op == "wgrad":
        assert rel_max(a, b) < 2e-3
    else:
        d = (a.float() - b.float()).abs()
        scale = b.float().abs().max()
        assert float(d.max() / scale) < 1.0 / 64, float(d.max() / scale)
        assert float((d > 0).float().mean()) < 0.25          # most elements round identically
        assert float(d.mean() / b.float().abs().mean()) < 2e-3
