"""GPU parity of the whole painter path (PainterSpadeDecoder / OmniGenerator.paint + L1 + backward)
through libcgb200, against the committed golden vectors the unmodified reference produced and
against the oracle on fresh seeded inputs."""
import numpy as np
import pytest
import torch

from climategan_b200 import _lib
from climategan_b200 import ops
from climategan_b200.generator import OmniGenerator
from climategan_b200.utils import default_painter_opts
from tests.helpers import cosine, load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu

# Stated tolerances (max|delta| / max|ref|, SURVEY.md §8d form):
#   fp32 storage  : forward 1e-4, per-parameter gradients 1e-3
#   bf16 storage  : forward 5e-2 of full scale (measured 2.5e-2 max, 7.5e-3 rel-L2), gradients cosine >= 0.98
#                   and rel-L2 <= 0.25 (measured 0.986 / 0.17 worst on this 32x32 case: bf16 has 8 mantissa
#                   bits, ~25 conv layers deep, and leaky-relu sign flips dominate on tiny maps), loss 1e-2 rel.
# The 1e-3 bar of BASELINE.json is met by the fp32-storage mode (measured 3e-6); bf16's epsilon is 3.9e-3.
TOL = {torch.float32: dict(fwd=1e-4, loss=1e-5), torch.bfloat16: dict(fwd=5e-2, loss=1e-2)}


def _build(meta, sd, dtype, dev):
    opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"])
    G = OmniGenerator(opts, latent_shape=meta["size"], storage_dtype=dtype)
    G.painter.load_state_dict(sd, strict=True)
    return G.to(dev).train()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_paint_matches_reference_golden(cuda, dtype):
    meta, g, sd, (x, m, t) = load_golden()
    G = _build(meta, sd, dtype, cuda)
    x, m, t = x.to(cuda), m.to(cuda), t.to(cuda)
    out = G.paint(m, x)
    loss = ops.l1_loss(out, t)
    loss.backward()
    tol = TOL[dtype]
    assert out.shape == (meta["batch"], 3, meta["size"], meta["size"]) and out.dtype == torch.float32
    assert rel_max(out, torch.from_numpy(g["out"])) < tol["fwd"]
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < tol["loss"]
    params = dict(G.painter.named_parameters())
    norms_ref = dict(zip(meta["grad_keys"], g["grad_norms"]))
    # spectral-norm u/v advanced one power iteration, in place (norms.py:106-108)
    assert rel_max(params["head_0.conv_0.module.weight_u"], torch.from_numpy(g["u_after"])) < 1e-4
    assert rel_max(params["up_spades.0.conv_s.module.weight_v"], torch.from_numpy(g["v_after"])) < 1e-4
    for k, v in g.items():
        if not k.startswith("grad::"):
            continue
        gr = torch.from_numpy(v)
        gm = params[k[6:]].grad
        assert gm is not None, k
        if dtype == torch.float32:
            assert rel_max(gm, gr) < 1e-3, k
        else:
            assert cosine(gm, gr) > 0.98 and rel_l2(gm, gr) < 0.25, (k, cosine(gm, gr), rel_l2(gm, gr))
    # every trainable parameter got a gradient of the right magnitude
    bad = []
    for k in meta["grad_keys"]:
        gm = params[k].grad
        assert gm is not None, k
        nr = norms_ref[k]
        if nr < 1e-6:  # biases feeding an instance norm: mathematically zero gradient
            continue
        rtol = 2e-3 if dtype == torch.float32 else 0.1
        if abs(float(gm.norm()) - nr) / nr > rtol:
            bad.append((k, float(gm.norm()), nr))
    assert not bad, bad[:5]
    # second forward uses the advanced u/v
    with torch.no_grad():
        out2 = G.paint(m, x)
    assert rel_max(out2, torch.from_numpy(g["out_second_forward"])) < tol["fwd"]


def test_no_paste_and_painter_forward(cuda):
    meta, g, sd, (x, m, t) = load_golden()
    G = _build(meta, sd, torch.float32, cuda)
    with torch.no_grad():
        fake = G.paint(m.to(cuda), x.to(cuda), no_paste=True)
    assert rel_max(fake, torch.from_numpy(g["fake_no_paste"])) < 1e-4
    # PainterSpadeDecoder.forward(z=None, cond) keeps the reference NCHW contract
    G = _build(meta, sd, torch.float32, cuda)
    with torch.no_grad():
        fake2 = G.painter(None, (x * (1 - m)).to(cuda))
    assert rel_max(fake2, torch.from_numpy(g["fake_no_paste"])) < 1e-4


@pytest.mark.parametrize("latent,n_up,size,batch", [(16, 4, 64, 1), (24, 2, 16, 3)])
def test_paint_matches_oracle_fresh_inputs(cuda, latent, n_up, size, batch):
    """Other depths / ragged channel counts, checked against the oracle run on the same seeded data."""
    from oracle import painter_oracle as po
    from tests.golden.weights import fill_state_dict, synth_inputs

    opts = default_painter_opts(latent_dim=latent, spade_n_up=n_up)
    G = OmniGenerator(opts, latent_shape=size, storage_dtype=torch.float32)
    shapes = [(k, tuple(v.shape)) for k, v in G.painter.state_dict().items()]
    sd = fill_state_dict(shapes, seed=latent)
    G.painter.load_state_dict(sd)
    G = G.to(cuda)
    x, m, t = synth_inputs(batch, size, seed=size)
    sdr = {k: v.clone().requires_grad_(not k.endswith(("_u", "_v"))) for k, v in sd.items()}
    z = size // 2 ** n_up
    out_r = po.paint(sdr, m, x, z, z, po.n_up_spades_of(sdr))
    torch.nn.functional.l1_loss(out_r, t).backward()
    out = G.paint(m.to(cuda), x.to(cuda))
    ops.l1_loss(out, t.to(cuda)).backward()
    assert rel_max(out, out_r) < 1e-4
    for k, p in G.painter.named_parameters():
        if p.requires_grad and float(sdr[k].grad.norm()) > 1e-6:
            assert rel_max(p.grad, sdr[k].grad) < 2e-3, k


def test_full_size_properties(cuda):
    """640x640 (BASELINE.json size), bf16: size-independent properties — output range of tanh/paste,
    unmasked pixels reproduce x exactly, run-to-run reproducibility of the forward, finite gradients everywhere."""
    torch.manual_seed(0)
    opts = default_painter_opts()
    G = OmniGenerator(opts, latent_shape=640, storage_dtype=torch.bfloat16).to(cuda)
    sd0 = {k: v.clone() for k, v in G.painter.state_dict().items()}
    x = torch.rand(1, 3, 640, 640, device=cuda) * 2 - 1
    m = (torch.rand(1, 1, 640, 640, device=cuda) > 0.5).float()
    out = G.paint(m, x)
    assert out.shape == (1, 3, 640, 640)
    assert float(out.abs().max()) <= 1.0 + 1e-6
    keep = (m == 0).expand_as(x)
    assert torch.equal(out[keep], x[keep])  # paste_original_content (generator.py:295-296)
    ops.l1_loss(out, torch.zeros_like(out)).backward()
    for k, p in G.painter.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
    G.painter.load_state_dict(sd0)  # rewind spectral-norm u/v
    with torch.no_grad():
        out_b = G.paint(m, x)
    # statistics use fp64 atomics whose summation order varies run to run: reproducible to bf16 noise only
    assert rel_l2(out_b, out) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_painter_explicit_z_and_final_shortcut_match_reference_golden(cuda, dtype):
    """gen.p.no_z = false (an explicit latent replaces fc(cond), painter.py:149-152) with gen.p.use_final_shortcut (the last
    SPADE block conditioned on lrelu(BatchNorm(SN conv1x1(y))), painter.py:101-109,163-164 — a conditioning map that receives a
    gradient through SPADE's mlp_shared): train-mode forward + L1 + backward, then an eval-mode forward on the updated running
    statistics, against the reference modules (tests/golden/painter_z_shortcut.*).  Tolerances as in
    test_paint_matches_reference_golden."""
    meta, g, sd, (x, m, t) = load_golden("painter_z_shortcut")
    opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"])
    opts.gen.p.no_z = False
    opts.gen.p.use_final_shortcut = True
    G = OmniGenerator(opts, latent_shape=meta["size"], storage_dtype=dtype)
    assert [(k, tuple(v.shape)) for k, v in G.painter.state_dict().items()] == [(k, tuple(s)) for k, s in meta["shapes"]]
    G.painter.load_state_dict(sd, strict=True)
    G = G.to(cuda).train()
    x, m, t = x.to(cuda), m.to(cuda), t.to(cuda)
    z = torch.from_numpy(g["z"]).to(cuda)
    zs = G.sample_painter_z(meta["batch"], cuda)
    assert zs.shape == z.shape and zs.dtype == torch.float32
    out = G.painter(z, x * (1.0 - m))
    loss = ops.l1_loss(out, t)
    loss.backward()
    tol = TOL[dtype]
    assert rel_max(out, torch.from_numpy(g["out"])) < tol["fwd"]
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < tol["loss"]
    params = dict(G.painter.named_parameters())
    assert params["fc.weight"].grad is None   # bypassed by the explicit z, as in the reference
    bn = G.painter.final_shortcut[1]
    stat_tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert rel_max(bn.running_mean, torch.from_numpy(g["running_mean"])) < stat_tol
    assert rel_max(bn.running_var, torch.from_numpy(g["running_var"])) < stat_tol
    for k in meta["full"]:
        gr, gm = torch.from_numpy(g["grad::" + k]), params[k].grad
        assert gm is not None, k
        if dtype == torch.float32:
            assert rel_max(gm, gr) < 1e-3, (k, rel_max(gm, gr))
        else:
            assert cosine(gm, gr) > 0.98 and rel_l2(gm, gr) < 0.25, (k, cosine(gm, gr), rel_l2(gm, gr))
    norms_ref = dict(zip(meta["grad_keys"], g["grad_norms"]))
    scale = max(norms_ref.values())
    bad = []
    for k, r in norms_ref.items():
        if r < 1e-6:  # biases feeding an instance norm: mathematically zero gradient (rounding residue on both sides)
            continue
        a = float(params[k].grad.norm())
        rt = 2e-3 if dtype == torch.float32 else 0.25
        if abs(a - r) > rt * r + 1e-5 * scale:
            bad.append((k, a, r))
    assert not bad, bad[:10]
    G.eval()
    with torch.no_grad():
        out_eval = G.painter(z, x * (1.0 - m))
    assert rel_max(out_eval, torch.from_numpy(g["out_eval"])) < tol["fwd"]
    # the full paint() path draws its own z
    assert G.paint(m, x).shape == x.shape
