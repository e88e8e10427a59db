"""GPU parity of the v2 masker inference path (DeepLab-v2 ResNet encoder, DADA depth decoder, DeepLab-v2 segmentation
decoder, base mask decoder, make_m_cond) through libcgb200 against goldens from the reference's OmniGenerator (eval)."""
import pytest
import torch

from climategan_b200.generator import OmniGenerator
from climategan_b200.utils import default_masker_opts
from tests.helpers import load_golden, rel_l2, rel_max

pytestmark = pytest.mark.gpu


def _build(meta, sd, dtype, cuda):
    opts = default_masker_opts(nblocks=tuple(meta["nblocks"]), size=meta["size"])
    opts.data.transforms[-1].new_size.d = meta["d_size"]
    opts.data.transforms[-1].new_size.s = meta["s_size"]
    G = OmniGenerator(opts, storage_dtype=dtype)
    assert [(k, tuple(v.shape)) for k, v in G.state_dict().items()] == [(k, tuple(s)) for k, s in meta["shapes"]]
    G.load_state_dict(sd, strict=True)
    return G.to(cuda).eval()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_masker_decode_matches_reference_golden(cuda, dtype):
    meta, g, sd, (x, _, _) = load_golden("masker_small")
    G = _build(meta, sd, dtype, cuda)
    x = x.to(cuda)
    out = G.decode(x=x, return_z=True, return_z_depth=True)
    # stated tolerances: fp32 storage 2e-4 of full scale; bf16 storage 4e-2 (ResNet-101-style depth, ~30 convs, bf16 activations)
    tol = 2e-4 if dtype == torch.float32 else 4e-2
    assert out["d"].shape == (meta["batch"], 1, meta["d_size"], meta["d_size"])
    assert out["s"].shape == (meta["batch"], 11, meta["s_size"], meta["s_size"])
    assert out["m"].shape == (meta["batch"], 1, meta["size"], meta["size"])
    for k in ("d", "s", "m"):
        assert rel_max(out[k], torch.from_numpy(g[k])) < tol, (k, rel_max(out[k], torch.from_numpy(g[k])))
    z = out["z"].float().permute(0, 3, 1, 2)[:, :2048]
    assert rel_max(z[:, ::97], torch.from_numpy(g["z_sample"])) < tol
    zd = out["z_depth"].float().permute(0, 3, 1, 2)[:, :2048]
    assert rel_max(zd[:, ::97], torch.from_numpy(g["z_depth_sample"])) < tol
    logits = G.mask(z=out["z"], z_depth=out["z_depth"], sigmoid=False)
    ref = torch.from_numpy(g["m_logits"])
    if dtype == torch.float32:
        assert rel_max(logits, ref) < 2e-4
    else:
        assert rel_l2(logits - logits.mean(), ref.to(cuda) - ref.mean()) < 0.15
    cond = G.make_m_cond(out["d"], out["s"], x)
    assert cond.shape == (meta["batch"], 15, meta["s_size"], meta["s_size"])
    # normalize(d) divides by the (narrow) per-sample depth range, which amplifies bf16 rounding of d: 8e-2 in bf16
    assert rel_max(cond, torch.from_numpy(g["cond"])) < (2e-4 if dtype == torch.float32 else 8e-2)
    d2 = G.depth(x=x)
    assert rel_max(d2, torch.from_numpy(g["d"])) < tol


def test_masker_train_mode_uses_batch_statistics(cuda):
    """Train mode: BatchNorm normalises with BATCH statistics and updates the running ones (resnetmulti_v2.py:16-18 only
    freezes the affine parameters), gradients reach the encoder's first conv; frozen BN affine parameters get none."""
    meta, g, sd, (x, _, _) = load_golden("masker_small")
    G = _build(meta, sd, torch.float32, cuda).train()
    for m in G.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    q = meta["size"] // 4
    G.decoders["d"]._target_size = q     # native resolution of the decoders: no bicubic re-sampling on the training path
    G.decoders["s"]._target_size = q     # (ints, as find_target_size yields; depth.py:143 compares the width with an int)
    rm0 = G.encoder.model.bn1.running_mean.clone()
    z = G.encode(x.to(cuda))
    assert z.requires_grad
    assert not torch.equal(G.encoder.model.bn1.running_mean, rm0)
    assert int(G.encoder.model.bn1.num_batches_tracked) == 1
    d, z_depth = G.decode_d(z)
    s = G.decode_s(z, z_depth)
    m = G.decode_m(z)
    (d.mean() + s.mean() + m.mean()).backward()
    gw = G.encoder.model.conv1.weight.grad
    assert gw is not None and bool(torch.isfinite(gw).all()) and float(gw.abs().max()) > 0
    assert G.encoder.model.bn1.weight.grad is None
    assert G.decoders["d"].enc4_1.norm.weight.grad is not None


def test_masker_full_size_shapes(cuda):
    """640x640, ResNet-101 depth [3,4,23,3], bf16: shapes of SURVEY.md §8a (z 2048x80x80, d/s 160x160, m 640x640), finite."""
    torch.manual_seed(0)
    G = OmniGenerator(default_masker_opts(), storage_dtype=torch.bfloat16).to(cuda).eval()
    x = torch.rand(1, 3, 640, 640, device=cuda) * 2 - 1
    out = G.decode(x=x, return_z=True)
    assert out["z"].shape == (1, 80, 80, 2048)
    assert out["d"].shape == (1, 1, 160, 160) and out["s"].shape == (1, 11, 160, 160) and out["m"].shape == (1, 1, 640, 640)
    for k in ("d", "s", "m"):
        assert bool(torch.isfinite(out[k]).all()), k
    assert float(out["m"].min()) >= 0.0 and float(out["m"].max()) <= 1.0


@pytest.mark.parametrize("case", ["masker_spade", "masker_spade12"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_masker_spade_decoder_matches_reference_golden(cuda, dtype, case):
    """gen.m.use_spade (paper / release configuration): MaskSpadeDecoder conditioned on make_m_cond(d, s, x) — SPADE with a
    BatchNorm param-free norm read from running statistics — against the reference OmniGenerator.decode, two consecutive
    decodes (the spectral-norm vectors advance).  Tolerances: fp32 storage 2e-4 of full scale, bf16 4e-2."""
    from climategan_b200.utils import Dict

    meta, g, sd, (x, _, _) = load_golden(case)   # masker_spade12: cond_nc = 12, conditioning without x (reference scenario 14)
    opts = default_masker_opts(nblocks=tuple(meta["nblocks"]), size=meta["size"])
    opts.gen.m.use_spade = True
    opts.gen.m.spade.activations = Dict(all_lrelu=True)
    opts.gen.m.spade.cond_nc = meta.get("cond_nc", 15)
    G = OmniGenerator(opts, storage_dtype=dtype)
    G.load_state_dict(sd, strict=True)
    G = G.to(cuda).eval()
    tol = 2e-4 if dtype == torch.float32 else 4e-2
    o1 = G.decode(x=x.to(cuda))
    o2 = G.decode(x=x.to(cuda))
    assert o1["m"].shape == (meta["batch"], 1, meta["size"], meta["size"])
    assert rel_max(o1["m"], torch.from_numpy(g["m1"])) < tol, rel_max(o1["m"], torch.from_numpy(g["m1"]))
    assert rel_max(o2["m"], torch.from_numpy(g["m2"])) < tol, rel_max(o2["m"], torch.from_numpy(g["m2"]))
    # train mode runs (batch-statistics SPADE, running statistics updated); its parity is tests/test_gpu_full_step.py's
    rv = G.decoders["m"].spade_blocks[0].norm_0.param_free_norm.running_var.clone()
    o3 = G.train().decode(x=x.to(cuda))
    assert o3["m"].shape == o1["m"].shape and bool(torch.isfinite(o3["m"]).all())
    assert not torch.equal(rv, G.decoders["m"].spade_blocks[0].norm_0.param_free_norm.running_var)
