"""bn_fuse (climategan/bn_fusion.py:97-118): the module rewrite of this package against the reference's own function on the same
weights (when /root/reference is mounted), the bias-overwrite quirk behind compat=True, and the fused model through the
inference host path."""
import pytest
import torch
import torch.nn as nn

from climategan_b200.bn_fusion import _IdentityLayer, bn_fuse, get_bn_fused_count, is_fused
from climategan_b200.generator import OmniGenerator
from climategan_b200.utils import default_masker_opts, full_opts
from oracle import refshim
from tests.golden.weights import fill_state_dict


def _randomise(model, seed):
    sd = fill_state_dict([(k, tuple(v.shape)) for k, v in model.state_dict().items()], seed)
    for k, v in sd.items():
        if k.endswith("running_var"):
            sd[k] = v.abs() + 0.5
    model.load_state_dict(sd, strict=True)
    return sd


def test_bn_fuse_folds_conv_bn_pairs_and_keeps_the_original():
    opts = default_masker_opts(nblocks=(1, 1, 1, 1), size=64)
    G = OmniGenerator(opts).eval()
    _randomise(G, 3)
    n_bn = sum(isinstance(m, nn.BatchNorm2d) for m in G.modules())
    before = {k: v.clone() for k, v in G.state_dict().items()}
    F = bn_fuse(G)
    assert F is not G and all(torch.equal(v, before[k]) for k, v in G.state_dict().items())      # deep copy, original untouched
    n_id = sum(isinstance(m, _IdentityLayer) for m in F.modules())
    assert n_id == get_bn_fused_count(F) > 0 and n_id + sum(isinstance(m, nn.BatchNorm2d) for m in F.modules()) == n_bn
    blk, ref = F.encoder.model.layer1[0], G.encoder.model.layer1[0]
    assert is_fused(blk.bn1) and blk.conv1.bias is not None
    alpha = ref.bn1.weight / torch.sqrt(ref.bn1.running_var + ref.bn1.eps)
    assert torch.allclose(blk.conv1.weight, ref.conv1.weight * alpha.view(-1, 1, 1, 1))
    assert torch.allclose(blk.conv1.bias, ref.bn1.bias - ref.bn1.running_mean * alpha)


def test_bn_fuse_bias_overwrite_quirk_is_switchable():
    """bn_fusion.py:115 ``item.bias = Parameter(beta)`` drops an existing conv bias; compat=False folds it."""
    m = nn.Sequential(nn.Conv2d(4, 6, 3, bias=True), nn.BatchNorm2d(6)).eval()
    with torch.no_grad():
        m[0].bias.uniform_(1, 2)
        m[1].running_mean.uniform_(-1, 1)
        m[1].running_var.uniform_(0.5, 2)
        m[1].weight.uniform_(0.5, 2)
        m[1].bias.uniform_(-1, 1)
    x = torch.randn(2, 4, 8, 8)
    want = m(x)
    good, quirk = bn_fuse(m, compat=False), bn_fuse(m, compat=True)
    assert torch.allclose(good(x), want, atol=1e-5)
    alpha = m[1].weight / torch.sqrt(m[1].running_var + m[1].eps)
    assert torch.allclose(quirk(x), want - (m[0].bias * alpha).view(1, -1, 1, 1), atol=1e-5)    # exactly the dropped bias term


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("arch", ["deeplabv2", "deeplabv3"])
def test_bn_fuse_matches_the_reference_function(arch):
    """Same weights into the reference's OmniGenerator and ours; the reference's bn_fuse and ours (compat=True) must fold the
    same pairs into the same numbers: identical state_dicts afterwards."""
    gen_mod, fus_mod = refshim.load("generator", "bn_fusion")
    overrides = {"gen.encoder.architecture": arch, "gen.s.architecture": arch} if arch == "deeplabv3" else None
    opts = full_opts(size=64, tasks=("d", "s", "m"), overrides=overrides)
    ours = OmniGenerator(opts).eval()
    sd = _randomise(ours, 5)
    ref = gen_mod.create_generator(opts, device=torch.device("cpu"), no_init=True).eval()
    ref.load_state_dict(sd, strict=True)
    ref_f = fus_mod.bn_fuse(ref)
    our_f = bn_fuse(ours, compat=True)
    rs, os_ = ref_f.state_dict(), our_f.state_dict()
    assert list(rs) == list(os_)
    for k in rs:
        assert torch.allclose(rs[k].float(), os_[k].float(), atol=1e-6, rtol=1e-5), k
    assert sum(type(m).__name__ == "_IdentityLayer" for m in ref_f.modules()) == get_bn_fused_count(our_f)


def test_fused_generator_runs_the_inference_host_path():
    from climategan_b200.trainer import Trainer
    from tests.dryrun import noop_library

    opts = full_opts(size=128)
    with noop_library() as lib:
        t = Trainer(opts, device=torch.device("cpu")).setup(inference=True, input_shape=(128, 128))
        n_unfused = None
        x = torch.rand(2, 3, 128, 128) * 2 - 1
        t.infer_all(x, numpy=False)
        t.infer_all(x, numpy=False)
        a = sum(lib.calls.values())
        t.infer_all(x, numpy=False)
        n_unfused = sum(lib.calls.values()) - a
        t.G = bn_fuse(t.G)
        assert get_bn_fused_count(t.G) >= 30     # (the small test encoder: 2+2+3+2 bottlenecks)
        t.infer_all(x, numpy=False)
        b = sum(lib.calls.values())
        out = t.infer_all(x, numpy=False)
        assert sum(lib.calls.values()) - b == n_unfused          # the same launches: the fold moved from the packing to the weights
        assert tuple(out["flood"].shape) == (2, 3, 128, 128)
