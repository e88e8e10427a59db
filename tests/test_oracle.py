"""CPU tests: the oracle (oracle/painter_oracle.py) against the committed golden vectors that the
unmodified reference produced, and — when the reference tree is present — against the reference
modules themselves."""
import os

import numpy as np
import pytest
import torch

from oracle import painter_oracle as po
from oracle import refshim
from tests.golden.weights import fill_state_dict
from tests.helpers import GOLDEN, load_golden, rel_max


def _oracle_run(sd, x, m, target, meta):
    sd = {k: v.clone().requires_grad_(not (k.endswith("_u") or k.endswith("_v"))) for k, v in sd.items()}
    z = meta["size"] // 2 ** meta["spade_n_up"]
    out = po.paint(sd, m, x, z, z, po.n_up_spades_of(sd))
    loss = torch.nn.functional.l1_loss(out, target)
    loss.backward()
    return sd, out, loss


def test_oracle_matches_golden():
    meta, g, sd, (x, m, t) = load_golden()
    sd, out, loss = _oracle_run(sd, x, m, t, meta)
    assert rel_max(out, torch.from_numpy(g["out"])) < 1e-5
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    norms = np.array([float(sd[k].grad.norm()) for k in meta["grad_keys"]])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-4, atol=1e-7)
    for k, v in g.items():
        if k.startswith("grad::"):
            assert rel_max(sd[k[6:]].grad, torch.from_numpy(v)) < 2e-4, k
    # spectral-norm state advanced exactly one power iteration
    assert rel_max(sd["head_0.conv_0.module.weight_u"], torch.from_numpy(g["u_after"])) < 1e-5
    assert rel_max(sd["up_spades.0.conv_s.module.weight_v"], torch.from_numpy(g["v_after"])) < 1e-5


def test_oracle_second_forward_and_no_paste():
    meta, g, sd, (x, m, t) = load_golden()
    z = meta["size"] // 2 ** meta["spade_n_up"]
    n_up = po.n_up_spades_of(sd)
    with torch.no_grad():
        sd1 = {k: v.clone() for k, v in sd.items()}
        fake = po.paint(sd1, m, x, z, z, n_up, paste=False)
        assert rel_max(fake, torch.from_numpy(g["fake_no_paste"])) < 1e-5
        out2 = po.paint(sd1, m, x, z, z, n_up)  # u/v advanced by the first call
        assert rel_max(out2, torch.from_numpy(g["out_second_forward"])) < 1e-5


def test_spectral_norm_properties():
    torch.manual_seed(3)
    w = torch.randn(24, 16, 3, 3)
    u = po.l2normalize(torch.randn(24))
    v = po.l2normalize(torch.randn(16 * 9))
    for _ in range(200):
        wn, u, v = po.spectral_norm_weight(w, u, v)
    s = torch.linalg.svdvals(w.view(24, -1))[0]
    assert abs(float(torch.linalg.svdvals(wn.view(24, -1))[0]) - 1.0) < 1e-3
    assert abs(float((w / wn).flatten()[0]) - float(s)) / float(s) < 1e-3


@pytest.mark.reference
@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_oracle_vs_reference_modules():
    """Bit-level agreement of the restatement with the reference's own modules (CPU, fp32)."""
    from climategan_b200.utils import default_painter_opts

    painter_mod = refshim.load("painter")
    for latent, n_up, size in [(40, 3, 32), (16, 4, 64)]:
        opts = default_painter_opts(latent_dim=latent, spade_n_up=n_up)
        torch.manual_seed(1)
        ref = painter_mod.PainterSpadeDecoder(opts)
        ref.set_latent_shape(size, True)
        sd = {k: v.clone() for k, v in ref.state_dict().items()}
        cond = torch.rand(2, 3, size, size) * 2 - 1
        with torch.no_grad():
            a = ref(None, cond)
            b = po.painter_forward(sd, cond, ref.z_h, ref.z_w, po.n_up_spades_of(sd))
        assert float((a - b).abs().max()) < 1e-6
        # u/v state after the forward must agree too
        for k, v in ref.state_dict().items():
            if k.endswith("_u") or k.endswith("_v"):
                assert float((v - sd[k]).abs().max()) < 1e-6, k


def test_discriminator_oracle_matches_golden():
    """oracle/discriminator_oracle.py vs the reference's OmniDiscriminator['p'] + GANLoss + FeatMatchLoss + HingeLoss."""
    from oracle import discriminator_oracle as do

    meta, g, sd, _ = load_golden("disc_small")
    sd = {k[2:]: v.clone().requires_grad_(not k.endswith(("_u", "_v"))) for k, v in sd.items()}  # strip "p."
    real = torch.from_numpy(g["real"])
    fake = torch.from_numpy(g["fake"]).requires_grad_(True)
    out = do.multiscale_forward(sd, torch.cat([real, fake], 0))
    pred_real, pred_fake = do.divide_pred(out)
    g_gan = do.gan_loss(pred_fake, True)
    g_feat = do.feat_match_loss(pred_real, pred_fake)
    d_hinge = do.hinge_loss(pred_fake, False) + do.hinge_loss(pred_real, True)
    (g_gan + 10.0 * g_feat).backward()
    assert abs(float(g_gan) - float(g["g_gan"])) < 1e-6
    assert abs(float(g_feat) - float(g["g_feat"])) < 1e-5
    assert abs(float(d_hinge) - float(g["d_hinge"])) < 1e-5
    for i in range(meta["num_D"]):
        assert rel_max(out[i][-1], torch.from_numpy(g[f"pred_{i}"])) < 1e-5
        assert rel_max(out[i][1], torch.from_numpy(g[f"feat_{i}_1"])) < 1e-5
    assert rel_max(fake.grad, torch.from_numpy(g["fake_grad"])) < 1e-4
    norms = np.array([float(sd[k[2:]].grad.norm()) for k in meta["grad_keys"]])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=5e-4, atol=1e-7)


def test_trainer_oracle_matches_reference_step_golden():
    """oracle/trainer_oracle.py (painter G/D losses + ExtraAdam) vs 4 optimiser steps run with the reference's modules."""
    import json
    import os

    from oracle import painter_oracle as po
    from oracle import trainer_oracle as to
    from oracle.painter_oracle import SNState
    from tests.golden.weights import fill_state_dict, synth_inputs
    from tests.helpers import GOLDEN

    meta = json.load(open(os.path.join(GOLDEN, "painter_step.json")))
    g = dict(np.load(os.path.join(GOLDEN, "painter_step.npz")))

    def mk(shapes, seed):
        sd = fill_state_dict([(k, tuple(s)) for k, s in shapes], seed)
        return {k: v.clone().requires_grad_(not k.endswith(("_u", "_v"))) for k, v in sd.items()}

    gsd, dsd_full, vsd = mk(meta["g_shapes"], 11), mk(meta["d_shapes"], 12), mk(meta["v_shapes"], 13)
    for v in vsd.values():
        v.requires_grad_(False)
    dsd = {k[2:]: v for k, v in dsd_full.items()}
    x, m, _ = synth_inputs(meta["batch"], meta["size"], 5)
    z = meta["size"] // 2 ** meta["spade_n_up"]
    g_opt = to.ExtraAdam([v for v in gsd.values() if v.requires_grad], lr=0.00005, betas=(0.9, 0.999))
    d_opt = to.ExtraAdam(list(dsd.values()), lr=0.00002, betas=(0.5, 0.999))   # every D tensor, u / v included (optim.py:72)
    g_sn, d_sn = SNState(gsd), SNState(dsd)
    logs = []
    for it in range(4):
        gs = it // 2
        if it % 2 == 0:
            for v in gsd.values():
                v.grad = None
            for v in dsd.values():
                v.grad = None
            loss, terms = to.painter_g_loss(gsd, dsd, vsd, x, m, z, g_sn=g_sn, d_sn=d_sn)
            loss.backward()
            for v in dsd.values():
                v.grad = None  # D is frozen during update_G ...
                v.requires_grad_(True)  # ... and un-frozen wholesale afterwards: u / v train from here on (trainer.py:971-973)
            (g_opt.extrapolation if gs % 2 == 0 else g_opt.step)()
            logs += [float(terms["vgg"]), float(terms["gan"]), float(terms["featmatch"])]
        else:
            for v in dsd.values():
                v.grad = None
            ld = to.painter_d_loss(gsd, dsd, x, m, z, g_sn=g_sn, d_sn=d_sn)
            ld.backward()
            (d_opt.extrapolation if gs % 2 == 0 else d_opt.step)()
            logs.append(float(ld))
    np.testing.assert_allclose(np.array(logs), g["logs"], rtol=2e-5)
    for k, v in g.items():
        if k.startswith("G::"):
            assert rel_max(gsd[k[3:]], torch.from_numpy(v)) < 1e-5, k
        if k.startswith("D::"):
            assert rel_max(dsd_full[k[3:]], torch.from_numpy(v)) < 1e-5, k


def test_masker_oracle_matches_reference_golden():
    """oracle/masker_oracle.py vs the reference OmniGenerator.decode / make_m_cond / mask in eval mode."""
    from oracle import masker_oracle as mo

    meta, g, sd, (x, _, _) = load_golden("masker_small")
    with torch.no_grad():
        out = mo.decode(sd, x, meta["d_size"], meta["s_size"])
        cond = mo.make_m_cond(out["d"], out["s"], x)
    for k in ("d", "s", "m"):
        assert rel_max(out[k], torch.from_numpy(g[k])) < 1e-4, k
    assert rel_max(cond, torch.from_numpy(g["cond"])) < 1e-4
    assert rel_max(out["z"][:, ::97], torch.from_numpy(g["z_sample"])) < 1e-4
    assert rel_max(out["z_depth"][:, ::97], torch.from_numpy(g["z_depth_sample"])) < 1e-4


def test_full_step_oracle_matches_reference_trainer_golden():
    """oracle/full_step_oracle.py (restated masker losses, train-mode BatchNorm masker, G/D loss assembly) against the
    first iteration of the reference's own Trainer.update_G / update_D (tests/golden/full_step.*): every logged loss term,
    the gradient norm of every parameter, sampled full gradients."""
    import json as _json

    from climategan_b200.utils import full_opts, synth_batch
    from oracle import full_step_oracle as fo

    meta = _json.load(open(os.path.join(GOLDEN, "full_step.json")))
    g = dict(np.load(os.path.join(GOLDEN, "full_step.npz")))
    size, batch = meta["size"], meta["batch"]
    mk = lambda shapes, seed: fill_state_dict([(k, tuple(s)) for k, s in shapes], seed)  # noqa: E731
    gsd, dsd, vsd = mk(meta["g_shapes"], 21), mk(meta["d_shapes"], 22), mk(meta["v_shapes"], 23)
    g_names, d_names = meta["g_param_names"], meta["d_param_names"]
    frozen = {k for k in g_names if ".bn" in k or "downsample.1" in k}  # encoder BatchNorm affine params (resnetmulti_v2.py:16-18)
    for k in g_names:
        gsd[k].requires_grad_(not k.endswith(("weight_u", "weight_v")) and not (k.startswith("encoder.") and k in frozen))
    opts = full_opts(size=size)
    mdb = synth_batch(opts, batch, size, meta["seeds"]["inputs"])
    z_hw = size // 2 ** opts.gen.p.spade_n_up
    # ---- update_G (D frozen)
    loss, terms = fo.full_g_loss(gsd, dsd, vsd, mdb, z_hw)
    loss.backward()
    ref = meta["logs"][0]
    assert abs(float(loss) - ref["gen.total_loss"]) < 1e-5 * abs(ref["gen.total_loss"])
    pairs = {"d.s": "gen.task.d.s", "s.crossent.s": "gen.task.s.crossent.s", "s.minent.r": "gen.task.s.minent.r",
             "s.advent.r": "gen.task.s.advent.r", "m.tv.r": "gen.task.m.tv.r", "m.tv.s": "gen.task.m.tv.s",
             "m.bce.s": "gen.task.m.bce.s", "m.gi.r": "gen.task.m.gi.r", "m.minent.r": "gen.task.m.minent.r",
             "m.advent.r": "gen.task.m.advent.r", "p.vgg": "gen.p.vgg", "p.gan": "gen.p.gan", "p.featmatch": "gen.p.featmatch"}
    for ours, theirs in pairs.items():
        assert abs(float(terms[ours]) - ref[theirs]) <= 1e-5 * abs(ref[theirs]) + 1e-7, (ours, float(terms[ours]), ref[theirs])
    gn = g["G.gradnorm"]
    for name, r in zip(g_names, gn):
        p = gsd[name]
        if r < 0:
            assert p.grad is None or not p.requires_grad, name
        else:
            assert abs(float(p.grad.norm()) - r) <= 1e-4 * r + 1e-7 * gn.max(), (name, float(p.grad.norm()), r)
    for k in meta["full_g"]:
        a = gsd[k].grad.numpy().reshape(-1)
        a = a[::max(1, -(-a.size // 8192))]
        assert np.abs(a - g["G.grad::" + k]).max() <= 1e-4 * np.abs(g["G.grad::" + k]).max(), k


def test_masker_spade_oracle_matches_reference_golden():
    """oracle MaskSpadeDecoder path (SPADE with BatchNorm running statistics, 15-channel conditioning) vs the reference
    OmniGenerator.decode with gen.m.use_spade, two consecutive decodes."""
    from oracle import masker_oracle as mo

    meta, g, sd, (x, _, _) = load_golden("masker_spade")
    sn = mo.SNState(sd)
    with torch.no_grad():
        q = meta["size"] // 4
        o1 = mo.decode_spade(sd, x, q, q, sn)
        o2 = mo.decode_spade(sd, x, q, q, sn)
    assert rel_max(o1["m"], torch.from_numpy(g["m1"])) < 1e-5
    assert rel_max(o2["m"], torch.from_numpy(g["m2"])) < 1e-5
    assert rel_max(o1["d"], torch.from_numpy(g["d"])) < 1e-5 and rel_max(o1["s"], torch.from_numpy(g["s"])) < 1e-5


def _v3_functional(meta, d, s, m):
    rs = np.random.RandomState(meta["functional_seed"])
    wd, ws, wm = (torch.from_numpy(rs.standard_normal(size=tuple(t.shape)).astype(np.float32)) for t in (d, s, m))
    return wd, ws, wm


def test_masker_v3_oracle_matches_reference_golden():
    """oracle/masker_v3_oracle.py (the reference's default deeplabv3 masker) vs the reference modules: eval decode, and a
    train-mode forward/backward (batch-statistics BatchNorm) with the gradient norm of every parameter."""
    from oracle import masker_oracle as mo
    from oracle import masker_v3_oracle as v3

    meta, g, sd, (x, _, _) = load_golden("masker_v3")
    q = meta["size"] // 4
    sdt = {k: v.clone() for k, v in sd.items()}
    for k in meta["param_names"]:
        sdt[k].requires_grad_(not k.endswith(("weight_u", "weight_v")))
    sn = v3.SNState(sdt)   # one state through both passes: the eval decode advances the spectral-norm u / v first, as in the golden
    with torch.no_grad():
        out = v3.forward(sdt, x, q, q, sn)
    for k in ("d", "s", "m"):
        assert rel_max(out[k], torch.from_numpy(g[k])) < 1e-5, k
    with mo.train_mode():
        o = v3.forward(sdt, x, q, q, sn)
    wd, ws, wm = _v3_functional(meta, o["d"], o["s"], o["m_logits"])
    loss = (o["d"] * wd).mean() + (o["s"] * ws).mean() + (o["m_logits"] * wm).mean()
    loss.backward()
    assert abs(float(loss) - float(g["train_loss"])) < 1e-6 + 1e-5 * abs(float(g["train_loss"]))
    for name, r in zip(meta["param_names"], g["gradnorm"]):
        if r >= 0:
            assert abs(float(sdt[name].grad.norm()) - r) <= 1e-4 * r + 1e-9, (name, float(sdt[name].grad.norm()), r)
    for k in ("encoder.bn1.running_mean", "encoder.layer4.1.bn2.running_var", "decoders.s.aspp.conv_out.bn.running_var"):
        assert rel_max(sdt[k], torch.from_numpy(g["final::" + k])) < 1e-5, k


@pytest.mark.skipif(not refshim.available(), reason="reference tree not mounted")
def test_committed_fixtures_are_what_the_generating_script_produces(tmp_path, monkeypatch):
    """tests/golden/make_golden.py, re-run against /root/reference into a scratch directory, reproduces the committed
    fixtures (two of the small ones: a Trainer step fixture and an eval decode): same metadata, same arrays to 1e-6."""
    import json

    import tests.golden.make_golden as mg

    monkeypatch.setattr(mg, "HERE", str(tmp_path))
    mg.run_full_step_case(name="mask_only_step_v3", tasks=("m",), overrides=mg.V3_MASKER)
    mg.run_masker_spade_case(name="masker_spade12", cond_nc=12)
    for name in ("mask_only_step_v3", "masker_spade12"):
        new_meta = json.load(open(os.path.join(str(tmp_path), name + ".json")))
        old_meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
        for k in old_meta:
            if k not in ("logs", "torch", "reference"):
                assert new_meta[k] == old_meta[k], (name, k)
        if "logs" in old_meta:
            for a, b in zip(new_meta["logs"], old_meta["logs"]):
                assert a.keys() == b.keys()
                for k in a:
                    assert abs(a[k] - b[k]) <= 1e-5 * abs(b[k]) + 1e-7, (name, k, a[k], b[k])
        new, old = np.load(os.path.join(str(tmp_path), name + ".npz")), np.load(os.path.join(GOLDEN, name + ".npz"))
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            scale = max(float(np.abs(old[k]).max()), 1e-12)
            assert float(np.abs(new[k] - old[k]).max()) <= 2e-3 * scale + 1e-9, (name, k)
