"""Host-side behaviours the round-1 advisor flagged, on CPU through the dry-run library stand-in:
G.mask() builds the SPADE conditioning itself, resume() follows the reference's rules, VGGLoss loads / warns about weights."""
import warnings

import pytest
import torch

from climategan_b200.trainer import Trainer
from climategan_b200.utils import full_opts
from tests.dryrun import noop_library


def test_paint_and_mask_with_the_spade_masker():
    """generator.py:257-262 / trainer.py:1866: with gen.m.use_spade the conditioning tensor is built inside G.mask (cond=None),
    which needs x — Trainer.compute_flood passes it."""
    opts = full_opts(size=128, use_spade=True)
    with noop_library():
        t = Trainer(opts, device=torch.device("cpu")).setup(inference=True, input_shape=(128, 128))
        x = torch.rand(2, 3, 128, 128) * 2 - 1
        out = t.paint_and_mask(x)
        assert tuple(out.shape) == (2, 3, 128, 128)
        m = t.G.mask(x)
        assert tuple(m.shape) == (2, 1, 128, 128)
        with pytest.raises(ValueError):          # cond_nc == 15 without x (generator.py:220-225)
            t.G.mask(z=t.G.encode(x))


def _trainer(tmp_path, tasks=("d", "s", "m", "p")):
    opts = full_opts(size=128, tasks=tasks)
    opts.output_path = str(tmp_path)
    return Trainer(opts, device=torch.device("cpu")).setup(input_shape=(128, 128))


def test_resume_rounds_an_odd_step_and_warns_about_a_foreign_optimizer_state(tmp_path):
    with noop_library():
        t = _trainer(tmp_path)
        t.logger.global_step, t.logger.epoch = 7, 2
        path = t.save()
        ck = torch.load(path)
        assert set(ck) == {"epoch", "step", "G", "g_opt", "D", "d_opt"}      # the reference's checkpoint keys (trainer.py:403-412)
        # a reference checkpoint: per-parameter optimiser state, odd step
        ck["g_opt"] = {"state": {}, "param_groups": [{"lr": 1e-4, "params": []}]}
        ck["d_opt"] = {"state": {}, "param_groups": [{"lr": 1e-4, "params": []}]}
        torch.save(ck, path)
        t2 = _trainer(tmp_path)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            t2.resume()
        assert sum("per-parameter optimiser state" in str(x.message) for x in w) == 2
        assert t2.logger.global_step == 8 and t2.logger.epoch == 2            # trainer.py:575-577
        t2.g_opt_step()                                                       # an extrapolation: would raise on an odd step
        assert t2.g_opt._have_copy


def test_resume_round_trips_our_own_optimizer_state(tmp_path):
    with noop_library():
        t = _trainer(tmp_path, tasks=("d", "s", "m"))
        t.g_opt.zero_grad()
        t.g_opt.extrapolation()
        t.logger.global_step = 1
        t.g_opt._flat[0]["m"].fill_(0.25)
        path = t.save()
        t2 = _trainer(tmp_path, tasks=("d", "s", "m"))
        t2.resume(checkpoint_path=path)
        assert t2.logger.global_step == 1 and t2.g_opt._have_copy and t2.g_opt._steps == 1
        assert float(t2.g_opt._flat[0]["m"][0]) == 0.25
        t2.g_opt_step()          # the update step that belongs to the restored look-ahead copy


def test_resume_merges_masker_and_painter_checkpoints(tmp_path):
    """trainer.py:447-477: a P+M model resumed from load_paths.m and load_paths.p (directories or .pth files)."""
    with noop_library():
        tm = _trainer(tmp_path / "m", tasks=("d", "s", "m"))
        tp = _trainer(tmp_path / "p", tasks=("p",))
        with torch.no_grad():
            next(tm.G.decoders["m"].parameters()).fill_(0.5)
            next(tp.G.painter.parameters()).fill_(-0.5)
        tm.save()
        pp = tp.save()
        t = _trainer(tmp_path / "pm")
        t.opts.load_paths = {"m": str(tmp_path / "m"), "p": str(pp), "pm": "none"}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t.resume(inference=True)
        assert float(next(t.G.decoders["m"].parameters()).flatten()[0]) == 0.5
        assert float(next(t.G.painter.parameters()).flatten()[0]) == -0.5
        t.opts.load_paths = {"m": str(tmp_path / "m"), "p": str(tmp_path / "m"), "pm": "none"}
        with pytest.raises(ValueError):
            t.resume(inference=True)


def test_vggloss_loads_torchvision_keys_and_warns_without_weights():
    from climategan_b200.losses import VGGLoss, Vgg19

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        loss = VGGLoss(storage_dtype=torch.float32)
    assert not loss.pretrained and any("RANDOMLY INITIALISED" in str(x.message) for x in w)
    # torchvision layout: features.{idx}.weight / bias
    ref = Vgg19(storage_dtype=torch.float32)
    tv = {"features." + k.split(".", 1)[1]: v.clone() + 1.0 for k, v in ref.state_dict().items()}
    tv["classifier.0.weight"] = torch.zeros(1)
    loss.load_vgg19_weights(tv)
    assert loss.pretrained
    for k, v in loss.vgg.state_dict().items():
        assert torch.equal(v, ref.state_dict()[k] + 1.0), k
    with pytest.raises(KeyError):
        loss.load_vgg19_weights({"features.0.weight": tv["features.0.weight"]})
