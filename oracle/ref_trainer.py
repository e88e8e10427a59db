"""Build the UNMODIFIED reference ``climategan.trainer.Trainer`` on CPU (build container only) — the recipe of
SURVEY.md Appendix B: import through oracle/refshim.py, assemble ``G / D / g_opt / d_opt / losses`` by hand exactly as
``Trainer.setup`` does (trainer.py:721-770; setup itself needs dataset json files), and patch the hard-coded CUDA
assumptions listed in SURVEY.md §8c so the step runs on CPU.  Used to GENERATE golden vectors (tests/golden/make_golden.py)
and to pin the oracle restatements.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import torch

from oracle import refshim


from climategan_b200.utils import full_opts, synth_batch  # noqa: E402,F401  (shared with the product's bench / tests)


def build_reference_trainer(opts, size, vgg_seed=13, inference=False):
    """Returns the reference Trainer with G, D, optimisers and losses assembled (CPU, train mode, dropout disabled)."""
    trainer_mod, generator_mod, disc_mod, losses_mod, optim_mod, tutils_mod = refshim.load(
        "trainer", "generator", "discriminator", "losses", "optim", "tutils")
    import torchvision

    # ---- CPU / offline patches of hard-coded assumptions ----
    _orig_vgg19 = getattr(torchvision.models.vgg19, "_cgb_orig", torchvision.models.vgg19)   # (idempotent: this runs per trainer)

    def _vgg19_no_download(pretrained=True):
        return _orig_vgg19(weights=None)

    _vgg19_no_download._cgb_orig = _orig_vgg19
    losses_mod.models.vgg19 = _vgg19_no_download

    def vgg_preprocess_cpu(batch):
        (r, g, b) = torch.chunk(batch, 3, dim=1)
        batch = torch.cat((b, g, r), dim=1)
        batch = (batch + 1) * 255 * 0.5
        mean = torch.zeros_like(batch)
        mean[:, 0], mean[:, 1], mean[:, 2] = 103.939, 116.779, 123.680
        return batch.sub(mean)

    trainer_mod.vgg_preprocess = vgg_preprocess_cpu
    _sigm_init = losses_mod.SIGMLoss.__init__

    def sigm_init(self, gmweight=0.5, scale=4, device="cpu"):
        _sigm_init(self, gmweight, scale, "cpu")

    losses_mod.SIGMLoss.__init__ = sigm_init

    def custom_bce_call(self, prediction, target):  # losses.py:472-477 without .get_device() (-1 on CPU)
        return self.loss(prediction, torch.zeros_like(prediction).fill_(target))

    losses_mod.CustomBCELoss.__call__ = custom_bce_call

    class _CpuTimer:  # utils.Timer defaults to cuda=True and records CUDA events (utils.py:919-959): patch (5) of SURVEY.md §8c
        def __init__(self, name="", store=None, precision=3, ignore=False, cuda=False):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    trainer_mod.Timer = _CpuTimer
    dev = torch.device("cpu")
    t = trainer_mod.Trainer(opts, device=dev)
    t.G = generator_mod.create_generator(opts, device=dev, no_init=True)
    t.has_painter = "p" in opts.tasks
    if t.has_painter:
        t.G.painter.set_latent_shape(size, True)
    if inference:
        t.is_setup = True
        t.G.eval()
        return t
    t.D = disc_mod.create_discriminator(opts, dev, no_init=True)
    t.g_opt, t.g_scheduler, t.lr_names["G"] = optim_mod.get_optimizer(t.G, opts.gen.opt, opts.tasks)
    t.d_opt, t.d_scheduler, t.lr_names["D"] = optim_mod.get_optimizer(t.D, opts.dis.opt, opts.tasks, True)
    t.losses = losses_mod.get_losses(opts, 0, device=dev)
    t.is_setup = True
    t.G.train()
    t.D.train()
    for m in t.G.modules():  # parity runs: dropout off (its RNG stream cannot be shared), everything else in train mode
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return t


def load_weights(t, seeds=(21, 22, 23), d=True):
    from tests.golden.weights import fill_state_dict

    g_shapes = [(k, tuple(v.shape)) for k, v in t.G.state_dict().items()]
    t.G.load_state_dict(fill_state_dict(g_shapes, seeds[0]), strict=True)
    if not d:
        return g_shapes, None, None
    d_shapes = [(k, tuple(v.shape)) for k, v in t.D.state_dict().items()]
    t.D.load_state_dict(fill_state_dict(d_shapes, seeds[1]), strict=True)
    v_shapes = None
    if "p" in t.opts.tasks and "vgg" in t.losses["G"]["p"]:
        vgg = t.losses["G"]["p"]["vgg"].vgg
        v_shapes = [(k, tuple(v.shape)) for k, v in vgg.state_dict().items()]
        vgg.load_state_dict(fill_state_dict(v_shapes, seeds[2]), strict=True)
    return g_shapes, d_shapes, v_shapes


def run_steps(t, batch, n_iters=2):
    """What Trainer.run_epoch does per batch (trainer.py:946-980): freeze D, update_G, unfreeze D, update_D, step += 1.
    Returns the per-iteration loss logs (plain floats)."""
    import copy

    logs = []
    for _ in range(n_iters):
        for p in t.D.parameters():
            p.requires_grad = False
        t.update_G(batch)
        for p in t.D.parameters():
            p.requires_grad = True
        t.update_D(batch)
        t.logger.global_step += 1
        logs.append(copy.deepcopy(t.logger.losses.to_dict() if hasattr(t.logger.losses, "to_dict") else dict(t.logger.losses)))
    return logs
