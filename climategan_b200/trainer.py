"""Trainer — the step-level surface of ``climategan/trainer.py`` for the painter task (``opts.tasks == ["p"]``):
``update_G`` (:989-1015), ``update_D`` (:1017-1032), ``get_G_loss`` (:1162-1182), ``get_painter_loss`` (:1256-1387),
``get_D_loss`` (:1034-1160, painter branch :1071-1107), ``g_opt_step``/``d_opt_step`` (:674-694), ``batch_to_device``
(:609-631), with the reference's ``multi_domain_batch`` dict contract and ``logger.losses.{gen,disc}`` keys.

Every array op runs through libcgb200; the `.item()` syncs the reference performs per loss term are kept lazy here
(``logger.losses`` stores device scalars; ``Trainer.losses_to_host()`` materialises them once).
The masker tasks (m, s, d) are not built yet and raise.
"""
from __future__ import annotations

import torch

from . import ops
from .discriminator import create_discriminator
from .generator import create_generator
from .losses import FeatMatchLoss, GANLoss, HingeLoss, VGGLoss
from .optim import get_optimizer
from .utils import Dict


def divide_pred(disc_output):
    """tutils.py:443-470."""
    if type(disc_output) == list:
        half1 = [[t[: t.size(0) // 2] for t in p] for p in disc_output]
        half2 = [[t[t.size(0) // 2:] for t in p] for p in disc_output]
    else:
        half1 = disc_output[: disc_output.size(0) // 2]
        half2 = disc_output[disc_output.size(0) // 2:]
    return half1, half2


def get_losses(opts, verbose=0, device=None, storage_dtype=torch.bfloat16):
    """losses.py:353-441, painter entries."""
    losses = {"G": {"a": {}, "p": {}, "tasks": {}}, "D": {"default": {}, "advent": {}}}
    if "p" in opts.tasks:
        losses["G"]["p"]["gan"] = (HingeLoss() if opts.gen.p.loss == "hinge"
                                   else GANLoss(use_lsgan=False, soft_shift=opts.dis.soft_shift, flip_prob=opts.dis.flip_prob))
        if opts.train.lambdas.G.p.vgg != 0:
            losses["G"]["p"]["vgg"] = VGGLoss(device, storage_dtype=storage_dtype)
        losses["G"]["p"]["featmatch"] = FeatMatchLoss()
        losses["D"]["p"] = losses["G"]["p"]["gan"]
    return losses


class _Logger:
    """Keeps the nested ``losses`` dict and ``global_step`` of ``climategan.logger.Logger`` (comet upload is out of scope)."""

    def __init__(self):
        self.losses = Dict(gen=Dict(), disc=Dict())
        self.global_step = 0

    def log_losses(self, model_to_update="G", mode="train"):
        return None


class Trainer:
    def __init__(self, opts, comet_exp=None, verbose=0, device=None, storage_dtype=torch.bfloat16):
        self.opts = opts
        self.verbose = verbose
        self.is_setup = False
        self.storage_dtype = storage_dtype
        self.device = device or torch.device("cuda:0")
        self.logger = _Logger()
        self.G = self.D = self.g_opt = self.d_opt = self.losses = None
        self.g_scheduler = self.d_scheduler = None
        self.kitti_pretrain = False
        if any(t in opts.tasks for t in "msd"):
            raise NotImplementedError("Trainer is built for the painter task only so far (tasks=['p'])")

    @property
    def has_painter(self):
        return "p" in self.opts.tasks

    def setup(self, inference=False, input_shape=(640, 640)):
        """trainer.py:702-770 without the data loaders (bench / tests feed tensors directly)."""
        self.G = create_generator(self.opts, device=self.device, latent_shape=tuple(input_shape),
                                  storage_dtype=self.storage_dtype)
        if not inference:
            self.D = create_discriminator(self.opts, self.device, storage_dtype=self.storage_dtype)
            self.g_opt, self.g_scheduler, self.lr_names_g = get_optimizer(self.G, self.opts.gen.opt, self.opts.tasks)
            self.d_opt, self.d_scheduler, self.lr_names_d = get_optimizer(self.D, self.opts.dis.opt, self.opts.tasks, True)
            self.losses = get_losses(self.opts, self.verbose, device=self.device, storage_dtype=self.storage_dtype)
            self.G.train()
            self.D.train()
        else:
            self.G.eval()
        self.is_setup = True
        return self

    # ---------------------------------------------------------------- data
    def batch_to_device(self, b):
        for task, tensor in b["data"].items():
            b["data"][task] = tensor.to(self.device, non_blocking=True)
        return b

    # ---------------------------------------------------------------- optimiser steps
    def _opt_step(self, opt, name):
        if "extra" in name.lower() and self.logger.global_step % 2 == 0:
            opt.extrapolation()
        else:
            opt.step()

    def g_opt_step(self):
        self._opt_step(self.g_opt, self.opts.gen.opt.optimizer)

    def d_opt_step(self):
        self._opt_step(self.d_opt, self.opts.dis.opt.optimizer)

    @staticmethod
    def _set_requires_grad(net, flag):
        for p in net.parameters():
            if p.dtype.is_floating_point and not (p.requires_grad is False and getattr(p, "_cgb_frozen", False)):
                pass
        # spectral-norm u/v are permanent non-trainable parameters: never flip them
        for name, p in net.named_parameters():
            if name.endswith(("weight_u", "weight_v")):
                continue
            p.requires_grad_(flag)

    # ---------------------------------------------------------------- update steps
    def update_G(self, multi_domain_batch, verbose=0):
        self._set_requires_grad(self.D, False)   # run_epoch freezes D around update_G (trainer.py:960-962)
        self.g_opt.zero_grad()
        g_loss = self.get_G_loss(multi_domain_batch, verbose)
        g_loss.backward()
        self.g_opt_step()
        self._set_requires_grad(self.D, True)    # trainer.py:971-973
        self.logger.log_losses(model_to_update="G", mode="train")
        return g_loss

    def update_D(self, multi_domain_batch, verbose=0):
        self.d_opt.zero_grad()
        d_loss = self.get_D_loss(multi_domain_batch, verbose)
        d_loss.backward()
        self.d_opt_step()
        self.logger.losses.disc.total_loss = d_loss.detach()
        self.logger.log_losses(model_to_update="D", mode="train")
        return d_loss

    def get_G_loss(self, multi_domain_batch, verbose=0):
        g_loss = 0
        if "p" in self.opts.tasks and not self.kitti_pretrain:
            p_loss = self.get_painter_loss(multi_domain_batch)
            self.logger.losses.gen.painter = p_loss.detach()
            g_loss = g_loss + p_loss
        assert not isinstance(g_loss, int), "No update in get_G_loss!"
        self.logger.losses.gen.total_loss = g_loss.detach()
        return g_loss

    def get_painter_loss(self, multi_domain_batch):
        """trainer.py:1256-1387 (vgg, gan, featmatch; tv/context/reconstruction have lambda 0 in defaults.yaml:294-299)."""
        step_loss = 0
        lambdas = self.opts.train.lambdas
        batch = multi_domain_batch["rf"]
        x = batch["data"]["x"]
        m = batch["data"]["m"]
        fake_flooded = self.G.paint(m, x)
        if lambdas.G.p.vgg != 0:
            loss = self.losses["G"]["p"]["vgg"](fake_flooded, x, m) * lambdas.G.p.vgg
            self.logger.losses.gen.p.vgg = loss.detach()
            step_loss = step_loss + loss
        for name in ("tv", "context", "reconstruction"):
            if lambdas.G.p[name] != 0:
                raise NotImplementedError(f"painter loss '{name}' (lambda 0 in defaults.yaml) is not built")
        if self.opts.gen.p.diff_aug.use:
            raise NotImplementedError("gen.p.diff_aug (off in defaults.yaml:158) is not built")
        real_cat = torch.cat([m, x], axis=1)
        fake_cat = ops.cat_mask_image(m, fake_flooded)
        real_fake_d = self.D["p"](torch.cat([real_cat, fake_cat], dim=0))
        real_d, fake_d = divide_pred(real_fake_d)
        loss = self.losses["G"]["p"]["gan"](fake_d, True, False)
        self.logger.losses.gen.p.gan = loss.detach()
        step_loss = step_loss + loss
        if self.opts.dis.p.get_intermediate_features and lambdas.G.p.featmatch != 0:
            loss = self.losses["G"]["p"]["featmatch"](real_d, fake_d) * lambdas.G.p.featmatch
            self.logger.losses.gen.p.featmatch = loss.detach()
            step_loss = step_loss + loss
        return step_loss

    def get_D_loss(self, multi_domain_batch, verbose=0):
        """trainer.py:1034-1160, painter branch."""
        disc_loss = {"p": {"gan": 0}}
        for domain, batch in multi_domain_batch.items():
            x = batch["data"]["x"]
            if domain == "rf" and self.has_painter:
                m = batch["data"]["m"]
                with torch.no_grad():
                    fake = self.G.paint(m, x)
                fake = fake.detach()
                real_cat = torch.cat([m, x], axis=1)
                fake_cat = torch.cat([m, fake], axis=1)
                real_fake_d = self.D["p"](torch.cat([real_cat, fake_cat], dim=0))
                real_d, fake_d = divide_pred(real_fake_d)
                disc_loss["p"]["gan"] = (self.losses["D"]["p"](fake_d, False, True)
                                         + self.losses["D"]["p"](real_d, True, True))
        self.logger.losses.disc.update({dom: {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
                                        for dom, d in disc_loss.items()})
        return sum(v for d in disc_loss.values() for v in d.values())

    def losses_to_host(self):
        """One sync for all logged scalars (the reference calls .item() ~10x per step)."""
        def conv(d):
            return {k: (conv(v) if isinstance(v, dict) else (float(v) if isinstance(v, torch.Tensor) else v)) for k, v in d.items()}
        return conv(self.logger.losses)
