"""Trainer — the step-level surface of ``climategan/trainer.py``: ``update_G`` (:989-1015), ``update_D`` (:1017-1032),
``get_G_loss`` (:1162-1182), ``get_masker_loss`` (:1184-1254) with ``masker_{d,s,m}_loss`` (:1389-1616) and
``painter_loss_for_masker`` (:1618-1651), ``get_painter_loss`` (:1256-1387), ``get_D_loss`` (:1034-1160),
``g_opt_step``/``d_opt_step`` (:674-694), ``batch_to_device`` (:609-631), ``infer_all`` (:218-334) and the event compositing
(:1824-1939), ``save`` / ``resume`` / ``resume_from_path`` — with the reference's ``multi_domain_batch`` dict contract and
``logger.losses.{gen,disc}`` keys, for tasks d, s, m (base or SPADE mask decoder, deeplabv2 / deeplabv3 encoder) and p.

Every array op runs through libcgb200; the `.item()` syncs the reference performs per loss term are kept lazy here
(``logger.losses`` stores device scalars; ``Trainer.losses_to_host()`` materialises them once).
"""
from __future__ import annotations

import torch

from . import events, ops
from .discriminator import create_discriminator
from .generator import create_generator
from .losses import (ADVENTAdversarialLoss, BCEWithLogits, CrossEntropy, DADADepthLoss, FeatMatchLoss, GANLoss, GroundIntersectionLoss,
                     HingeLoss, MinentLoss, SIGMLoss, TVLoss, VGGLoss)
from .discriminator import fc_discriminator_forward
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
        losses["G"]["p"]["tv"] = TVLoss()
        losses["D"]["p"] = losses["G"]["p"]["gan"]
    if "d" in opts.tasks:
        if opts.gen.d.classify.enable:
            losses["G"]["tasks"]["d"] = CrossEntropy()   # losses.py:399-405: bucketised log-depth, loss name ignored
        elif opts.gen.d.loss == "dada":
            losses["G"]["tasks"]["d"] = DADADepthLoss()
        else:
            losses["G"]["tasks"]["d"] = SIGMLoss(opts.train.lambdas.G.d.gml)
    if "s" in opts.tasks:
        losses["G"]["tasks"]["s"] = {"crossent": CrossEntropy(), "minent": MinentLoss(),
                                     "advent": ADVENTAdversarialLoss(opts, gan_type=opts.dis.s.gan_type)}
    if "m" in opts.tasks:
        losses["G"]["tasks"]["m"] = {
            "bce": BCEWithLogits(),
            "minent": (MinentLoss(version=2, lambda_var=opts.train.lambdas.advent.ent_var) if opts.gen.m.use_minent_var
                       else MinentLoss()),
            "tv": TVLoss(),
            "advent": ADVENTAdversarialLoss(opts, gan_type=opts.dis.m.gan_type),
            "gi": GroundIntersectionLoss(),
        }
    if "m" in opts.tasks or "s" in opts.tasks:
        losses["D"]["advent"] = ADVENTAdversarialLoss(opts)   # losses.py:440: always gan_type="GAN" (BCE), as the reference
    return losses


class _Logger:
    """Keeps the nested ``losses`` dict and ``global_step`` of ``climategan.logger.Logger`` (comet upload is out of scope)."""

    def __init__(self):
        self.losses = Dict(gen=Dict(), disc=Dict())
        self.global_step = 0
        self.epoch = 0

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
        self.kitti_pretrain = bool(opts.train.kitti.pretrain)   # trainer.py:101
        self.use_pl4m = False
        self.data_parallel = False
        self._use_graphs = False
        self._graphs, self._graph_seen, self._graph_pool = {}, {}, None
        self.pseudo_training_tasks = set(opts.train.pseudo.tasks or [])
        self.domain_labels = {"s": 0, "r": 1}

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
            if sum(p.numel() for p in self.D.parameters()) > 0:
                self.d_opt, self.d_scheduler, self.lr_names_d = get_optimizer(self.D, self.opts.dis.opt, self.opts.tasks, True)
            else:   # no adversarial loss configured (trainer.py:762-767): nothing to optimise, update_D is never called
                self.d_opt, self.d_scheduler = None, None
            self.losses = get_losses(self.opts, self.verbose, device=self.device, storage_dtype=self.storage_dtype)
            self.diff_transforms = None
            if "p" in self.opts.tasks and self.opts.gen.p.diff_aug.use:                # trainer.py:772-773
                from .transforms import DiffTransforms

                self.diff_transforms = DiffTransforms(self.opts.gen.p.diff_aug)
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
        """run_epoch's freeze / un-freeze of the discriminator (trainer.py:958-973): EVERY parameter, the spectral-norm u / v
        vectors included — after the first un-freeze they are trained by ExtraAdam like any weight (reference behaviour)."""
        for p in net.parameters():
            p.requires_grad_(flag)

    # ---------------------------------------------------------------- data parallelism
    def enable_data_parallel(self, group=None):
        """One process per GPU, each with its own slice of every domain batch: after each backward the flat G (or D) gradient
        buffer is averaged over the ranks (NCCL all-reduce), then every rank applies the same ExtraAdam update.
        BatchNorm statistics stay per-rank, as wrapping the reference in DDP (no SyncBN) would leave them."""
        self.data_parallel = True
        self._dp_group = group
        return self

    def _sync_grads(self, opt):
        if self.data_parallel:
            from .parallel import allreduce_flat_grads

            allreduce_flat_grads(opt, self._dp_group)

    # ---------------------------------------------------------------- CUDA-graph replay of forward + backward
    def enable_cuda_graphs(self, flag=True):
        """Replay ``zero_grad`` + loss + ``backward`` of update_G / update_D from a captured CUDA graph (graphs.py): the first
        call with a given batch signature runs eagerly (warm-up), the second captures, later ones replay.  The optimiser
        update and the data-parallel all-reduce stay eager.  Numerically identical to the eager step — same kernels, same
        order, and the host-side random draws (GANLoss labels, dropout seeds) are re-made before every replay in the eager
        order (tests/test_gpu_graphs.py).  A graph is keyed on everything that shapes the step's control flow: the batch's
        domains / tasks / shapes, pl4m, kitti pre-training, pseudo-label tasks, train / eval mode."""
        self._use_graphs = bool(flag)
        self._graphs, self._graph_seen = {}, {}
        return self

    def reset_graphs(self):
        """Drop every captured graph (call after changing options that the captured control flow depends on)."""
        self._graphs, self._graph_seen = {}, {}

    def _fwd_bwd(self, kind, data):
        """The captured region: data = {domain: {task: tensor}}."""
        mdb = {dom: {"data": d, "domain": [dom] * next(iter(d.values())).shape[0]} for dom, d in data.items()}
        if kind == "G":
            self.g_opt.zero_grad()
            loss = self.get_G_loss(mdb)
        else:
            self.d_opt.zero_grad()
            loss = self.get_D_loss(mdb)
        loss.backward()
        return loss.detach()

    def _loss_backward(self, kind, multi_domain_batch):
        data = {dom: {k: v for k, v in b["data"].items() if isinstance(v, torch.Tensor)} for dom, b in multi_domain_batch.items()}
        if not self._use_graphs:
            return self._fwd_bwd(kind, data)
        from . import graphs

        key = (kind, graphs.tree_signature(data), self.use_pl4m, self.kitti_pretrain, tuple(sorted(self.pseudo_training_tasks)),
               self.G.training, self.D.training if self.D is not None else None)
        seen = self._graph_seen.get(key, 0)
        self._graph_seen[key] = seen + 1
        if seen == 0:
            return self._fwd_bwd(kind, data)    # eager warm-up: lazy initialisation (flat buffers, kernel attributes) happens here
        step = self._graphs.get(key)
        if step is None:
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            step = self._graphs[key] = graphs.GraphedStep(lambda d, kind=kind: self._fwd_bwd(kind, d), data, self.device,
                                                          pool=self._graph_pool)
        return step(data)

    # ---------------------------------------------------------------- update steps
    def update_G(self, multi_domain_batch, verbose=0):
        self._set_requires_grad(self.D, False)   # run_epoch freezes D around update_G (trainer.py:960-962)
        g_loss = self._loss_backward("G", multi_domain_batch)
        self._sync_grads(self.g_opt)
        self.g_opt_step()
        self._set_requires_grad(self.D, True)    # trainer.py:971-973
        self.logger.log_losses(model_to_update="G", mode="train")
        return g_loss

    def update_D(self, multi_domain_batch, verbose=0):
        if self.d_opt is None:   # run_epoch only calls update_D when there is a discriminator optimiser (trainer.py:971)
            return None
        d_loss = self._loss_backward("D", multi_domain_batch)
        self._sync_grads(self.d_opt)
        self.d_opt_step()
        self.logger.losses.disc.total_loss = d_loss.detach()
        self.logger.log_losses(model_to_update="D", mode="train")
        return d_loss

    # ---------------------------------------------------------------- epoch driver
    def update_learning_rates(self):
        """trainer.py:696-700."""
        if self.g_scheduler is not None:
            self.g_scheduler.step()
        if self.d_scheduler is not None:
            self.d_scheduler.step()

    def run_epoch(self, train_loaders):
        """trainer.py:924-987 for an iterable of multi-batch tuples (what ``zip(*loaders["train"].values())`` yields, :633): per
        tuple — random domain order (tutils.shuffle_batch_tuple :318-328), batches to the device, update_G (with D frozen),
        update_D unless the VKITTI2 pre-training is on (:971), global_step += 1 — then the learning-rate schedulers.  Data
        loading, comet logging and the progress bar stay with the caller."""
        import numpy as np

        assert self.is_setup
        self.G.train()
        if self.D is not None:
            self.D.train()
        for multi_batch_tuple in train_loaders:
            assert isinstance(multi_batch_tuple, (tuple, list)) and len(multi_batch_tuple) > 0
            perm = np.random.permutation(len(multi_batch_tuple))
            multi_domain_batch = {multi_batch_tuple[i]["domain"][0]: self.batch_to_device(multi_batch_tuple[i]) for i in perm}
            self.update_G(multi_domain_batch)
            if self.d_opt is not None and not self.kitti_pretrain:
                self.update_D(multi_domain_batch)
            self.logger.global_step += 1
        if not self.kitti_pretrain:
            self.update_learning_rates()

    def train(self, train_loaders, epochs=None, on_epoch_end=None):
        """trainer.py:888-922: ``epochs`` (default opts.train.epochs) calls of :meth:`run_epoch` with the epoch-boundary switches
        of the reference — the painter loss for the masker turns on at epoch ``gen.p.pl4m_epoch`` (when ``gen.m.use_pl4m``), the
        VKITTI2 pre-training ends after ``train.kitti.epochs`` epochs, pseudo-label training after ``train.pseudo.epochs``.
        ``train_loaders``: an iterable of multi-batch tuples, or a callable ``epoch -> iterable`` (the caller swaps the kitti
        loaders for the base ones there: ``switch_data``, :817-845).  ``on_epoch_end(trainer)`` stands where the reference
        evaluates and saves."""
        assert self.is_setup
        n = self.opts.train.epochs if epochs is None else epochs
        for self.logger.epoch in range(self.logger.epoch, self.logger.epoch + n):
            if (self.logger.epoch == self.opts.gen.p.pl4m_epoch and "p" in self.opts.tasks and self.opts.gen.m.use_pl4m
                    and sum(p.numel() for p in self.G.painter.parameters()) > 0):
                self.use_pl4m = True
            self.run_epoch(train_loaders(self.logger.epoch) if callable(train_loaders) else train_loaders)
            if on_epoch_end is not None:
                on_epoch_end(self)
            if self.opts.train.kitti.epochs and self.logger.epoch == self.opts.train.kitti.epochs - 1:
                self.kitti_pretrain = False
            if self.opts.train.pseudo.epochs and self.logger.epoch == self.opts.train.pseudo.epochs - 1:
                self.pseudo_training_tasks = set()

    def get_G_loss(self, multi_domain_batch, verbose=0):
        """trainer.py:1162-1182."""
        g_loss = 0
        if any(t in self.opts.tasks for t in "msd"):
            m_loss = self.get_masker_loss(multi_domain_batch)
            self.logger.losses.gen.masker = m_loss.detach()
            g_loss = g_loss + m_loss
        if "p" in self.opts.tasks and not self.kitti_pretrain:
            p_loss = self.get_painter_loss(multi_domain_batch)
            self.logger.losses.gen.painter = p_loss.detach()
            g_loss = g_loss + p_loss
        assert not isinstance(g_loss, int), "No update in get_G_loss!"
        self.logger.losses.gen.total_loss = g_loss.detach()
        return g_loss

    # ---------------------------------------------------------------- masker
    def _advent_D(self, task):
        net = self.D[task]["Advent"]
        return lambda t: fc_discriminator_forward(net, t, self.storage_dtype)

    def get_masker_loss(self, multi_domain_batch):
        """trainer.py:1184-1254."""
        m_loss = 0
        for domain, batch in multi_domain_batch.items():
            if domain == "rf":
                continue
            x = batch["data"]["x"]
            z = self.G.encode(x)
            d_pred = s_pred = z_depth = None
            for task in ["d", "s", "m"]:
                if task not in batch["data"] or task not in self.opts.tasks:
                    continue
                target = batch["data"][task]
                if task == "d":
                    loss, d_pred, z_depth = self.masker_d_loss(x, z, target, domain, "G")
                    m_loss = m_loss + loss
                    self.logger.losses.gen.task["d"][domain] = loss.detach()
                elif task == "s":
                    loss, s_pred = self.masker_s_loss(x, z, d_pred, z_depth, target, domain, "G")
                    m_loss = m_loss + loss
                    self.logger.losses.gen.task["s"][domain] = loss.detach()
                elif task == "m":
                    cond = None
                    if self.opts.gen.m.use_spade:
                        # trainer.py:1235-1239 (the .clone() there only shields d_pred / s_pred from in-place edits)
                        cond = self.G.make_m_cond(d_pred, s_pred, x)
                    loss, _ = self.masker_m_loss(x, z, target, domain, "G", cond=cond, z_depth=z_depth, depth_preds=d_pred)
                    m_loss = m_loss + loss
                    self.logger.losses.gen.task["m"][domain] = loss.detach()
        return m_loss

    def _zero(self):
        return torch.zeros((), dtype=torch.float32, device=self.device)

    def masker_d_loss(self, x, z, target, domain, for_="G"):
        """trainer.py:1389-1407."""
        assert for_ in {"G", "D"}
        assert x.shape[0] == target.shape[0]
        weight = self.opts.train.lambdas.G.d.main
        prediction, z_depth = self.G.decode_d(z)
        if weight == 0 or (domain == "r" and "d" not in self.pseudo_training_tasks):
            return self._zero(), prediction, z_depth    # the reference evaluates the loss and discards it
        if self.opts.gen.d.classify.enable:
            target = target.squeeze(1)                  # trainer.py:1398-1399 (bucket indices [B,1,H,W] -> [B,H,W])
        full_loss = self.losses["G"]["tasks"]["d"](prediction, target) * weight
        return full_loss, prediction, z_depth

    def masker_s_loss(self, x, z, depth_preds, z_depth, target, domain, for_="G"):
        """trainer.py:1409-1516."""
        assert for_ in {"G", "D"}
        assert domain in {"r", "s"}
        full_loss = self._zero()
        softmax_preds = None
        pred = None
        lam = self.opts.train.lambdas
        if for_ == "G" or self.opts.gen.s.use_advent:
            pred = self.G.decode_s(z, z_depth)
        if for_ == "G":
            if domain == "s" or "s" in self.pseudo_training_tasks:
                key = "crossent" if domain == "s" else "crossent_pseudo"
                weight = lam.G["s"][key]
                if weight != 0:
                    loss = self.losses["G"]["tasks"]["s"]["crossent"](pred, target.squeeze(1)) * weight
                    full_loss = full_loss + loss
                    self.logger.losses.gen.task["s"][key][domain] = loss.detach()
            if domain == "r":
                weight = lam.G["s"]["minent"]
                if self.opts.gen.s.use_minent and weight != 0:
                    softmax_preds = ops.softmax_nchw(pred)
                    loss = self.losses["G"]["tasks"]["s"]["minent"](softmax_preds) * weight
                    full_loss = full_loss + loss
                    self.logger.losses.gen.task["s"]["minent"]["r"] = loss.detach()
        if self.opts.gen.s.use_advent:
            if self.opts.gen.s.use_dada and depth_preds is not None:
                depth_preds = depth_preds.detach()
            else:
                depth_preds = None
            if for_ == "D":
                domain_label = domain
                logger = {}
                loss_func = self.losses["D"]["advent"]
                pred = pred.detach()
                weight = lam.advent.adv_main
            else:
                domain_label = "s"
                logger = self.logger.losses.gen.task["s"]["advent"]
                loss_func = self.losses["G"]["tasks"]["s"]["advent"]
                weight = lam.G["s"]["advent"]
            if (for_ == "D" or domain == "r") and weight != 0:
                if softmax_preds is None:
                    softmax_preds = ops.softmax_nchw(pred)
                loss = loss_func(softmax_preds, self.domain_labels[domain_label], self._advent_D("s"), depth_preds) * weight
                full_loss = full_loss + loss
                logger[domain] = loss.detach()
                # trainer.py:1487: `gan_type == "GAN" or "WGAN_norm"` is always true -> never any clipping / gradient penalty
        return full_loss, pred

    def masker_m_loss(self, x, z, target, domain, for_="G", cond=None, z_depth=None, depth_preds=None):
        """trainer.py:1518-1616."""
        assert for_ in {"G", "D"}
        assert domain in {"r", "s"}
        full_loss = self._zero()
        lam = self.opts.train.lambdas
        pred_logits = self.G.decode_m(z, cond=cond, z_depth=z_depth)
        prob = ops.sigmoid_pair(pred_logits)            # cat[sigmoid(l), 1 - sigmoid(l)]
        pred_prob = prob[:, :1]
        if for_ == "G":
            weight = lam.G.m.tv
            if weight != 0:
                loss = self.losses["G"]["tasks"]["m"]["tv"](pred_prob) * weight
                full_loss = full_loss + loss
                self.logger.losses.gen.task["m"]["tv"][domain] = loss.detach()
            weight = lam.G.m.bce
            if domain == "s" and weight != 0:
                loss = self.losses["G"]["tasks"]["m"]["bce"](pred_logits, target) * weight
                full_loss = full_loss + loss
                self.logger.losses.gen.task["m"]["bce"]["s"] = loss.detach()
            if domain == "r":
                weight = lam.G["m"]["gi"]
                if self.opts.gen.m.use_ground_intersection and weight != 0:
                    loss = self.losses["G"]["tasks"]["m"]["gi"](pred_prob, target) * weight
                    full_loss = full_loss + loss
                    self.logger.losses.gen.task["m"]["gi"]["r"] = loss.detach()
                weight = lam.G.m.pl4m
                if self.use_pl4m and weight != 0:
                    loss = self.painter_loss_for_masker(x, pred_prob) * weight
                    full_loss = full_loss + loss
                    self.logger.losses.gen.task["m"]["pl4m"]["r"] = loss.detach()
                weight = lam.advent.ent_main
                if self.opts.gen.m.use_minent and weight != 0:
                    loss = self.losses["G"]["tasks"]["m"]["minent"](prob) * weight
                    full_loss = full_loss + loss
                    self.logger.losses.gen.task["m"]["minent"]["r"] = loss.detach()
        if self.opts.gen.m.use_advent:
            if self.opts.gen.m.use_dada and depth_preds is not None:
                dp = ops.to_storage(depth_preds.detach(), self.storage_dtype)
                depth_preds = ops.from_storage(ops.resize_nearest(dp, x.shape[-2], x.shape[-1]), 1)
            else:
                depth_preds = None
            if for_ == "D":
                domain_label = domain
                logger = {}
                loss_func = self.losses["D"]["advent"]
                prob = prob.detach()
                weight = lam.advent.adv_main
            else:
                domain_label = "s"
                logger = self.logger.losses.gen.task["m"]["advent"]
                loss_func = self.losses["G"]["tasks"]["m"]["advent"]
                weight = lam.advent.adv_main
            if (for_ == "D" or domain == "r") and weight != 0:
                loss = loss_func(prob, self.domain_labels[domain_label], self._advent_D("m"), depth_preds) * weight
                full_loss = full_loss + loss
                logger[domain] = loss.detach()
        return full_loss, prob

    def painter_loss_for_masker(self, x, m):
        """trainer.py:1618-1651 (pl4m; switched on at epoch gen.p.pl4m_epoch when gen.m.use_pl4m, trainer.py:899-909): the frozen
        painter paints with the masker's PREDICTED mask and the painter discriminator scores the result; the gradient reaches
        the masker through x (1 - m), the SPADE conditioning, the paste and the mask channel of D's input."""
        frozen = [p for p in self.G.painter.parameters() if p.requires_grad]
        for p in frozen:
            p.requires_grad = False
        try:
            fake_flooded = self.G.paint(m, x)
            if self.opts.dis.p.use_local_discriminator:                                # trainer.py:1628-1636
                fake_d_global = self.D["p"]["global"](fake_flooded)
                fake_d_local = self.D["p"]["local"](ops.paste(torch.zeros_like(x), m, fake_flooded))
                return (self.losses["G"]["p"]["gan"](fake_d_global, True, False)
                        + self.losses["G"]["p"]["gan"](fake_d_local, True, False))
            real_cat = torch.cat([m, x], axis=1)
            fake_cat = torch.cat([m, fake_flooded], axis=1)
            real_fake_d = self.D["p"](torch.cat([real_cat, fake_cat], dim=0))
            _, fake_d = divide_pred(real_fake_d)
            return self.losses["G"]["p"]["gan"](fake_d, True, False)
        finally:
            # (the reference re-enables EVERY painter parameter, u / v included — harmless there: g_opt only holds what was
            # trainable at construction; here the previous flags are restored)
            for p in frozen:
                p.requires_grad = True

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
        # tv / context / reconstruction have lambda 0 in defaults.yaml:294-299.  `t * mask` is the paste kernel on a zero
        # background (x (1 - m) + t m with x = 0): gradient to t, none to the mask, as in the reference where m is data.
        zeros = torch.zeros_like(x)
        mf = m.to(x.dtype)
        if lambdas.G.p.tv != 0:                                                       # trainer.py:1292-1296, losses.py:140-171
            loss = self.losses["G"]["p"]["tv"](ops.paste(zeros, mf, fake_flooded)) * lambdas.G.p.tv
            self.logger.losses.gen.p.tv = loss.detach()
            step_loss = step_loss + loss
        if lambdas.G.p.context != 0:                                                  # masked L1 off the water, losses.py:281-287
            loss = ops.l1_loss(ops.paste(zeros, 1.0 - mf, fake_flooded), ops.paste(zeros, 1.0 - mf, x).detach()) * lambdas.G.p.context
            self.logger.losses.gen.p.context = loss.detach()
            step_loss = step_loss + loss
        if lambdas.G.p.reconstruction != 0:                                           # masked L1 on the water, losses.py:290-296
            loss = ops.l1_loss(ops.paste(zeros, mf, fake_flooded), ops.paste(zeros, mf, x).detach()) * lambdas.G.p.reconstruction
            self.logger.losses.gen.p.reconstruction = loss.detach()
            step_loss = step_loss + loss
        if self.opts.gen.p.diff_aug.use:   # trainer.py:1319-1321: both D inputs, independent draws; the mask channel is not moved
            fake_flooded = self.diff_transforms(fake_flooded)
            x = self.diff_transforms(x)
        if self.opts.dis.p.use_local_discriminator:                                   # trainer.py:1322-1356
            fake_d_global = self.D["p"]["global"](fake_flooded)
            fake_d_local = self.D["p"]["local"](ops.paste(zeros, mf, fake_flooded))
            real_d_global = self.D["p"]["global"](x)
            loss = (self.losses["G"]["p"]["gan"](fake_d_global, True, False)
                    + self.losses["G"]["p"]["gan"](fake_d_local, True, False)) * lambdas.G["p"]["gan"]
            self.logger.losses.gen.p.gan = loss.detach()
            step_loss = step_loss + loss
            if self.opts.dis.p.get_intermediate_features:                             # (on the global discriminator only)
                loss = self.losses["G"]["p"]["featmatch"](real_d_global, fake_d_global) * lambdas.G["p"]["featmatch"]
                self.logger.losses.gen.p.featmatch = loss.detach() if isinstance(loss, torch.Tensor) else loss
                step_loss = step_loss + loss
            return step_loss
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
        """trainer.py:1034-1160."""
        disc_loss = {"m": {"Advent": 0}, "s": {"Advent": 0},
                     "p": {"global": 0, "local": 0} if self.opts.dis.p.use_local_discriminator else {"gan": 0}}   # trainer.py:1058-1066
        lam = self.opts.train.lambdas
        for domain, batch in multi_domain_batch.items():
            x = batch["data"]["x"]
            if domain == "rf" and self.has_painter:
                m = batch["data"]["m"]
                with torch.no_grad():
                    fake = self.G.paint(m, x)
                    if self.opts.gen.p.diff_aug.use:                                   # trainer.py:1079-1081
                        fake = self.diff_transforms(fake)
                        x = self.diff_transforms(x)
                fake = fake.detach()
                if self.opts.dis.p.use_local_discriminator:                            # trainer.py:1084-1098
                    zeros, mf = torch.zeros_like(x), m.to(x.dtype)
                    Dp, Lp = self.D["p"], self.losses["D"]["p"]
                    disc_loss["p"]["global"] = (Lp(Dp["global"](fake), False, True) + Lp(Dp["global"](x), True, True))
                    disc_loss["p"]["local"] = (Lp(Dp["local"](ops.paste(zeros, mf, fake)), False, True)
                                               + Lp(Dp["local"](ops.paste(zeros, mf, x)), True, True))
                    continue
                real_cat = torch.cat([m, x], axis=1)
                fake_cat = torch.cat([m, fake], axis=1)
                real_fake_d = self.D["p"](torch.cat([real_cat, fake_cat], dim=0))
                real_d, fake_d = divide_pred(real_fake_d)
                disc_loss["p"]["gan"] = (self.losses["D"]["p"](fake_d, False, True)
                                         + self.losses["D"]["p"](real_d, True, True))
            elif domain != "rf" and any(t in self.opts.tasks for t in "msd"):
                # every generator output is detached before it reaches a discriminator (trainer.py:1467,1585), so the
                # generator runs without an autograd tape here; BatchNorm still runs on (and updates) batch statistics.
                with torch.no_grad():
                    z = self.G.encode(x)
                    s_pred = d_pred = cond = z_depth = None
                    if "s" in batch["data"] and "s" in self.opts.tasks:
                        if "d" in self.opts.tasks and self.opts.gen.s.use_dada:
                            d_pred, z_depth = self.G.decode_d(z)
                if "s" in batch["data"] and "s" in self.opts.tasks:
                    step_loss, s_pred = self._no_g_tape(self.masker_s_loss, x, z, d_pred, z_depth, None, domain, for_="D")
                    disc_loss["s"]["Advent"] = disc_loss["s"]["Advent"] + step_loss * lam.advent.adv_main
                if "m" in batch["data"] and "m" in self.opts.tasks:
                    if "d" in self.opts.tasks and (self.opts.gen.m.use_spade or self.opts.gen.m.use_dada):   # trainer.py:1128-1136
                        with torch.no_grad():
                            if d_pred is None:
                                d_pred, z_depth = self.G.decode_d(z)
                            if self.opts.gen.m.use_spade:
                                cond = self.G.make_m_cond(d_pred, s_pred, x)
                    step_loss, _ = self._no_g_tape(self.masker_m_loss, x, z, None, domain, for_="D", cond=cond,
                                                   z_depth=z_depth, depth_preds=d_pred)
                    disc_loss["m"]["Advent"] = disc_loss["m"]["Advent"] + step_loss * lam.advent.adv_main
        self.logger.losses.disc.update({dom: {k: (v.detach() if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
                                        for dom, d in disc_loss.items()})
        return sum(v for d in disc_loss.values() for v in d.values())

    def _no_g_tape(self, fn, *args, **kwargs):
        """Run a masker_*_loss(for_="D") with the generator's decoders un-taped (their outputs are detached by the loss)."""
        G = self.G
        orig = G._grad_ctx
        G._grad_ctx = torch.no_grad
        try:
            return fn(*args, **kwargs)
        finally:
            G._grad_ctx = orig

    # ---------------------------------------------------------------- inference (trainer.py:218-334, 1824-1939)
    def infer_all(self, x, numpy=True, stores={}, bin_value=-1, half=False, xla=False, cloudy=False, auto_resize_640=False,
                  ignore_event=set(), return_masks=False):
        """Dictionary of events ("flood", "wildfire", "smog") from a numpy array or tensor, single image or batch, HWC or CHW.
        ``half=True`` (trainer.py:263-264, ``apply_events.py --half``): the generator runs with fp16 activation / weight storage
        (tcgen05 ``kind::f16`` on fp16 operands, fp32 accumulation and statistics) for this call, whatever the trainer's
        ``storage_dtype`` is; the compositing stays fp32 as in the reference (events are computed on ``.float()`` tensors)."""
        import numpy as np

        if half and self.G.storage_dtype != torch.float16:
            saved = (self.G.storage_dtype, getattr(self.G.painter, "storage_dtype", None))
            self.G.storage_dtype = torch.float16
            if saved[1] is not None:
                self.G.painter.storage_dtype = torch.float16
            try:
                return self.infer_all(x, numpy, stores, bin_value, True, xla, cloudy, auto_resize_640, ignore_event, return_masks)
            finally:
                self.G.storage_dtype = saved[0]
                if saved[1] is not None:
                    self.G.painter.storage_dtype = saved[1]

        assert self.is_setup
        assert len(x.shape) in {3, 4}, f"Unknown Data shape {x.shape}"
        if not isinstance(x, torch.Tensor):
            x = torch.tensor(x, device=self.device)
        if len(x.shape) == 3:
            x = x.unsqueeze(0)
        if x.shape[1] != 3:
            assert x.shape[-1] == 3, f"Unknown x shape to permute {x.shape}"
            x = x.permute(0, 3, 1, 2)
        x = x.to(self.device).float().contiguous()
        if auto_resize_640 and (x.shape[-1] != 640 or x.shape[-2] != 640):
            xs = ops.resize_bilinear(ops.to_storage(x, torch.float32), 640, 640, align_corners=False)
            x = ops.from_storage(xs, 3)
        if xla:
            raise NotImplementedError("xla=True has no meaning here (sm_100a only)")
        with torch.no_grad():
            if self.has_painter:
                self.G.painter.set_latent_shape(x.shape, True)
            z = self.G.encode(x)
            depth, z_depth = self.G.decode_d(z)
            segmentation = self.G.decode_s(z, z_depth)
            cond = self.G.make_m_cond(depth, segmentation, x) if self.opts.gen.m.use_spade else None
            mask = self.G.mask(z=z, cond=cond, z_depth=z_depth)
            wildfire = smog = flood = None
            if "wildfire" not in ignore_event:
                wildfire = self.compute_fire(x, seg_preds=segmentation)
            if "smog" not in ignore_event:
                smog = self.compute_smog(x, d=depth, s=segmentation)
            if "flood" not in ignore_event:
                flood = self.compute_flood(x, m=mask, s=segmentation, cloudy=cloudy, bin_value=bin_value)
            m8 = events.mask_to_uint8(mask, bin_value) if return_masks else None
            if numpy:
                # normalize -> NHWC -> uint8 on the device, then asynchronous D2H copies into pinned staging buffers and ONE
                # synchronisation for all events (the reference does .cpu() per event, each a blocking pageable copy)
                dev8 = {"flood": flood, "smog": smog, "wildfire": wildfire}
                dev8 = {k: events.to_uint8_nhwc(v) for k, v in dev8.items() if v is not None}
                if m8 is not None:
                    dev8["mask"] = m8
                host8 = {k: self._pinned(k, v).copy_(v, non_blocking=True) for k, v in dev8.items()}
                torch.cuda.current_stream().synchronize()
                arr = {k: v.numpy().copy() for k, v in host8.items()}
                flood, smog, wildfire = arr.get("flood"), arr.get("smog"), arr.get("wildfire")
                m8 = arr.get("mask")
            output_data = {"flood": flood, "wildfire": wildfire, "smog": smog}
            if return_masks:
                output_data["mask"] = m8 if numpy else m8.cpu().numpy().astype(np.uint8)
        return output_data

    def _pinned(self, key, like):
        """Pinned host staging buffer of ``like``'s shape / dtype, cached per output (cudaHostAlloc is expensive)."""
        cache = self.__dict__.setdefault("_pinned_cache", {})
        buf = cache.get(key)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = cache[key] = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
        return buf

    def compute_fire(self, x, seg_preds=None, z=None, z_depth=None):
        """trainer.py:1824-1841."""
        if seg_preds is None:
            if z is None:
                z = self.G.encode(x)
            seg_preds = self.G.decode_s(z, z_depth)
        return events.add_fire(x, seg_preds, self.opts.events.fire)

    def compute_flood(self, x, z=None, z_depth=None, m=None, s=None, cloudy=None, bin_value=-1):
        """trainer.py:1843-1877."""
        if m is None:
            if z is None:
                z = self.G.encode(x)
            if "d" in self.opts.tasks and self.opts.gen.m.use_dada and z_depth is None:
                _, z_depth = self.G.decode_d(z)
            m = self.G.mask(x=x, z=z, z_depth=z_depth)     # trainer.py:1866 passes x: the SPADE masker's conditioning needs it
        if bin_value >= 0:
            m = (m > bin_value).to(m.dtype)
        with torch.no_grad():
            if cloudy:
                assert s is not None
                return self.G.paint_cloudy(m, x, s)
            return self.G.paint(m, x)

    def compute_smog(self, x, z=None, d=None, s=None, use_sky_seg=False):
        """trainer.py:1879-1939 (use_sky_seg is a TODO in the reference: sky_mask stays None)."""
        if d is None:
            if z is None:
                z = self.G.encode(x)
            d, _ = self.G.decode_d(z)
        return events.add_smog(x, d, self.opts.events.smog)

    def paint_and_mask(self, x):
        """Mask then paint (the flood event without the other two)."""
        return self.compute_flood(x)

    # ---------------------------------------------------------------- validation metrics (trainer.py:1706-1799)
    def eval_images(self, mode, domain):
        """trainer.py:1706-1799: accuracy and mIOU of the segmentation (and of the bucketised depth on the sim domain) over
        ``self.display_images[mode][domain]``, one image at a time as in the reference.  Predictions stay on the device: each
        (prediction, label) pair costs one pass over the logits (``eval_metrics.accuracy_and_mIOU``) and one 1 KB copy.
        The mask block is guarded by ``"m" in self.opts`` exactly as the reference writes it (:1753) — false for the stock option
        tree, whose top level has no such key, so the reference reports -1 for both mask metrics; same here."""
        import numpy as np

        from .eval_metrics import accuracy_and_mIOU

        display_images = getattr(self, "display_images", None) or {}
        if domain == "s" and self.kitti_pretrain:
            domain = "kitti"
        if domain == "rf" or domain not in display_images.get(mode, {}):
            return
        scores = {"m": {}}
        if "s" in self.opts.tasks:
            scores["s"] = {}
        if "d" in self.opts.tasks and domain == "s" and self.opts.gen.d.classify.enable:
            scores["d"] = {}
        for task in scores:
            scores[task] = {"accuracy": [], "mIOU": []}

        def record(task, pred, label):
            acc, miou = accuracy_and_mIOU(pred, label)
            scores[task]["accuracy"].append(acc)
            scores[task]["mIOU"].append(miou)

        with torch.no_grad():
            for im_set in display_images[mode][domain]:
                x = im_set["data"]["x"].unsqueeze(0).to(self.device).float().contiguous()
                z = self.G.encode(x)
                s_pred = d_pred = z_depth = None
                if "d" in scores:
                    d_pred, z_depth = self.G.decode_d(z)
                    if domain == "s":
                        record("d", d_pred, im_set["data"]["d"].unsqueeze(0))
                if "s" in scores:
                    if z_depth is None and self.opts.gen.s.use_dada and "d" in self.opts.tasks:
                        _, z_depth = self.G.decode_d(z)
                    s_pred = self.G.decode_s(z, z_depth)
                    record("s", s_pred, im_set["data"]["s"].unsqueeze(0))
                if "m" in self.opts:
                    cond = self.G.make_m_cond(d_pred, s_pred, x) if s_pred is not None and d_pred is not None else None
                    if z_depth is None and self.opts.gen.m.use_dada and "d" in self.opts.tasks:
                        _, z_depth = self.G.decode_d(z)
                    pred_mask = (self.G.mask(z=z, cond=cond, z_depth=z_depth) > 0.5).float()
                    m = im_set["data"]["m"].unsqueeze(0)
                    acc, _ = accuracy_and_mIOU(pred_mask, m)                                     # one channel: see accuracy()
                    _, miou = accuracy_and_mIOU(torch.cat([1 - pred_mask, pred_mask], dim=1), m)
                    scores["m"]["accuracy"].append(acc)
                    scores["m"]["mIOU"].append(miou)
        scores = {task: {k: (float(np.mean(v)) if v else float("nan")) for k, v in met.items()} for task, met in scores.items()}
        scores = {task: {k: (v if not np.isnan(v) else -1) for k, v in met.items()} for task, met in scores.items()}
        flat = {f"{task}.{k}": v for task, met in scores.items() for k, v in met.items()}
        self.metrics = getattr(self, "metrics", {})
        self.metrics[f"metrics_{mode}_{domain}"] = flat
        exp = getattr(self, "exp", None)
        if exp is not None:
            exp.log_metrics(flat, prefix=f"metrics_{mode}_{domain}", step=self.logger.global_step)
        elif self.verbose > 0:
            print(f"metrics_{mode}_{domain}")
            print(flat)
        return 0

    # ---------------------------------------------------------------- checkpoints (trainer.py:337-412, 470-600)
    def save(self):
        """``{output_path}/checkpoints/latest_ckpt.pth`` with the reference's keys: G, g_opt, D, d_opt, epoch, step."""
        from pathlib import Path

        save_dir = Path(self.opts.output_path) / "checkpoints"
        save_dir.mkdir(parents=True, exist_ok=True)
        d = {"epoch": getattr(self.logger, "epoch", 0), "step": self.logger.global_step, "G": self.G.state_dict()}
        if self.g_opt is not None:
            d["g_opt"] = self.g_opt.state_dict()
        if self.D is not None and sum(p.numel() for p in self.D.parameters()) > 0:
            d["D"] = self.D.state_dict()
            if self.d_opt is not None:
                d["d_opt"] = self.d_opt.state_dict()
        torch.save(d, save_dir / "latest_ckpt.pth")
        return save_dir / "latest_ckpt.pth"

    @staticmethod
    def _merge(source, destination):
        """climategan/utils.py:68-105: recursive dict merge (source wins)."""
        for key, value in source.items():
            if isinstance(value, dict):
                Trainer._merge(value, destination.setdefault(key, {}))
            else:
                destination[key] = value
        return destination

    def _checkpoint_from_load_paths(self, checkpoint_path=None):
        """trainer.py:422-533: which file(s) to load — ``checkpoint_path`` (ours), else ``opts.load_paths.{m,p,pm}`` with the
        reference's rules (a P+M model may be assembled from a masker and a painter checkpoint: their dicts are merged), else
        ``{output_path}/checkpoints/latest_ckpt.pth``.  A directory stands for its ``checkpoints/latest_ckpt.pth``."""
        from pathlib import Path

        def as_file(p):
            p = Path(p)
            assert p.exists(), f"{p} does not exist"
            p = p / "checkpoints/latest_ckpt.pth" if p.is_dir() else p
            assert p.suffix == ".pth", p
            return p

        def load(p):
            return torch.load(as_file(p), map_location=self.device)

        if checkpoint_path is not None:
            return load(checkpoint_path)
        lp = self.opts.load_paths if "load_paths" in self.opts else {}
        m_path, p_path, pm_path = (str(lp.get(k, "none") or "none") for k in ("m", "p", "pm"))
        latest = Path(self.opts.output_path) / "checkpoints" / "latest_ckpt.pth"
        if "m" in self.opts.tasks and "p" in self.opts.tasks:
            if m_path == p_path == pm_path == "none":
                return load(latest)
            if pm_path != "none":
                return load(pm_path)
            if m_path != p_path:
                print(f"Resuming P+M model from \n  -{p_path} \nand \n  -{m_path}")
                return self._merge(load(m_path), load(p_path))
            raise ValueError("Cannot resume a P+M model with provided load_paths:\n{}".format(dict(lp)))
        if m_path != "none" and p_path != "none":
            raise ValueError("Opts tasks are {} but received 2 values for the load_paths".format(self.opts.tasks))
        if m_path != "none":
            assert "m" in self.opts.tasks
            return load(m_path)
        if p_path != "none":
            assert "p" in self.opts.tasks
            return load(p_path)
        return load(latest)

    def resume(self, inference=False, checkpoint_path=None):
        """trainer.py:422-588.  The model state_dicts of a reference checkpoint load as they are (same keys and shapes).  The
        OPTIMISER state only round-trips between runs of this package: ExtraAdam keeps its moments in one flat buffer per
        parameter group (optim.py), a layout the reference's per-parameter state cannot be poured into — such a state is skipped
        WITH a warning and the moments restart from zero (a plain ``torch.optim.Adam`` state loads normally).  As in the
        reference, an odd step is rounded up to an even one so that extragradient resumes on an extrapolation."""
        import warnings

        ckpt = self._checkpoint_from_load_paths(checkpoint_path)
        if inference:
            bad = self.G.load_state_dict(ckpt["G"], strict=False)
            if bad.missing_keys:
                print("WARNING: Missing keys in self.G.load_state_dict, keeping inits", bad.missing_keys)
            if bad.unexpected_keys:
                print("WARNING: Ignoring Unexpected keys in self.G.load_state_dict", bad.unexpected_keys)
            return self
        self.G.load_state_dict(ckpt["G"])
        if "D" in ckpt and self.D is not None and sum(p.numel() for p in self.D.parameters()) > 0:
            self.D.load_state_dict(ckpt["D"])
        from .optim import ExtraAdam

        restored_lookahead = {}
        for opt, key in ((self.g_opt, "g_opt"), (self.d_opt, "d_opt")):
            if opt is None:
                continue
            state = ckpt.get(key)
            if not isinstance(state, dict):
                warnings.warn(f"resume: the checkpoint holds no '{key}': the optimiser restarts from zero moments")
            elif isinstance(opt, ExtraAdam):
                if "flat" in state:
                    opt.load_state_dict(state)
                    restored_lookahead[key] = bool(opt._have_copy)
                else:
                    warnings.warn(f"resume: '{key}' is a per-parameter optimiser state (a reference checkpoint); ExtraAdam here keeps "
                                  "flat per-group moments and cannot load it — moments and step count restart from zero")
            else:
                try:
                    opt.load_state_dict(state)                       # torch.optim.Adam & co: the reference's own format
                except (ValueError, KeyError) as e:
                    warnings.warn(f"resume: could not restore '{key}' ({e}); the optimiser restarts from zero moments")
        self.logger.epoch = int(ckpt.get("epoch", 0))
        self.logger.global_step = int(ckpt.get("step", 0))
        for _ in range(self.logger.epoch + 1):                           # trainer.py:553-555: replay the schedulers
            self.update_learning_rates()
        if self.logger.global_step % 2 != 0 and not any(restored_lookahead.values()):
            self.logger.global_step += 1                                 # trainer.py:575-577: round to even for extragradient
        for opt in (self.g_opt, self.d_opt):                             # a look-ahead copy without its parity makes no sense
            if isinstance(opt, ExtraAdam) and self.logger.global_step % 2 == 0:
                opt._have_copy = False
        ops.invalidate_weight_cache()
        return self

    @classmethod
    def resume_from_path(cls, path, overrides={}, setup=True, inference=False, new_exp=False, device=None, verbose=1,
                         storage_dtype=torch.bfloat16, input_shape=(640, 640)):
        """trainer.py:337-396: ``path`` holds ``opts.yaml`` and ``checkpoints/latest_ckpt.pth`` (comet re-attachment is out of
        scope)."""
        from pathlib import Path

        import yaml

        p = Path(path).expanduser().resolve()
        assert p.exists() and (p / "checkpoints").is_dir(), f"{p} must contain checkpoints/"
        cands = sorted(p.glob("opts*.yaml"))
        assert cands, f"no opts*.yaml in {p}"
        with open(cands[-1]) as f:
            opts = Dict(yaml.safe_load(f))
        opts.update(overrides or {})
        opts.output_path = str(p)
        opts.train.resume = True
        t = cls(opts, device=device, verbose=verbose, storage_dtype=storage_dtype)
        if setup:
            t.setup(inference=inference, input_shape=input_shape)
            t.resume(inference=inference)
        return t

    def losses_to_host(self):
        """One sync for all logged scalars (the reference calls .item() ~10x per step)."""
        def conv(d):
            return {k: (conv(v) if isinstance(v, dict) else (float(v) if isinstance(v, torch.Tensor) else v)) for k, v in d.items()}
        return conv(self.logger.losses)
