"""Oracle for the painter train step: functional restatement of

  climategan/trainer.py   get_painter_loss :1256-1387 (vgg + gan + featmatch), get_D_loss painter branch :1071-1107,
                          g_opt_step/d_opt_step :674-694
  climategan/losses.py    Vgg19 :304-336, VGGLoss :338-350
  climategan/tutils.py    vgg_preprocess :416-427
  climategan/optim.py     ExtraAdam.update :242-291, Extragradient.extrapolation/step :153-197

built on oracle/painter_oracle.py and oracle/discriminator_oracle.py.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import discriminator_oracle as do
from oracle import painter_oracle as po

VGG_CONVS = {1: [0], 2: [2, 5], 3: [7, 10], 4: [12, 14, 16, 19], 5: [21, 23, 25, 28]}
VGG_POOL_BEFORE = {5, 10, 19, 28}  # conv indices preceded by a MaxPool2d(2,2) in torchvision's vgg19.features


def vgg_preprocess(batch):
    r, g, b = torch.chunk(batch, 3, dim=1)
    batch = torch.cat((b, g, r), dim=1)
    batch = (batch + 1) * 255 * 0.5
    mean = torch.tensor([103.939, 116.779, 123.680], dtype=batch.dtype).view(1, 3, 1, 1)
    return batch - mean


def vgg_features(vsd, x):
    outs = []
    for k in range(1, 6):
        for idx in VGG_CONVS[k]:
            if idx in VGG_POOL_BEFORE:
                x = F.max_pool2d(x, 2, 2)
            x = F.relu(F.conv2d(x, vsd[f"slice{k}.{idx}.weight"], vsd[f"slice{k}.{idx}.bias"], padding=1))
        outs.append(x)
    return outs


def vgg_loss(vsd, x, y):
    weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
    xv, yv = vgg_features(vsd, x), vgg_features(vsd, y)
    return sum(w * F.l1_loss(a, b.detach()) for w, a, b in zip(weights, xv, yv))


def painter_g_loss(gsd, dsd, vsd, x, m, z, lam_vgg=10.0, lam_feat=10.0, g_sn=None, d_sn=None):
    fake = po.paint(gsd, m, x, z, z, po.n_up_spades_of(gsd), sn=g_sn)
    terms = {}
    terms["vgg"] = lam_vgg * vgg_loss(vsd, vgg_preprocess(fake * m), vgg_preprocess(x * m))
    real_cat = torch.cat([m, x], 1)
    fake_cat = torch.cat([m, fake], 1)
    out = do.multiscale_forward(dsd, torch.cat([real_cat, fake_cat], 0), sn=d_sn)
    real_d, fake_d = do.divide_pred(out)
    terms["gan"] = do.gan_loss(fake_d, True)
    terms["featmatch"] = lam_feat * do.feat_match_loss(real_d, fake_d)
    return sum(terms.values()), terms


def painter_d_loss(gsd, dsd, x, m, z, g_sn=None, d_sn=None):
    with torch.no_grad():
        fake = po.paint(gsd, m, x, z, z, po.n_up_spades_of(gsd), sn=g_sn)
    real_cat = torch.cat([m, x], 1)
    fake_cat = torch.cat([m, fake], 1)
    out = do.multiscale_forward(dsd, torch.cat([real_cat, fake_cat], 0), sn=d_sn)
    real_d, fake_d = do.divide_pred(out)
    return do.gan_loss(fake_d, False) + do.gan_loss(real_d, True)


class ExtraAdam:
    """optim.py:137-291 on a dict of tensors (those with requires_grad)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = params
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = {id(p): dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p)) for p in params}
        self.copy = []

    def _update(self, p):
        if p.grad is None:
            return None
        s = self.state[id(p)]
        b1, b2 = self.betas
        s["step"] += 1
        s["m"].mul_(b1).add_(p.grad, alpha=1 - b1)
        s["v"].mul_(b2).addcmul_(p.grad, p.grad, value=1 - b2)
        denom = s["v"].sqrt().add_(self.eps)
        step_size = self.lr * math.sqrt(1 - b2 ** s["step"]) / (1 - b1 ** s["step"])
        return -step_size * s["m"] / denom

    @torch.no_grad()
    def extrapolation(self):
        empty = len(self.copy) == 0
        for p in self.params:
            u = self._update(p)
            if empty:
                self.copy.append(p.data.clone())
            if u is not None:
                p.data.add_(u)

    @torch.no_grad()
    def step(self):
        assert self.copy, "Need to call extrapolation before calling step."
        for i, p in enumerate(self.params):
            u = self._update(p)
            if u is not None:
                p.data = self.copy[i].add_(u)
        self.copy = []
