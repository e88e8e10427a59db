"""Oracle for the FULL train step (tasks d, s, m, p): functional fp32 restatement of

  climategan/trainer.py   get_G_loss :1162-1182, get_masker_loss :1184-1254, masker_d_loss :1389-1407,
                          masker_s_loss :1409-1516, masker_m_loss :1518-1616, get_D_loss :1034-1160,
                          update_G/update_D :989-1032, run_epoch's D freeze :958-973
  climategan/losses.py    CrossEntropy :106-112, TVLoss :140-171, MinentLoss :177-196, SIGMLoss :232-278,
                          GroundIntersectionLoss :449-455, prob_2_entropy :466-471, CustomBCELoss :472-477,
                          ADVENTAdversarialLoss :480-524 (WGAN form for G; losses["D"]["advent"] is always BCE, :440)
  climategan/discriminator.py  get_fc_discriminator :327-361

on reference-layout state_dicts, built on the masker / painter / discriminator / trainer oracles.  Dropout is the identity
here (parity runs set p = 0 on both sides).  Pinned by tests/golden/full_step.* (the reference's own Trainer).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import discriminator_oracle as do
from oracle import masker_oracle as mo
from oracle import trainer_oracle as to
from oracle.painter_oracle import SNState

LAMBDAS = dict(d_main=1.0, d_gml=0.5, s_crossent=1.0, s_minent=0.001, s_advent=0.001, m_bce=1.0, m_tv=1.0, m_gi=0.05,
               ent_main=0.5, ent_var=0.1, adv_main=1.0)
DOMAIN_LABELS = {"s": 0, "r": 1}


def sigm_loss(prediction, target, gmweight=0.5, scale=4):
    t_pred, t_targ = torch.median(prediction), torch.median(target)
    s_pred, s_targ = torch.mean(torch.abs(prediction - t_pred)), torch.mean(torch.abs(target - t_targ))
    R = (prediction - t_pred) / s_pred - (target - t_targ) / s_targ
    bs, num_pix = prediction.shape[0], prediction.shape[-1] * prediction.shape[-2]
    sx = torch.tensor([[1.0, 0, -1], [2, 0, -2], [1, 0, -1]]).expand(bs, 1, 3, 3)
    sy = torch.tensor([[1.0, 2, 1], [0, 0, 0], [-1, -2, -1]]).expand(bs, 1, 3, 3)
    gm = 0
    for k in range(scale):
        R_ = F.interpolate(R, scale_factor=1 / 2 ** k)
        gm = gm + torch.sum(torch.abs(F.conv2d(R_, sx)) + torch.abs(F.conv2d(R_, sy)))
    return 0.5 / num_pix * torch.sum(torch.abs(R)) + gmweight / num_pix * gm


def prob_2_entropy(prob):
    c = prob.shape[1]
    return -torch.mul(prob, torch.log2(prob + 1e-30)) / math.log2(c)


def minent(pred, version=1, lambda_var=0.1):
    n, c, h, w = pred.shape
    e = prob_2_entropy(pred)
    if version == 1:
        return torch.sum(e) / (n * h * w)
    dm = e - torch.sum(e) / (n * h * w)
    return torch.sum(e + lambda_var * dm * dm) / (n * h * w)


def tv_loss(x):
    b, c, h, w = x.shape
    h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, : h - 1, :], 2).sum()
    w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, : w - 1], 2).sum()
    return 2 * (h_tv / (c * (h - 1) * w) + w_tv / (c * h * (w - 1))) / b


def fc_discriminator(dsd, sn, prefix, x):
    for i in range(0, 9, 2):
        p = f"{prefix}.{i}"
        if p + ".module.weight_bar" in dsd:
            w, b = sn.weight(p + ".module"), dsd[p + ".module.bias"]
        else:
            w, b = dsd[p + ".weight"], dsd[p + ".bias"]
        x = F.conv2d(x, w, b, stride=2, padding=1)
        if i < 8:
            x = F.leaky_relu(x, 0.2)
    return x


def advent(prob, target, dsd, d_sn, prefix, depth=None, wgan=True):
    d_in = prob_2_entropy(prob)
    if depth is not None:
        d_in = d_in * depth
    d_out = fc_discriminator(dsd, d_sn, prefix, d_in)
    if wgan:
        return -torch.mean(target * d_out + (1 - target) * (1 - d_out))
    return F.binary_cross_entropy_with_logits(d_out, torch.full_like(d_out, float(target)))


def masker_forward(gsd, g_sn, x, d_target, s_target):
    z = mo.encoder(gsd, x)
    d, z_depth = mo.depth_decoder(gsd, g_sn, z, d_target)
    s = mo.seg_decoder(gsd, z, z_depth, s_target)
    logits = mo.mask_decoder(gsd, g_sn, z)
    return z, d, z_depth, s, logits


def masker_g_loss(gsd, dsd, batch, g_sn, d_sn, lam=LAMBDAS):
    """get_masker_loss: per domain (r, s) d -> s -> m, in the reference's order of spectral-norm power iterations."""
    total = 0
    terms = {}
    with mo.train_mode():
        for domain in ("r", "s"):
            if domain not in batch:
                continue
            data = batch[domain]["data"]
            x = data["x"]
            q = data["d"].shape[-1]
            z = mo.encoder(gsd, x)
            d, z_depth = mo.depth_decoder(gsd, g_sn, z, q)
            if domain == "s":
                l_d = sigm_loss(d, data["d"], lam["d_gml"]) * lam["d_main"]
                terms["d.s"] = l_d
                total = total + l_d
            s = mo.seg_decoder(gsd, z, z_depth, data["s"].shape[-1])
            if domain == "s":
                l = F.cross_entropy(s, data["s"].squeeze(1).long()) * lam["s_crossent"]
                terms["s.crossent.s"] = l
                total = total + l
            else:
                sm = torch.softmax(s, 1)
                l1 = minent(sm) * lam["s_minent"]
                l2 = advent(sm, DOMAIN_LABELS["s"], dsd, d_sn, "s.Advent", d.detach(), wgan=True) * lam["s_advent"]
                terms["s.minent.r"], terms["s.advent.r"] = l1, l2
                total = total + l1 + l2
            logits = mo.mask_decoder(gsd, g_sn, z)
            p = torch.sigmoid(logits)
            prob = torch.cat([p, 1 - p], 1)
            l = tv_loss(p) * lam["m_tv"]
            terms[f"m.tv.{domain}"] = l
            total = total + l
            if domain == "s":
                l = F.binary_cross_entropy_with_logits(logits, data["m"]) * lam["m_bce"]
                terms["m.bce.s"] = l
                total = total + l
            else:
                l_gi = torch.mean(1.0 * ((data["m"] - p) > 0.5)) * lam["m_gi"]
                l_me = minent(prob, 2, lam["ent_var"]) * lam["ent_main"]
                l_adv = advent(prob, DOMAIN_LABELS["s"], dsd, d_sn, "m.Advent", None, wgan=True) * lam["adv_main"]
                terms["m.gi.r"], terms["m.minent.r"], terms["m.advent.r"] = l_gi, l_me, l_adv
                total = total + l_gi + l_me + l_adv
    return total, terms


def masker_d_loss(gsd, dsd, batch, g_sn, d_sn, lam=LAMBDAS):
    """get_D_loss, masker branch: D["s"]/D["m"] AdvEnt on detached predictions; losses["D"]["advent"] is BCE (losses.py:440);
    adv_main is applied twice (inside masker_*_loss and again at trainer.py:1123,1146)."""
    out = {"s": 0, "m": 0}
    with mo.train_mode():
        for domain in ("r", "s"):
            if domain not in batch:
                continue
            data = batch[domain]["data"]
            with torch.no_grad():
                z = mo.encoder(gsd, data["x"])
                d, z_depth = mo.depth_decoder(gsd, g_sn, z, data["d"].shape[-1])
                s = mo.seg_decoder(gsd, z, z_depth, data["s"].shape[-1])
            sm = torch.softmax(s, 1)
            out["s"] = out["s"] + advent(sm, DOMAIN_LABELS[domain], dsd, d_sn, "s.Advent", d, wgan=False) * lam["adv_main"] ** 2
            with torch.no_grad():
                p = torch.sigmoid(mo.mask_decoder(gsd, g_sn, z))
            prob = torch.cat([p, 1 - p], 1)
            out["m"] = out["m"] + advent(prob, DOMAIN_LABELS[domain], dsd, d_sn, "m.Advent", None, wgan=False) * lam["adv_main"] ** 2
    return out


def full_g_loss(gsd, dsd, vsd, batch, z_hw, g_sn=None, d_sn=None):
    """get_G_loss: masker loss then painter loss (trainer.py:1168-1176)."""
    g_sn, d_sn = g_sn or SNState(gsd), d_sn or SNState(dsd)
    m_loss, terms = masker_g_loss(gsd, dsd, batch, g_sn, d_sn)
    psd = _Prefixed(gsd, "painter.")
    dpsd = _Prefixed(dsd, "p.")
    rf = batch["rf"]["data"]
    p_loss, p_terms = to.painter_g_loss(psd, dpsd, vsd, rf["x"], rf["m"], z_hw, g_sn=SNState(psd), d_sn=SNState(dpsd))
    terms.update({"p." + k: v for k, v in p_terms.items()})
    return m_loss + p_loss, terms


def full_d_loss(gsd, dsd, batch, z_hw, g_sn=None, d_sn=None):
    """get_D_loss in the order multi_domain_batch iterates (r, s, rf)."""
    g_sn, d_sn = g_sn or SNState(gsd), d_sn or SNState(dsd)
    md = masker_d_loss(gsd, dsd, batch, g_sn, d_sn)
    psd = _Prefixed(gsd, "painter.")
    dpsd = _Prefixed(dsd, "p.")
    rf = batch["rf"]["data"]
    pd = to.painter_d_loss(psd, dpsd, rf["x"], rf["m"], z_hw, g_sn=SNState(psd), d_sn=SNState(dpsd))
    return md["m"] + md["s"] + pd, {"m.Advent": md["m"], "s.Advent": md["s"], "p.gan": pd}


class _Prefixed(dict):
    """View of a state_dict under a key prefix that writes through (spectral-norm u/v updates land in the parent)."""

    def __init__(self, parent, prefix):
        super().__init__({k[len(prefix):]: v for k, v in parent.items() if k.startswith(prefix)})
        self._parent, self._prefix = parent, prefix

    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        self._parent[self._prefix + k] = v
