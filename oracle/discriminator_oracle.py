"""Oracle for the painter's discriminator and GAN losses: functional fp32 restatement of

  climategan/discriminator.py  NLayerDiscriminator.forward :172-182 (groups built :100-166),
                               MultiscaleDiscriminator.forward :227-239 (AvgPool2d 3/s2/p1, count_include_pad=False :223-225)
  climategan/tutils.py         divide_pred :443-470
  climategan/losses.py         GANLoss :13-83 (no soft shift / flip), FeatMatchLoss :86-103, HingeLoss :550-593

TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle.painter_oracle import SNState


def nlayer_forward(sd, sn: SNState, prefix: str, x):
    """All groups of one PatchGAN: model0 = SN conv s2 + lrelu; model1..n-1 = SN conv s2 + IN + lrelu;
    model n = SN conv s1 + IN + lrelu; last = SN conv s1."""
    groups = sorted({int(k[len(prefix) + 6:].split(".")[0]) for k in sd if k.startswith(prefix + ".model")})
    last = groups[-1]
    feats = []
    for g in groups:
        p = f"{prefix}.model{g}.0.module"
        w = sn.weight(p)
        b = sd.get(p + ".bias")
        stride = 2 if g < last - 1 else 1
        x = F.conv2d(x, w, b, stride=stride, padding=1)
        if 0 < g < last:
            x = F.instance_norm(x, eps=1e-5)
        if g < last:
            x = F.leaky_relu(x, 0.2)
        feats.append(x)
    return feats


def multiscale_forward(sd, x, prefix="", sn: SNState = None):
    sn = sn or SNState(sd)
    pre = prefix + "." if prefix else ""
    names = sorted({k[len(pre):].split(".")[0] for k in sd if k.startswith(pre + "discriminator_")})
    result = []
    for name in names:
        result.append(nlayer_forward(sd, sn, pre + name, x))
        x = F.avg_pool2d(x, 3, stride=2, padding=[1, 1], count_include_pad=False)
    return result


def divide_pred(disc_output):
    half1 = [[t[: t.size(0) // 2] for t in p] for p in disc_output]
    half2 = [[t[t.size(0) // 2:] for t in p] for p in disc_output]
    return half1, half2


def gan_loss(preds, target_is_real, use_lsgan=False, real_label=1.0, fake_label=0.0):
    total = 0
    for p in preds:
        p = p[-1] if isinstance(p, list) else p
        t = torch.full_like(p, real_label if target_is_real else fake_label)
        total = total + (F.mse_loss(p, t) if use_lsgan else F.binary_cross_entropy_with_logits(p, t))
    return total / len(preds)


def hinge_loss(preds, target_is_real, for_discriminator=True):
    total = 0
    for p in preds:
        p = p[-1] if isinstance(p, list) else p
        if for_discriminator:
            total = total + (-torch.mean(torch.min(p - 1, torch.zeros_like(p))) if target_is_real
                             else -torch.mean(torch.min(-p - 1, torch.zeros_like(p))))
        else:
            total = total - torch.mean(p)
    return total / len(preds)


def feat_match_loss(pred_real, pred_fake):
    num_d = len(pred_fake)
    total = 0.0
    for i in range(num_d):
        for j in range(len(pred_fake[i]) - 1):
            total = total + F.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) / num_d
    return total
