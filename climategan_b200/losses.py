"""GAN-side losses of the painter step — ``climategan/losses.py`` GANLoss (:13-83), FeatMatchLoss (:86-103), HingeLoss
(:550-593) — with the reference call signatures, evaluated by libcgb200 reduce kernels (value + gradient in one pass).
"""
from __future__ import annotations

from random import random as rand

import torch
import torch.nn as nn

from . import ops


class GANLoss(nn.Module):
    """BCE-with-logits (use_lsgan=False) or MSE (use_lsgan=True) against a constant real/fake label, with the
    reference's label smoothing (``soft_shift``) and label flipping (``flip_prob``) draws in the same order."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, soft_shift=0.0, flip_prob=0.0,
                 verbose=0):
        super().__init__()
        self.soft_shift = soft_shift
        self.verbose = verbose
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.kind = ops.LOSS_MSE if use_lsgan else ops.LOSS_BCE_LOGITS
        self.flip_prob = flip_prob

    def get_target_value(self, target_is_real):
        soft_change = float(torch.FloatTensor(1).uniform_(0, self.soft_shift))  # same RNG draw as losses.py:57
        return float(self.real_label) - soft_change if target_is_real else float(self.fake_label) + soft_change

    def __call__(self, input, target_is_real, *args, **kwargs):
        r = rand()
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                if r < self.flip_prob:
                    target_is_real = not target_is_real
                loss = loss + ops.const_target_loss(pred_i, self.kind, self.get_target_value(target_is_real))
            return loss / len(input)
        if r < self.flip_prob:
            target_is_real = not target_is_real
        return ops.const_target_loss(input, self.kind, self.get_target_value(target_is_real))


class HingeLoss(nn.Module):
    def __init__(self, tensor=None):
        super().__init__()

    def loss(self, input, target_is_real, for_discriminator=True):
        if for_discriminator:
            return ops.const_target_loss(input, ops.LOSS_HINGE_D_REAL if target_is_real else ops.LOSS_HINGE_D_FAKE)
        assert target_is_real, "The generator's hinge loss must be aiming for real"
        return ops.const_target_loss(input, ops.LOSS_NEG_MEAN)

    def __call__(self, input, target_is_real, for_discriminator=True):
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                loss = loss + self.loss(pred_i, target_is_real, for_discriminator)
            return loss / len(input)
        return self.loss(input, target_is_real, for_discriminator)


class FeatMatchLoss(nn.Module):
    """L1 between the discriminator's intermediate features of the fake and (detached) real halves, summed over
    layers, divided by num_D.  Accepts NCHW fp32 feature lists (reference contract)."""

    def __call__(self, pred_real, pred_fake):
        num_D = len(pred_fake)
        total = 0.0
        for i in range(num_D):
            for j in range(len(pred_fake[i]) - 1):
                total = total + ops.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) / num_D
        return total
