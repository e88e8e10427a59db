"""BatchNorm fusion — ``climategan/bn_fusion.py`` (``bn_fuse`` :97-118, used by ``apply_events.py --fuse``, :466).

Two things live here:

* ``bn_fuse(model)`` — the reference's MODULE REWRITE: walk the leaves of the module tree in registration order
  (``FlattableModel._flatten_model`` :25-35), and wherever an ``nn.Conv2d`` is immediately followed by an ``nn.BatchNorm2d``
  with matching channel count fold the running statistics into the conv (``weight *= alpha``, ``bias = beta``,
  ``_calculate_alpha_beta`` :121-133) and replace the BatchNorm by an identity layer.  The result is an inference-only model
  whose convs run WITHOUT any BatchNorm work; the forwards of this package recognise the identity layer (``is_fused``).

  ``compat=True`` (default) reproduces the reference bit for bit, including its quirk: ``item.bias = Parameter(beta)`` (:115)
  OVERWRITES an existing conv bias instead of folding it (``beta + bias * alpha``) — harmless for the reference's own models,
  whose BatchNorm'd convs have ``bias=False`` except the deeplabv3 ASPP convs.  ``compat=False`` folds
  the bias correctly.  Unlike the reference the model is NOT moved to the CPU (:98) — it stays where it is — and, like the
  reference, a deep copy is returned and the original left untouched.

* Without ``bn_fuse`` nothing is lost: every eval-mode forward of this package already folds BatchNorm into the packed weights
  on the fly (``deeplab.resnetmulti_v2.fold_bn``, cached until a parameter or a running statistic changes).
"""
from __future__ import annotations

from copy import deepcopy

import torch
import torch.nn as nn


class _IdentityLayer(nn.Module):
    """What a fused BatchNorm2d becomes (bn_fusion.py:136-138)."""

    def forward(self, input):
        return input


def is_fused(bn) -> bool:
    return isinstance(bn, _IdentityLayer)


def _leaves(module, prefix=()):
    """(path, leaf module) in the reference's flattening order: depth-first over named_children; childless modules that are
    not Sequential / ModuleList containers are the leaves (bn_fusion.py:17-35)."""
    out = []
    children = list(module.named_children())
    order = getattr(module, "_reference_leaf_order", None)
    if order is not None:
        # a module of this package that registers fewer / differently ordered children than its reference counterpart states
        # the reference's leaf sequence itself; None stands for a reference leaf that has no counterpart here (a padding or
        # activation module) and only serves to break adjacency
        children = order()
    for name, c in children:
        if c is None:
            out.append((prefix + (name,), None))
        else:
            out += _leaves(c, prefix + (name,))
    if not children and not isinstance(module, (nn.Sequential, nn.ModuleList)) and not hasattr(module, "_restricted"):
        out = [(prefix, module)]
    return out


def _calculate_alpha_beta(bn):
    """bn_fusion.py:121-133."""
    std = torch.sqrt(bn.running_var + bn.eps)
    alpha = bn.weight.data / std
    beta = -(bn.weight.data * bn.running_mean) / std + bn.bias.data
    return alpha, beta


def bn_fuse(model: nn.Module, compat: bool = True) -> nn.Module:
    model = deepcopy(model)
    leaves = _leaves(model)
    fused = 0
    for i, (path, item) in enumerate(leaves[:-1]):
        nxt_path, nxt = leaves[i + 1]
        if not (isinstance(item, nn.Conv2d) and isinstance(nxt, nn.BatchNorm2d)):
            continue
        if nxt.weight is None or nxt.running_mean is None:
            continue   # a parameter-free / statistics-free BatchNorm (SPADE's): the reference's fusion raises on it; skipped here
        alpha, beta = _calculate_alpha_beta(nxt)
        if item.weight.shape[0] != alpha.shape[0]:
            continue   # something else sat between the two in the forward (bn_fusion.py:110-113)
        with torch.no_grad():
            if not compat and item.bias is not None:
                beta = beta + item.bias.data * alpha
            item.weight.data = item.weight.data * alpha.view(-1, 1, 1, 1)
            item.bias = nn.Parameter(beta.clone())
        parent = model
        for name in nxt_path[:-1]:
            parent = getattr(parent, name)
        setattr(parent, nxt_path[-1], _IdentityLayer())
        fused += 1
    model.eval()
    model._cgb_bn_fused = fused
    from . import ops

    ops.invalidate_weight_cache()
    return model


def get_bn_fused_count(model: nn.Module) -> int:
    """Number of Conv2d + BatchNorm2d pairs ``bn_fuse`` folded."""
    return int(getattr(model, "_cgb_bn_fused", 0))
