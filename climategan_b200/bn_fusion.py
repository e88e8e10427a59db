"""BatchNorm fusion — the surface of ``climategan/bn_fusion.py`` (``bn_fuse`` :97-118, used by ``apply_events.py --fuse``).

The reference rewrites the module tree (Conv2d+BatchNorm2d pairs that are adjacent in leaf order become one Conv2d, the
BatchNorm an Identity).  Here eval-mode forwards ALREADY run every conv with its BatchNorm folded into the packed weights and
bias (``deeplab.resnetmulti_v2.fold_bn``, ``blocks.Conv2dBlock.forward_infer``: one kernel launch per conv+BN+activation), and
the folded packings are cached between forwards.  ``bn_fuse`` therefore keeps the module tree (and the state_dict) intact:
it puts the model in eval mode and returns it, so ``--fuse`` code paths work unchanged.  (The reference's fusion overwrites an
existing conv bias — SURVEY.md §7 "bug-compatibility decisions" — which the in-kernel folding does not reproduce: it computes
``beta + (b - mean) * gamma / sqrt(var + eps)``.)"""
from __future__ import annotations

import torch.nn as nn


def bn_fuse(model: nn.Module) -> nn.Module:
    n_bn = sum(1 for m in model.modules() if isinstance(m, nn.BatchNorm2d))
    model.eval()
    model._cgb_bn_fused = n_bn
    return model


def get_bn_fused_count(model: nn.Module) -> int:
    """Number of BatchNorm2d layers that eval-mode forwards fold into their convolutions."""
    return int(getattr(model, "_cgb_bn_fused", 0))
