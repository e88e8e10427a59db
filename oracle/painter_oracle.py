"""Oracle for the SPADE painter path: functional fp32 restatement of

  climategan/norms.py      SpectralNorm._update_u_v :100-112, SPADE.forward :174-186
  climategan/blocks.py     SPADEResnetBlock.forward/shortcut/activation :369-395, InterpolateNearest2d :28-43
  climategan/painter.py    PainterSpadeDecoder.forward :149-168
  climategan/generator.py  OmniGenerator.paint :279-297

All functions take a reference-layout ``state_dict`` (the same keys the reference modules
produce) plus NCHW fp32 tensors, so one set of weights drives the reference, this oracle and the
CUDA path.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def l2normalize(v: Tensor, eps: float = 1e-12) -> Tensor:
    """norms.py:80-81."""
    return v / (v.norm() + eps)


def spectral_norm_weight(w_bar: Tensor, u: Tensor, v: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """norms.py:100-112 with power_iterations=1.  Returns (w, u_new, v_new); u/v are treated as
    constants for autograd (the reference updates ``.data``), sigma depends on w_bar."""
    height = w_bar.shape[0]
    w2 = w_bar.view(height, -1)
    with torch.no_grad():
        v_new = l2normalize(torch.mv(w2.t(), u))
        u_new = l2normalize(torch.mv(w2, v_new))
    sigma = u_new.dot(w2.mv(v_new))
    return w_bar / sigma.expand_as(w_bar), u_new, v_new


class SNState:
    """Tracks the mutable u/v vectors across forwards (they change on every call, norms.py:106-108)."""

    def __init__(self, sd: Dict[str, Tensor]):
        self.sd = sd

    def weight(self, prefix: str) -> Tensor:
        """One power iteration + w_bar / sigma.  u and v are PERSISTENT tensors whose ``.data`` is swapped in place, exactly as
        ``SpectralNorm._update_u_v`` does (norms.py:106-108): autograd saved them by reference for ``sigma = u.(W v)``, so
        when a layer runs twice before one backward (the mask decoder and the AdvEnt discriminators see the r and the s
        batch), BOTH backward passes use the u, v of the LAST forward in d(sigma)/dW = u v^T.  Reference behaviour, kept."""
        w_bar, u, v = self.sd[prefix + ".weight_bar"], self.sd[prefix + ".weight_u"], self.sd[prefix + ".weight_v"]
        w2 = w_bar.view(w_bar.shape[0], -1)
        with torch.no_grad():
            v_new = l2normalize(torch.mv(w2.t(), u))
            u_new = l2normalize(torch.mv(w2, v_new))
        v.data = v_new
        u.data = u_new
        sigma = u.dot(w2.mv(v))
        return w_bar / sigma.expand_as(w_bar)


def spade(sd: Dict[str, Tensor], prefix: str, x: Tensor, segmap: Tensor) -> Tensor:
    """norms.py:174-186 (instance flavour: nn.InstanceNorm2d(affine=False), eps 1e-5 :151)."""
    normalized = F.instance_norm(x, eps=1e-5)
    segmap = F.interpolate(segmap, size=x.size()[2:], mode="nearest")
    actv = F.relu(F.conv2d(segmap, sd[prefix + ".mlp_shared.0.weight"], sd[prefix + ".mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[prefix + ".mlp_gamma.weight"], sd[prefix + ".mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[prefix + ".mlp_beta.weight"], sd[prefix + ".mlp_beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


def spade_resnet_block(sd, sn: SNState, prefix: str, x: Tensor, seg: Tensor, spectral: bool = True) -> Tensor:
    """blocks.py:369-395."""

    def conv(name, inp, padding):
        p = f"{prefix}.{name}"
        if spectral:
            w = sn.weight(p + ".module")
            b = sd.get(p + ".module.bias")
        else:
            w, b = sd[p + ".weight"], sd.get(p + ".bias")
        return F.conv2d(inp, w, b, padding=padding)

    learned_shortcut = (f"{prefix}.conv_s.module.weight_bar" in sd) or (f"{prefix}.conv_s.weight" in sd)
    if learned_shortcut:
        x_s = conv("conv_s", spade(sd, prefix + ".norm_s", x, seg), 0)
    else:
        x_s = x
    dx = conv("conv_0", F.leaky_relu(spade(sd, prefix + ".norm_0", x, seg), 2e-1), 1)
    dx = conv("conv_1", F.leaky_relu(spade(sd, prefix + ".norm_1", dx, seg), 2e-1), 1)
    return x_s + dx


def upsample2(x: Tensor) -> Tensor:
    """blocks.py:39-43."""
    return F.interpolate(x, size=(x.shape[-2] * 2, x.shape[-1] * 2), mode="nearest")


def painter_forward(sd: Dict[str, Tensor], cond: Tensor, z_h: int, z_w: int, n_up_spades: int,
                    sn: SNState = None) -> Tensor:
    """painter.py:149-168 with z=None (no_z) and no final shortcut."""
    sn = sn or SNState(sd)
    z = F.conv2d(F.interpolate(cond, size=(z_h, z_w)), sd["fc.weight"], sd["fc.bias"], padding=1)
    y = spade_resnet_block(sd, sn, "head_0", z, cond)
    y = upsample2(y)
    y = spade_resnet_block(sd, sn, "G_middle_0", y, cond)
    y = upsample2(y)
    y = spade_resnet_block(sd, sn, "G_middle_1", y, cond)
    for i in range(n_up_spades):
        y = upsample2(y)
        y = spade_resnet_block(sd, sn, f"up_spades.{i}", y, cond)
    y = spade_resnet_block(sd, sn, "final_spade", y, cond)
    y = F.conv2d(F.leaky_relu(y, 2e-1), sd["conv_img.weight"], sd["conv_img.bias"], padding=1)
    return torch.tanh(y)


def paint(sd, m: Tensor, x: Tensor, z_h: int, z_w: int, n_up_spades: int, paste: bool = True,
          sn: SNState = None) -> Tensor:
    """generator.py:279-297 (painter keys un-prefixed)."""
    m = m.to(x.dtype)
    fake = painter_forward(sd, x * (1.0 - m), z_h, z_w, n_up_spades, sn)
    if paste:
        return x * (1.0 - m) + fake * m
    return fake


def n_up_spades_of(sd) -> int:
    return len({k.split(".")[1] for k in sd if k.startswith("up_spades.")})
