"""Normalisation layers of the hot path — same class names, constructor arguments, parameter
names and state_dict layout as ``climategan/norms.py`` in the reference, computing through
libcgb200 on NHWC storage tensors (see :mod:`climategan_b200.ops`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


def _pad_c(t, cs):
    """[.., C] -> [.., Cs] zero padded (pad channels of a storage tensor must stay zero: scale = shift = 0 there)."""
    return t if t.shape[-1] == cs else torch.nn.functional.pad(t, (0, cs - t.shape[-1]))


def instance_norm_act(x, c, eps=1e-5, weight=None, bias=None, act=_lib.ACT_NONE, slope=0.2):
    """act(instance_norm(x) [* weight[n, c] + bias[n, c]]) on a storage tensor with c logical channels: biased variance, eps
    inside the square root (nn.InstanceNorm2d / F.batch_norm in training mode).  The statistics -> scale / shift algebra runs
    on [N, C] tensors under autograd; the tensor-sized work is one moments pass and one affine pass (ops.moments / affine_nc)."""
    cs = x.shape[-1]
    m1, m2 = ops.moments(x)
    var = (m2 - m1 * m1).clamp_min(0.0)
    scale = torch.rsqrt(var + eps)
    shift = -m1 * scale
    if weight is not None:
        w, b = _pad_c(weight.view(x.shape[0], -1), cs), _pad_c(bias.view(x.shape[0], -1), cs)
        scale, shift = scale * w, shift * w + b
    live = (torch.arange(cs, device=x.device) < c).to(scale.dtype)
    return ops.affine_nc(x, scale * live, shift * live, act, slope)


class AdaptiveInstanceNorm2d(nn.Module):
    """``climategan.norms.AdaptiveInstanceNorm2d`` (norms.py:8-46): instance norm whose per-(sample, channel) ``weight`` /
    ``bias`` ([B*C] tensors) are assigned from outside before the call; same dummy ``running_mean`` / ``running_var`` buffers
    (state_dict surface)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features = num_features
        self.eps = eps
        self.momentum = momentum
        self.weight = None
        self.bias = None
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x, act=_lib.ACT_NONE, slope=0.2):
        assert self.weight is not None and self.bias is not None, "Please assign weight and bias before calling AdaIN!"
        return instance_norm_act(x, self.num_features, self.eps, self.weight, self.bias, act, slope)

    def __repr__(self):
        return self.__class__.__name__ + "(" + str(self.num_features) + ")"


class LayerNorm(nn.Module):
    """``climategan.norms.LayerNorm`` (norms.py:49-81; MUNIT's): per-SAMPLE mean and UNBIASED standard deviation over (C, H, W),
    ``(x - mean) / (std + eps)`` — eps outside the square root — then a per-channel ``gamma`` / ``beta``."""

    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__()
        self.num_features = num_features
        self.affine = affine
        self.eps = eps
        if self.affine:
            self.gamma = nn.Parameter(torch.Tensor(num_features).uniform_())
            self.beta = nn.Parameter(torch.zeros(num_features))

    def forward(self, x, act=_lib.ACT_NONE, slope=0.2):
        n, h, w, cs = x.shape
        c = self.num_features
        m1, m2 = ops.moments(x)
        mu = m1[:, :c].mean(1, keepdim=True)                     # every channel has the same pixel count
        e2 = m2[:, :c].mean(1, keepdim=True)
        cnt = float(c * h * w)
        var = ((e2 - mu * mu) * (cnt / max(cnt - 1.0, 1.0))).clamp_min(0.0)   # torch.std: unbiased
        inv = 1.0 / (torch.sqrt(var) + self.eps)
        scale = inv.expand(n, c)
        shift = (-mu * inv).expand(n, c)
        if self.affine:
            scale, shift = scale * self.gamma.view(1, c), shift * self.gamma.view(1, c) + self.beta.view(1, c)
        return ops.affine_nc(x, _pad_c(scale, cs), _pad_c(shift, cs), act, slope)


class SpectralNorm(nn.Module):
    """Drop-in for ``climategan.norms.SpectralNorm`` (norms.py:84-143).

    Parameters live on the wrapped module exactly as in the reference: ``module.weight_bar``
    (trainable), ``module.weight_u`` / ``module.weight_v`` (``requires_grad=False``), in that
    registration order after ``bias`` (norms.py:123-139), so checkpoints interchange.
    The power iteration runs on every :meth:`effective_weight` call — train *and* eval — and
    mutates ``u``/``v`` in place, as ``_update_u_v`` does (norms.py:100-112).
    """

    def __init__(self, module, name="weight", power_iterations=1):
        super().__init__()
        if power_iterations != 1:
            raise NotImplementedError("the reference only ever uses power_iterations=1")
        self.module = module
        self.name = name
        self.power_iterations = power_iterations
        if not self._made_params():
            self._make_params()

    def _made_params(self):
        return all(hasattr(self.module, self.name + s) for s in ("_u", "_v", "_bar"))

    def _make_params(self):
        w = getattr(self.module, self.name)
        height = w.data.shape[0]
        width = w.view(height, -1).data.shape[1]
        # same RNG draws, in the same order, as the reference
        u = nn.Parameter(w.data.new(height).normal_(0, 1), requires_grad=False)
        v = nn.Parameter(w.data.new(width).normal_(0, 1), requires_grad=False)
        u.data = l2normalize(u.data)
        v.data = l2normalize(v.data)
        w_bar = nn.Parameter(w.data)
        del self.module._parameters[self.name]
        self.module.register_parameter(self.name + "_u", u)
        self.module.register_parameter(self.name + "_v", v)
        self.module.register_parameter(self.name + "_bar", w_bar)

    def effective_weight(self) -> torch.Tensor:
        m = self.module
        # u / v are passed as the Parameters themselves: the kernel updates their storage in place (as ``u.data = ...`` does,
        # norms.py:106-108) and, when the trainer has flipped their requires_grad (discriminators), they receive gradients
        return ops.spectral_weight(getattr(m, self.name + "_bar"), getattr(m, self.name + "_u"), getattr(m, self.name + "_v"))

    @property
    def bias(self):
        return self.module.bias


def conv_weight_bias(conv):
    """(weight, bias) of an nn.Conv2d or a SpectralNorm-wrapped one (runs the power iteration)."""
    if isinstance(conv, SpectralNorm):
        return conv.effective_weight(), conv.module.bias
    return conv.weight, conv.bias


class SPADE(nn.Module):
    """Drop-in for ``climategan.norms.SPADE`` (norms.py:146-186), instance-norm flavour.

    ``forward(x, segmap, stats=None, act=ACT_NONE)`` takes/returns NHWC storage tensors; ``segmap``
    must already be at x's resolution (the caller resizes once per resolution and shares it, the
    reference re-interpolates per layer, norms.py:179 — same values).  ``stats`` lets norm_0 and
    norm_s of a block share one statistics pass over the same x.
    """

    def __init__(self, param_free_norm_type, kernel_size, norm_nc, cond_nc):
        super().__init__()
        if param_free_norm_type == "instance":
            self.param_free_norm = nn.InstanceNorm2d(norm_nc, affine=False)  # holds no state
        elif param_free_norm_type == "batch":
            self.param_free_norm = nn.BatchNorm2d(norm_nc, affine=False)  # running statistics: the masker's SPADE decoder
        else:
            raise ValueError("%s is not a recognized param-free norm type in SPADE" % param_free_norm_type)
        nhidden = 128
        pw = kernel_size // 2
        self.norm_nc = norm_nc
        self.mlp_shared = nn.Sequential(nn.Conv2d(cond_nc, nhidden, kernel_size=kernel_size, padding=pw), nn.ReLU())
        self.mlp_gamma = nn.Conv2d(nhidden, norm_nc, kernel_size=kernel_size, padding=pw)
        self.mlp_beta = nn.Conv2d(nhidden, norm_nc, kernel_size=kernel_size, padding=pw)

    def forward(self, x, segmap, stats=None, act=_lib.ACT_NONE, slope=0.2, seg_col=None):
        """seg_col: optional im2col patches of segmap (ops.im2col) shared by all SPADE layers of a resolution."""
        sh = self.mlp_shared[0]
        k = sh.kernel_size[0]
        if seg_col is None:
            if segmap.shape[1:3] != x.shape[1:3]:
                segmap = ops.resize_nearest(segmap, x.shape[1], x.shape[2])
            if k * k * sh.in_channels <= 64 and not (torch.is_grad_enabled() and segmap.requires_grad):
                seg_col = ops.im2col(segmap, sh.in_channels, k, k // 2)
        batch_stats = False
        if isinstance(self.param_free_norm, nn.BatchNorm2d):
            bn = self.param_free_norm
            if bn.training or bn.running_mean is None:
                # train mode (norms.py:154-155, 177): statistics over (N, H, W), running statistics updated; the backward
                # differentiates through the batch statistics
                mean, rstd = ops.batchnorm_stats_update(x, bn)
                batch_stats = True
            else:
                n, cs = x.shape[0], x.shape[-1]
                mean = torch.zeros(n, cs, dtype=torch.float32, device=x.device)
                rstd = torch.ones(n, cs, dtype=torch.float32, device=x.device)
                mean[:, : self.norm_nc] = bn.running_mean
                rstd[:, : self.norm_nc] = torch.rsqrt(bn.running_var + bn.eps)
        else:
            mean, rstd = stats if stats is not None else ops.instnorm_stats(x, self.param_free_norm.eps)
        seg_in, is_col = (seg_col, True) if seg_col is not None else (segmap, False)
        return ops.spade(x, mean, rstd, seg_in, sh.weight, sh.bias, self.mlp_gamma.weight, self.mlp_gamma.bias,
                         self.mlp_beta.weight, self.mlp_beta.bias, act, slope, seg_is_col=is_col, batch_stats=batch_stats)
