"""Conv / SPADE building blocks — drop-ins for ``climategan/blocks.py`` (same class names,
constructor signatures and state_dict keys), operating on NHWC storage tensors through libcgb200.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .norms import SPADE, AdaptiveInstanceNorm2d, LayerNorm, SpectralNorm, conv_weight_bias, instance_norm_act

_ACTS = {
    "relu": (_lib.ACT_RELU, 0.0),
    "lrelu": (_lib.ACT_LRELU, 0.2),
    "tanh": (_lib.ACT_TANH, 0.0),
    "sigmoid": (_lib.ACT_SIGMOID, 0.0),
    "none": (_lib.ACT_NONE, 0.0),
    "selu": (_lib.ACT_SELU, 0.0),
    "prelu": (_lib.ACT_NONE, 0.0),   # nn.PReLU(): a learnable slope, applied after the conv / norm (Conv2dBlock._prelu)
}


class InterpolateNearest2d(nn.Module):
    """``climategan.blocks.InterpolateNearest2d`` (blocks.py:11-43) on storage tensors."""

    def __init__(self, scale_factor=2):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x):
        return ops.resize_nearest(x, x.shape[1] * self.scale_factor, x.shape[2] * self.scale_factor)


class Conv2dBlock(nn.Module):
    """``climategan.blocks.Conv2dBlock`` (blocks.py:49-147): pad -> conv -> norm -> activation.

    Every option of the reference: pad_type zero / reflect / replicate; norm none / spectral / batch / instance / layer / adain
    (and the ``spectral_`` prefixed combinations); activation relu / lrelu / prelu / selu / tanh / sigmoid / none.  On the
    default path (no norm, or BatchNorm) bias + activation ride in the conv epilogue and the BatchNorm statistics in it too;
    the MUNIT-era normalisations (instance / layer / adain: no configuration of the reference selects them) run as a moments
    pass + one per-(sample, channel) affine pass (norms.instance_norm_act, LayerNorm, AdaptiveInstanceNorm2d).
    """

    def __init__(self, input_dim, output_dim, kernel_size, stride=1, padding=0, dilation=1, norm="none",
                 activation="relu", pad_type="zero", bias=True):
        super().__init__()
        self.use_bias = bias
        self.replicate = pad_type == "replicate"
        if pad_type == "reflect":
            self.pad_mode = _lib.PAD_REFLECT
        elif pad_type in ("zero", "replicate"):      # replicate: an explicit padded copy (ops.replicate_pad), then a pad-0 conv
            self.pad_mode = _lib.PAD_ZERO
        else:
            assert 0, "Unsupported padding type: {}".format(pad_type)
        self.padding, self.stride, self.dilation = padding, stride, dilation
        use_spectral_norm = False
        if norm.startswith("spectral_"):
            norm = norm.replace("spectral_", "")
            use_spectral_norm = True
        if norm in ("spectral", "none"):
            self.norm = None
        elif norm == "batch":
            self.norm = nn.BatchNorm2d(output_dim)  # eval: folded into the conv (forward_infer); train: batch statistics
        elif norm == "instance":
            self.norm = nn.InstanceNorm2d(output_dim)   # affine=False, no running statistics (blocks.py:80-82): holds no state
        elif norm == "layer":
            self.norm = LayerNorm(output_dim)
        elif norm == "adain":
            self.norm = AdaptiveInstanceNorm2d(output_dim)
        else:
            raise ValueError("Unsupported normalization: {}".format(norm))
        self.output_dim = output_dim
        self.kernel_size = kernel_size
        if activation not in _ACTS:
            raise ValueError("Unsupported activation: {}".format(activation))
        self.act, self.slope = _ACTS[activation]
        self.activation = None if activation == "none" else (nn.PReLU() if activation == "prelu" else activation)
        if norm == "spectral" or use_spectral_norm:   # blocks.py:117-127: the spectral branch keeps the bias even with BatchNorm
            self.conv = SpectralNorm(nn.Conv2d(input_dim, output_dim, kernel_size, stride, dilation=dilation, bias=self.use_bias))
        else:
            self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride, dilation=dilation,
                                  bias=self.use_bias if norm != "batch" else False)

    def _reference_leaf_order(self):
        """Leaf order of the reference Conv2dBlock for bn_fusion.bn_fuse: pad, norm, activation, conv (blocks.py:66-136 registers
        them in that order, so a block's own BatchNorm is never adjacent to its conv, nor to the previous block's)."""
        order = [("pad", None)]
        if self.norm is not None:
            order.append(("norm", self.norm))
        if self.activation is not None:
            order.append(("activation", None))
        order.append(("conv", self.conv))
        return order

    def forward(self, x, residual=None):
        """Training-capable forward (autograd tape): explicit reflect pad -> conv (tcgen05, pad 0) -> [train-mode
        BatchNorm + activation as one apply pass] ; bias + activation ride in the conv epilogue when there is no norm."""
        w, b = conv_weight_bias(self.conv)
        pad, pad_mode = self.padding, self.pad_mode
        if pad_mode == _lib.PAD_REFLECT and pad > 0:
            x = ops.reflect_pad(x, pad)
            pad, pad_mode = 0, _lib.PAD_ZERO
        elif self.replicate and pad > 0:
            x = ops.replicate_pad(x, pad)
            pad = 0
        prelu = isinstance(self.activation, nn.PReLU)
        if self.norm is not None and not isinstance(self.norm, nn.BatchNorm2d):
            # instance / layer / adain: conv (bias in the epilogue) -> moments -> one affine + activation pass
            y = ops.conv2d(x, w, b, None, stride=self.stride, dil=self.dilation, pad=pad, pad_mode=pad_mode)
            if isinstance(self.norm, nn.InstanceNorm2d):
                y = instance_norm_act(y, self.output_dim, self.norm.eps, act=self.act, slope=self.slope)
            else:
                y = self.norm(y, act=self.act, slope=self.slope)
            y = self._prelu(y) if prelu else y
            return y if residual is None else y + residual
        if prelu:
            y = ops.conv2d(x, w, b, None, stride=self.stride, dil=self.dilation, pad=pad, pad_mode=pad_mode)
            if self.norm is not None:
                y = ops.batchnorm_act(y, self.norm, None, _lib.ACT_NONE)
            y = self._prelu(y)
            return y if residual is None else y + residual
        if self.norm is not None:
            y = ops.conv_bn_act(x, w, self.norm, b, stride=self.stride, dil=self.dilation, pad=pad, pad_mode=pad_mode,
                                act=self.act, slope=self.slope)
            return y if residual is None else y + residual
        return ops.conv2d(x, w, b, residual, stride=self.stride, dil=self.dilation, pad=pad, pad_mode=pad_mode,
                          act=self.act, slope=self.slope)

    def _prelu(self, x):
        """nn.PReLU() (one learnable slope a): max(0, x) + a * min(0, x), as three affine passes so that autograd yields da."""
        n, cs = x.shape[0], x.shape[-1]
        one = torch.ones((n, cs), device=x.device)
        zero = torch.zeros((n, cs), device=x.device)
        pos = ops.affine_nc(x, one, zero, _lib.ACT_RELU)
        neg = ops.affine_nc(x, -one, zero, _lib.ACT_RELU)                  # relu(-x) = -min(0, x)
        return pos + ops.affine_nc(neg, -self.activation.weight.view(1, 1).expand(n, cs), zero)

    def forward_infer(self, x, residual=None):
        """Inference forward (no autograd tape): pad -> conv [-> eval BatchNorm folded] -> activation [+ residual]."""
        import torch

        if self.replicate or isinstance(self.activation, nn.PReLU) or (self.norm is not None and not isinstance(self.norm, nn.BatchNorm2d)):
            with torch.no_grad():     # the options off the default path share the training forward
                return self.forward(x, residual)
        with torch.no_grad():
            if self.norm is not None and not isinstance(self.conv, SpectralNorm):
                from .deeplab.resnetmulti_v2 import fold_bn

                wp, bp = fold_bn(self.conv, self.norm, x.dtype, cis=x.shape[-1])   # cached folded packing
            else:
                w, b = conv_weight_bias(self.conv)  # runs the spectral-norm power iteration, as the reference does in eval
                if self.norm is not None:
                    from .deeplab.resnetmulti_v2 import eval_bn_fold

                    w, b = eval_bn_fold(w, b, self.norm)
                wp = ops.pack_weight_cached(w, x.dtype, cis=x.shape[-1])
                bp = ops.pad_bias(b, wp.shape[0])
            pad, pad_mode = self.padding, self.pad_mode
            if pad_mode == _lib.PAD_REFLECT and pad > 0:
                # explicit reflect-padded copy + pad-0 conv: keeps the conv on the tcgen05 engine (the in-loader reflect
                # padding only exists in the SIMT engine: 512->512 3x3 @80x80 took 13.7 ms there)
                x = ops.reflect_pad(x, pad)
                pad, pad_mode = 0, _lib.PAD_ZERO
            return ops.conv2d_infer(x, wp, bp, residual, k=self.kernel_size, stride=self.stride,
                                    dil=self.dilation, pad=pad, pad_mode=pad_mode, act=self.act,
                                    slope=self.slope)


class SPADEResnetBlock(nn.Module):
    """``climategan.blocks.SPADEResnetBlock`` (blocks.py:325-398).

    forward (blocks.py:369-392):  x_s = conv_s(norm_s(x)) | x ;  dx = conv_0(lrelu(norm_0(x))) ;
    dx = conv_1(lrelu(norm_1(dx))) ; out = x_s + dx.  Here the leaky-relu rides in the SPADE
    modulation kernel, the residual add in conv_1's epilogue, and norm_0 / norm_s share one
    instance-norm statistics pass over x.
    """

    def __init__(self, fin, fout, cond_nc, spade_use_spectral_norm, spade_param_free_norm, spade_kernel_size,
                 last_activation=None):
        super().__init__()
        self.fin = fin
        self.fout = fout
        self.use_spectral_norm = spade_use_spectral_norm
        self.param_free_norm = spade_param_free_norm
        self.kernel_size = spade_kernel_size
        self.learned_shortcut = fin != fout
        self.last_activation = last_activation
        if last_activation not in (None, "lrelu"):
            raise NotImplementedError("The type of activation is not supported: {}".format(last_activation))
        fmiddle = min(fin, fout)
        # same construction order as the reference so a shared RNG seed gives identical initial weights
        self.conv_0 = nn.Conv2d(fin, fmiddle, kernel_size=3, padding=1)
        self.conv_1 = nn.Conv2d(fmiddle, fout, kernel_size=3, padding=1)
        if self.learned_shortcut:
            self.conv_s = nn.Conv2d(fin, fout, kernel_size=1, bias=False)
        if spade_use_spectral_norm:
            self.conv_0 = SpectralNorm(self.conv_0)
            self.conv_1 = SpectralNorm(self.conv_1)
            if self.learned_shortcut:
                self.conv_s = SpectralNorm(self.conv_s)
        self.norm_0 = SPADE(spade_param_free_norm, spade_kernel_size, fin, cond_nc)
        self.norm_1 = SPADE(spade_param_free_norm, spade_kernel_size, fmiddle, cond_nc)
        if self.learned_shortcut:
            self.norm_s = SPADE(spade_param_free_norm, spade_kernel_size, fin, cond_nc)

    def forward(self, x, seg, seg_col=None):
        """x: storage [N,H,W,round8(fin)] ; seg: storage conditioning at the same H,W ; seg_col: its im2col
        patches (ops.im2col), computed here when not supplied and shared by the block's 2-3 SPADE layers."""
        if seg_col is None and not (torch.is_grad_enabled() and seg.requires_grad):
            # (a conditioning tensor that needs a gradient takes the direct 3x3 mlp_shared path: im2col has no adjoint here)
            sh = self.norm_0.mlp_shared[0]
            if sh.kernel_size[0] ** 2 * sh.in_channels <= 64:
                seg_col = ops.im2col(seg, sh.in_channels, sh.kernel_size[0], sh.kernel_size[0] // 2)
        # norm_0 / norm_s see the same x: with instance norm they share one statistics pass; the batch flavour (masker) has
        # its own running statistics per SPADE layer
        stats = ops.instnorm_stats(x) if self.param_free_norm == "instance" else None
        if self.learned_shortcut:
            # reference order (blocks.py:370,389): shortcut first -> conv_s's power iteration runs first
            w_s, _ = conv_weight_bias(self.conv_s)
            x_s = ops.conv2d(self.norm_s(x, seg, stats, _lib.ACT_NONE, 0.2, seg_col), w_s, None)
        else:
            x_s = x
        w0, b0 = conv_weight_bias(self.conv_0)
        dx = ops.conv2d(self.norm_0(x, seg, stats, _lib.ACT_LRELU, 0.2, seg_col), w0, b0, pad=1)
        w1, b1 = conv_weight_bias(self.conv_1)
        out = ops.conv2d(self.norm_1(dx, seg, None, _lib.ACT_LRELU, 0.2, seg_col), w1, b1, x_s, pad=1)
        if self.last_activation == "lrelu":
            out = ops.activation(out, _lib.ACT_LRELU, 0.2)
        return out


class ResBlock(nn.Module):
    """``climategan.blocks.ResBlock`` (blocks.py:174-197): two 3x3 Conv2dBlocks + in-place residual (fused in the 2nd
    conv's epilogue)."""

    def __init__(self, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.dim, self.norm, self.activation = dim, norm, activation
        self.model = nn.Sequential(
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation=activation, pad_type=pad_type),
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation="none", pad_type=pad_type),
        )

    def forward(self, x):
        return self.model[1](self.model[0](x), residual=x)

    def forward_infer(self, x):
        return self.model[1].forward_infer(self.model[0].forward_infer(x), residual=x)


class ResBlocks(nn.Module):
    """``climategan.blocks.ResBlocks`` (blocks.py:153-171)."""

    def __init__(self, num_blocks, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = nn.Sequential(*[ResBlock(dim, norm=norm, activation=activation, pad_type=pad_type)
                                     for _ in range(num_blocks)])

    def forward(self, x):
        for blk in self.model:
            x = blk(x)
        return x

    def forward_infer(self, x):
        for blk in self.model:
            x = blk.forward_infer(x)
        return x


class BaseDecoder(nn.Module):
    """``climategan.blocks.BaseDecoder`` (blocks.py:206-318) without the v3 low-level-feature branch: proj 1x1 ->
    ResBlocks -> n_upsample x [nearest x2, 3x3 halving conv] -> 3x3 output conv."""

    def __init__(self, n_upsample=4, n_res=4, input_dim=2048, proj_dim=64, output_dim=3, norm="batch", activ="relu",
                 pad_type="zero", output_activ="tanh", low_level_feats_dim=-1, use_dada=False):
        super().__init__()
        self.low_level_feats_dim = low_level_feats_dim
        self.use_dada = use_dada
        if proj_dim != -1:
            self.proj_conv = Conv2dBlock(input_dim, proj_dim, 1, 1, 0, norm=norm, activation=activ)
        else:
            self.proj_conv = None
            proj_dim = input_dim
        if low_level_feats_dim > 0:   # deeplabv3 encoder: the backbone's layer1 features join the latent (blocks.py:237-258)
            self.low_level_conv = Conv2dBlock(input_dim=low_level_feats_dim, output_dim=proj_dim, kernel_size=3, stride=1,
                                              padding=1, pad_type=pad_type, norm=norm, activation=activ)
            self.merge_feats_conv = Conv2dBlock(input_dim=2 * proj_dim, output_dim=proj_dim, kernel_size=1, stride=1, padding=0,
                                                pad_type=pad_type, norm=norm, activation=activ)
        else:
            self.low_level_conv = None
        model = [ResBlocks(n_res, proj_dim, norm, activ, pad_type=pad_type)]
        dim = proj_dim
        for _ in range(n_upsample):
            model += [InterpolateNearest2d(scale_factor=2),
                      Conv2dBlock(input_dim=dim, output_dim=dim // 2, kernel_size=3, stride=1, padding=1, pad_type=pad_type,
                                  norm=norm, activation=activ)]
            dim //= 2
        model += [Conv2dBlock(input_dim=dim, output_dim=output_dim, kernel_size=3, stride=1, padding=1, pad_type=pad_type,
                              norm="none", activation=output_activ)]
        self.model = nn.Sequential(*model)
        self.output_dim = output_dim

    def forward_storage(self, z, cond=None, z_depth=None):
        """blocks.py:291-318.  Train mode (or grad enabled on a training module): autograd forwards; eval: fused inference."""
        import torch

        train = self.training
        run = (lambda blk, t: blk(t)) if train else (lambda blk, t: blk.forward_infer(t))
        low = None
        if isinstance(z, (list, tuple)):
            if self.low_level_conv is None:
                z = z[0]
            else:
                z, low = z
                low = run(self.low_level_conv, low)
                low = ops.resize_bilinear(low, z.shape[1], z.shape[2], align_corners=False)
        if z_depth is not None and self.use_dada:
            z = ops.mul(z, z_depth)
        if self.proj_conv is not None:
            z = run(self.proj_conv, z)
        if low is not None:
            z = run(self.merge_feats_conv, torch.cat([low, z], dim=-1))
        for m in self.model:
            if isinstance(m, InterpolateNearest2d) or train:
                z = m(z)
            else:
                z = m.forward_infer(z)
        return z
