"""SPADE painter — drop-in for ``climategan/painter.py`` (PainterSpadeDecoder, create_painter):
same constructor, attributes (``z_h``, ``z_w``, ``set_latent_shape``), submodule names and
state_dict keys.  ``forward(z, cond)`` keeps the reference contract — NCHW fp32 in, NCHW fp32
out — and runs every layer in between on NHWC storage tensors through libcgb200.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .blocks import InterpolateNearest2d, SPADEResnetBlock
from .norms import SpectralNorm, conv_weight_bias


def create_painter(opts, no_init=False, verbose=0):
    if verbose > 0:
        print("  - Add PainterSpadeDecoder Painter")
    return PainterSpadeDecoder(opts)


class PainterSpadeDecoder(nn.Module):
    """See ``climategan/painter.py:16-168``.  ``storage_dtype`` selects the activation storage
    type of the CUDA path (bf16 is the performance mode, fp32 the tight-parity mode)."""

    def __init__(self, opts, storage_dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        latent_dim = opts.gen.p.latent_dim
        cond_nc = 3
        spade_n_up = opts.gen.p.spade_n_up
        spade_use_spectral_norm = opts.gen.p.spade_use_spectral_norm
        spade_param_free_norm = opts.gen.p.spade_param_free_norm
        spade_kernel_size = 3
        self.storage_dtype = storage_dtype
        self.z_nc = latent_dim
        self.spade_n_up = spade_n_up
        self.z_h = self.z_w = None

        def block(fin, fout):
            return SPADEResnetBlock(fin, fout, cond_nc, spade_use_spectral_norm, spade_param_free_norm,
                                    spade_kernel_size)

        self.fc = nn.Conv2d(3, latent_dim, 3, padding=1)
        self.head_0 = block(self.z_nc, self.z_nc)
        self.G_middle_0 = block(self.z_nc, self.z_nc)
        self.G_middle_1 = block(self.z_nc, self.z_nc)
        self.up_spades = nn.Sequential(
            *[block(self.z_nc // 2 ** i, self.z_nc // 2 ** (i + 1)) for i in range(spade_n_up - 2)]
        )
        self.final_nc = self.z_nc // 2 ** (spade_n_up - 2)
        self.final_spade = block(self.final_nc, self.final_nc)
        self.final_shortcut = None
        if opts.gen.p.use_final_shortcut:   # painter.py:101-109: the last block is conditioned on a learned 3-channel map of y
            self.final_shortcut = nn.Sequential(SpectralNorm(nn.Conv2d(self.final_nc, 3, 1)), nn.BatchNorm2d(3),
                                                nn.LeakyReLU(0.2, True))
        self.conv_img = nn.Conv2d(self.final_nc, 3, 3, padding=1)
        self.upsample = InterpolateNearest2d(scale_factor=2)

    def set_latent_shape(self, shape, is_input=True):
        """painter.py:115-136."""
        if isinstance(shape, (list, tuple)):
            self.z_h = shape[-2]
            self.z_w = shape[-1]
        elif isinstance(shape, int):
            self.z_h = self.z_w = shape
        else:
            raise ValueError("Unknown shape type:", shape)
        if is_input:
            self.z_h = self.z_h // (2 ** self.spade_n_up)
            self.z_w = self.z_w // (2 ** self.spade_n_up)

    # -- storage-level forward (used by OmniGenerator.paint to skip a layout round trip) --------
    def forward_storage(self, cond_st: torch.Tensor, z_st: torch.Tensor = None) -> torch.Tensor:
        """cond_st: storage [N,H,W,8] conditioning; z_st: optional storage latent [N,z_h,z_w,round8(latent_dim)] (gen.p.no_z =
        false: OmniGenerator.sample_painter_z) replacing fc(cond).  Returns storage [N,H,W,8] holding tanh(conv_img)."""
        assert self.z_h is not None and self.z_w is not None
        segs = {}

        def seg_at(h, w):
            if (h, w) not in segs:
                segs[(h, w)] = cond_st if (h, w) == tuple(cond_st.shape[1:3]) else ops.resize_nearest(cond_st, h, w)
            return segs[(h, w)]

        cols = {}
        # a conditioning tensor that carries a gradient (painter loss for the masker: cond = x (1 - predicted mask)) takes the
        # direct 3x3 mlp_shared path in every SPADE layer — its dgrad is the gradient w.r.t. the conditioning
        cond_grad = torch.is_grad_enabled() and cond_st.requires_grad

        def run(blk, y):
            hw = (y.shape[1], y.shape[2])
            seg = seg_at(*hw)
            if cond_grad:
                return blk(y, seg)
            if hw not in cols:  # im2col patches of the conditioning: once per resolution, shared by all SPADEs
                cols[hw] = ops.im2col(seg, 3, 3, 1)
            return blk(y, seg, cols[hw])

        if z_st is None:   # painter.py:150-152
            z = ops.conv2d(seg_at(self.z_h, self.z_w), self.fc.weight, self.fc.bias, pad=1)
        else:
            z = z_st
        y = run(self.head_0, z)
        y = self.upsample(y)
        y = run(self.G_middle_0, y)
        y = self.upsample(y)
        y = run(self.G_middle_1, y)
        for up in self.up_spades:
            y = self.upsample(y)
            y = run(up, y)
        if self.final_shortcut is not None:
            # painter.py:163-164: cond <- lrelu(BatchNorm(SN conv1x1(y))), a DIFFERENTIABLE conditioning map: the block takes the
            # direct 3x3 mlp_shared path (SPADE's gradient w.r.t. the conditioning), not the shared im2col patches
            sn, bn = self.final_shortcut[0], self.final_shortcut[1]
            w_fs, b_fs = conv_weight_bias(sn)
            c_st = ops.batchnorm_act(ops.conv2d(y, w_fs, b_fs), bn, None, _lib.ACT_LRELU, 0.2)
            y = self.final_spade(y, c_st)
        else:
            y = run(self.final_spade, y)
        y = ops.activation(y, _lib.ACT_LRELU, 0.2)
        return ops.conv2d(y, self.conv_img.weight, self.conv_img.bias, pad=1, act=_lib.ACT_TANH)

    def forward(self, z, cond):
        cond_st = ops.to_storage(cond, self.storage_dtype)
        z_st = ops.to_storage(z, self.storage_dtype) if z is not None else None
        return ops.from_storage(self.forward_storage(cond_st, z_st), 3)
