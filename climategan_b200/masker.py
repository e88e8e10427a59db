"""Flood-mask decoders (``climategan/masker.py``): MaskBaseDecoder (:25-56) = BaseDecoder (``blocks.py:206-318``) with
the masker options, and MaskSpadeDecoder (:59-231; the paper / release configuration): spectral-norm +
BatchNorm ``fc_conv`` (deeplabv2 latent) or low-level / high-level / merge convs (deeplabv3 pair), ``num_layers`` x [SPADEResnetBlock conditioned on make_m_cond's 15-channel tensor, nearest x2], spectral
``mask_conv``.  Eval mode reads the SPADE layers' BatchNorm running statistics through the fused inference kernels; train mode
runs them on batch statistics with a gradient into both the latent and the conditioning tensor."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .blocks import BaseDecoder, Conv2dBlock, InterpolateNearest2d, SPADEResnetBlock


def create_mask_decoder(opts, no_init=False, verbose=0):
    if opts.gen.m.use_spade:
        return MaskSpadeDecoder(opts)
    return MaskBaseDecoder(opts)


class MaskBaseDecoder(BaseDecoder):
    def __init__(self, opts):
        use_v3 = opts.gen.encoder.architecture == "deeplabv3"
        if use_v3 and opts.gen.deeplabv3.backbone == "mobilenet":
            raise NotImplementedError("the deeplabv3 mobilenet backbone is not built")
        low_level_feats_dim = 256 if (use_v3 and opts.gen.m.use_low_level_feats) else -1   # masker.py:27-44
        use_dada = ("d" in opts.tasks) and opts.gen.m.use_dada
        super().__init__(n_upsample=opts.gen.m.n_upsample, n_res=opts.gen.m.n_res, input_dim=2048,
                         proj_dim=opts.gen.m.proj_dim, output_dim=opts.gen.m.output_dim, norm=opts.gen.m.norm,
                         activ=opts.gen.m.activ, pad_type=opts.gen.m.pad_type, output_activ="none",
                         low_level_feats_dim=low_level_feats_dim, use_dada=use_dada)


class MaskSpadeDecoder(nn.Module):
    def __init__(self, opts):
        super().__init__()
        self.opts = opts
        sp = opts.gen.m.spade
        latent_dim, cond_nc = sp.latent_dim, sp.cond_nc
        spade_activation = "lrelu" if sp.activations.all_lrelu else None
        self.num_layers = sp.num_layers
        self.z_nc = latent_dim
        sb = dict(padding=1, activation="lrelu", pad_type="reflect", norm="spectral_batch")
        if opts.gen.encoder.architecture == "deeplabv3":     # masker.py:93-158: z = (latent, backbone layer1 features)
            if opts.gen.deeplabv3.backbone == "mobilenet":
                raise NotImplementedError("the deeplabv3 mobilenet backbone is not built")
            self.input_dim = [2048, 256]
            if opts.gen.m.use_proj:
                proj_dim = opts.gen.m.proj_dim
                self.low_level_conv = Conv2dBlock(self.input_dim[1], proj_dim, 3, **sb)
                self.high_level_conv = Conv2dBlock(self.input_dim[0], proj_dim, 3, **sb)
                self.merge_feats_conv = Conv2dBlock(proj_dim * 2, self.z_nc, 3, **sb)
            else:
                self.low_level_conv = Conv2dBlock(self.input_dim[1], self.input_dim[0], 3, **sb)
                self.merge_feats_conv = Conv2dBlock(self.input_dim[0] * 2, self.z_nc, 3, **sb)
        elif opts.gen.encoder.architecture == "deeplabv2":
            self.input_dim = 2048
            self.fc_conv = Conv2dBlock(self.input_dim, self.z_nc, 3, **sb)
        else:
            raise ValueError("Unknown encoder type")
        self.spade_blocks = nn.Sequential(*[
            SPADEResnetBlock(int(self.z_nc / (2 ** i)), int(self.z_nc / (2 ** (i + 1))), cond_nc, sp.spade_use_spectral_norm,
                             sp.spade_param_free_norm, 3, spade_activation) for i in range(self.num_layers)])
        self.final_nc = int(self.z_nc / (2 ** self.num_layers))
        self.mask_conv = Conv2dBlock(self.final_nc, 1, 3, padding=1, activation="none", pad_type="reflect", norm="spectral")
        self.upsample = InterpolateNearest2d(scale_factor=2)

    def _latent(self, z, run):
        """masker.py:213-224: merge the v3 (latent, low-level) pair, or fc_conv on the v2 latent."""
        if isinstance(z, (list, tuple)):
            z_h, z_l = z
            z_l = run(self.low_level_conv, z_l)
            z_l = ops.resize_bilinear(z_l, z_h.shape[1], z_h.shape[2], align_corners=False)
            if self.opts.gen.m.use_proj:
                z_h = run(self.high_level_conv, z_h)
            return run(self.merge_feats_conv, torch.cat([z_h, z_l], dim=-1))
        return run(self.fc_conv, z)

    def forward_storage(self, z, cond, z_depth=None):
        """masker.py:212-231.  z: storage [N,h,w,2048] (or the v3 pair); cond: NCHW fp32 conditioning from
        ``OmniGenerator.make_m_cond``."""
        if cond is None:
            raise ValueError("MaskSpadeDecoder needs the conditioning tensor (OmniGenerator.make_m_cond)")
        dt = (z[0] if isinstance(z, (list, tuple)) else z).dtype
        if self.training:
            # autograd forwards (the caller decides whether a tape is recorded): train-mode BatchNorm in the latent convs and in
            # every SPADE layer (batch statistics + running update), gradient into z AND into cond (gen.m.spade.detach = false)
            y = self._latent(z, lambda blk, t: blk(t))
            seg = ops.to_storage(cond, dt)
            for blk in self.spade_blocks:
                seg_r = seg if seg.shape[1:3] == y.shape[1:3] else ops.resize_nearest(seg, y.shape[1], y.shape[2])
                y = blk(y, seg_r)
                y = self.upsample(y)
            return self.mask_conv(y)
        with torch.no_grad():
            y = self._latent(z, lambda blk, t: blk.forward_infer(t))
            seg = ops.to_storage(cond, dt)
            for blk in self.spade_blocks:
                seg_r = seg if seg.shape[1:3] == y.shape[1:3] else ops.resize_nearest(seg, y.shape[1], y.shape[2])
                y = blk(y, seg_r)
                y = self.upsample(y)
            return self.mask_conv.forward_infer(y)
