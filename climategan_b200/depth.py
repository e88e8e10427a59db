"""Depth decoders (``climategan/depth.py``): DADADepthDecoder (:25-158) and BaseDepthDecoder (:161-230), same module trees /
state_dict keys; forwards on storage tensors (eval: Conv2dBlock(norm="batch") = conv + folded BatchNorm + leaky-relu in one
launch; train: batch-statistics BatchNorm on the autograd tape)."""
from __future__ import annotations

import torch.nn as nn

from . import _lib, ops
from .blocks import BaseDecoder, Conv2dBlock, InterpolateNearest2d
from .deeplab.deeplab_v2 import find_target_size


def create_depth_decoder(opts, no_init=False, verbose=0):
    if opts.gen.d.architecture == "base":
        return BaseDepthDecoder(opts)
    return DADADepthDecoder(opts)


class BaseDepthDecoder(BaseDecoder):
    """``climategan.depth.BaseDepthDecoder`` (depth.py:161-230): a BaseDecoder (1x1 projection, ResBlocks, 0 or 1 nearest-x2 +
    3x3 halving conv, 3x3 head) followed by a bilinear (align_corners) resize to the depth target size.  One output channel
    (regression), or ``gen.d.classify.linspace.buckets`` logits when depth is classified (``classify.enable``; the loss is then
    a cross-entropy against bucketised log-depth, losses.py:399-405, transforms.py:264-291)."""

    def __init__(self, opts):
        low_level_feats_dim = -1
        if opts.gen.encoder.architecture == "deeplabv3":
            if opts.gen.deeplabv3.backbone == "mobilenet":
                raise NotImplementedError("the deeplabv3 mobilenet backbone is not built")
            if opts.gen.d.use_low_level_feats:
                low_level_feats_dim = 256
        output_dim = 1 if not opts.gen.d.classify.enable else opts.gen.d.classify.linspace.buckets
        super().__init__(n_upsample=1 if opts.gen.d.upsample_featuremaps else 0, n_res=opts.gen.d.n_res, input_dim=2048,
                         proj_dim=opts.gen.d.proj_dim, output_dim=output_dim, norm=opts.gen.d.norm, activ=opts.gen.d.activ,
                         pad_type=opts.gen.d.pad_type, output_activ="none", low_level_feats_dim=low_level_feats_dim)
        self._target_size = find_target_size(opts, "d")

    def set_target_size(self, size):
        self._target_size = size[:2] if isinstance(size, (list, tuple)) else (size, size)

    def forward_storage(self, z, cond=None, z_depth=None):
        """-> (depth / depth logits storage [N,T,T,round8(output_dim)], None): this decoder has no DADA feature to share."""
        if self._target_size is None:
            raise ValueError("self._target_size should be set with self.set_target_size()")
        d = BaseDecoder.forward_storage(self, z)
        ts = self._target_size
        th, tw = (ts, ts) if isinstance(ts, int) else ts
        return ops.resize_bilinear(d, th, tw, align_corners=True), None


class DADADepthDecoder(nn.Module):
    def __init__(self, opts):
        super().__init__()
        res_dim, mid_dim = 2048, 512
        self.do_feat_fusion = False
        if opts.gen.m.use_dada or ("s" in opts.tasks and opts.gen.s.use_dada):
            self.do_feat_fusion = True
            self.dec4 = Conv2dBlock(128, res_dim, 1, stride=1, padding=0, bias=True, activation="lrelu", norm="none")
        self.relu = nn.ReLU(inplace=True)
        kw = dict(stride=1, bias=False, activation="lrelu", pad_type="reflect", norm="batch")
        self.enc4_1 = Conv2dBlock(res_dim, mid_dim, 1, padding=0, **kw)
        self.enc4_2 = Conv2dBlock(mid_dim, mid_dim, 3, padding=1, **kw)
        self.enc4_3 = Conv2dBlock(mid_dim, 128, 1, padding=0, **kw)
        self.upsample = None
        if opts.gen.d.upsample_featuremaps:
            self.upsample = nn.Sequential(InterpolateNearest2d(), Conv2dBlock(128, 32, 3, padding=1, **kw),
                                          nn.Conv2d(32, 1, kernel_size=1, stride=1, padding=0))
        self._target_size = find_target_size(opts, "d")

    def set_target_size(self, size):
        self._target_size = size[:2] if isinstance(size, (list, tuple)) else (size, size)

    def forward_storage(self, z):
        """z storage [N,h,w,2048] -> (depth storage [N,T,T,8] (1 real channel), z_depth storage or None)."""
        if isinstance(z, (list, tuple)):   # deeplabv3 encoder: (z, low_level_feat) (depth.py:129-130)
            z = z[0]
        run = (lambda blk, t: blk(t)) if self.training else (lambda blk, t: blk.forward_infer(t))
        z4 = run(self.enc4_3, run(self.enc4_2, run(self.enc4_1, z)))
        z_depth = run(self.dec4, z4) if self.do_feat_fusion else None
        c_log = 128
        if self.upsample is not None:
            y = self.upsample[0](z4)
            y = run(self.upsample[1], y)
            last = self.upsample[2]
            if self.training:
                z4 = ops.conv2d(y, last.weight, last.bias)
            else:
                wl = ops.pack_weight(last.weight, y.dtype, cis=y.shape[-1])
                z4 = ops.conv2d_infer(y, wl, ops.pad_bias(last.bias, wl.shape[0]), k=1)
            c_log = 1
        depth = ops.channel_mean(z4, c_log)               # torch.mean(z4_enc, dim=1, keepdim=True)
        ts = self._target_size
        if depth.shape[2] != ts:                          # depth.py:143 compares the width with the int target size
            depth = ops.resize_bicubic(depth, 384, 384)   # MiDaS inference size, bicubic, align_corners=False
            depth = ops.resize_nearest(depth, ts, ts)
        return depth, z_depth
