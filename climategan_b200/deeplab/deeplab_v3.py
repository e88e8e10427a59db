"""DeepLab-v3+ segmentation head for the ResNet backbone (``climategan/deeplab/deeplab_v3.py``): ConvBNReLU (:33-62; conv
with bias + BatchNorm and — despite the name — NO ReLU), ASPPv3Plus (:65-112), Decoder (:115-136), DeepLabV3Decoder (:142-266).
Same module tree / state_dict keys.  Kept bug-compatible with the reference: ``ASPPv3Plus.conv_out`` is a 1x1
conv built with ConvBNReLU's default ``padding=1`` (:84), so an 80x80 map leaves the ASPP as 82x82; and
``DeepLabV3Decoder.forward`` calls ``self.decoder(z_high, z_low)`` (:258) while ``Decoder.forward(feat_low, feat_aspp)``
(:126) — so ``conv_low`` runs on the ASPP output and the backbone's low-level features are the ones resized to 82x82."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops
from .deeplab_v2 import find_target_size
from .resnet101_v3 import conv_bn


class ConvBNReLU(nn.Module):
    def __init__(self, in_chan, out_chan, ks=3, stride=1, padding=1, dilation=1, *args, **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_chan, out_chan, kernel_size=ks, stride=stride, padding=padding, dilation=dilation, bias=True)
        self.bn = nn.BatchNorm2d(out_chan)
        nn.init.kaiming_normal_(self.conv.weight, a=1)
        nn.init.constant_(self.conv.bias, 0)

    def forward_storage(self, x):
        return conv_bn(x, self.conv, self.bn, _lib.ACT_NONE, training=self.training)


class ASPPv3Plus(nn.Module):
    def __init__(self, backbone, no_init):
        super().__init__()
        in_chan = 320 if backbone == "mobilenet" else 2048
        self.with_gp = False
        self.conv1 = ConvBNReLU(in_chan, 256, ks=1, dilation=1, padding=0)
        self.conv2 = ConvBNReLU(in_chan, 256, ks=3, dilation=6, padding=6)
        self.conv3 = ConvBNReLU(in_chan, 256, ks=3, dilation=12, padding=12)
        self.conv4 = ConvBNReLU(in_chan, 256, ks=3, dilation=18, padding=18)
        self.conv_out = ConvBNReLU(256 * 4, 256, ks=1)   # padding defaults to 1: the output grows by 2 pixels (reference quirk)

    def forward_storage(self, x):
        feats = [m.forward_storage(x) for m in (self.conv1, self.conv2, self.conv3, self.conv4)]
        return self.conv_out.forward_storage(torch.cat(feats, dim=-1))


class Decoder(nn.Module):
    def __init__(self, n_classes):
        super().__init__()
        self.conv_low = ConvBNReLU(256, 48, ks=1, padding=0)
        self.conv_cat = nn.Sequential(ConvBNReLU(304, 256, ks=3, padding=1), ConvBNReLU(256, 256, ks=3, padding=1))
        self.conv_out = nn.Conv2d(256, n_classes, kernel_size=1, bias=False)

    def forward_storage(self, feat_low, feat_aspp):
        h, w = feat_low.shape[1:3]
        feat_low = self.conv_low.forward_storage(feat_low)
        feat_aspp_up = ops.resize_bilinear(feat_aspp, h, w, align_corners=True)
        y = torch.cat([feat_low, feat_aspp_up], dim=-1)
        for m in self.conv_cat:
            y = m.forward_storage(y)
        if self.training:
            return ops.conv2d(y, self.conv_out.weight, None)
        wp = ops.pack_weight_cached(self.conv_out.weight, y.dtype, cis=y.shape[-1])
        return ops.conv2d_infer(y, wp, None, k=1)


class DeepLabV3Decoder(nn.Module):
    def __init__(self, opts, no_init=False, freeze_bn=False):
        super().__init__()
        num_classes = opts.gen.s.output_dim
        self.backbone = opts.gen.deeplabv3.backbone
        self.use_dada = ("d" in opts.tasks) and opts.gen.s.use_dada
        if self.backbone != "resnet":
            raise NotImplementedError("DeepLabV3Decoder: only the resnet backbone is built (mobilenet head is not)")
        self.aspp = ASPPv3Plus(self.backbone, no_init)
        self.decoder = Decoder(num_classes)
        self.freeze_bn = freeze_bn
        self.output_dim = num_classes
        self._target_size = find_target_size(opts, "s")
        if not no_init:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    nn.init.kaiming_normal_(m.weight, mode="fan_out")
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)
                elif isinstance(m, nn.BatchNorm2d):
                    nn.init.ones_(m.weight)
                    nn.init.zeros_(m.bias)

    def set_target_size(self, size):
        self._target_size = size[:2] if isinstance(size, (list, tuple)) else (size, size)

    def forward_storage(self, z, z_depth=None):
        assert isinstance(z, (tuple, list))
        if self._target_size is None:
            raise ValueError("self._target_size should be set with self.set_target_size()")
        z_high, z_low = z
        if z_depth is not None and self.use_dada:
            z_high = ops.mul(z_high, z_depth)
        z_high = self.aspp.forward_storage(z_high)
        s = self.decoder.forward_storage(z_high, z_low)      # (feat_low=ASPP output, feat_aspp=low-level features): as called
        ts = self._target_size
        th, tw = (ts, ts) if isinstance(ts, int) else ts
        return ops.resize_bilinear(s, th, tw, align_corners=True)
