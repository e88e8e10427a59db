"""DeepLab-v2 segmentation head (``climategan/deeplab/deeplab_v2.py``): ASPP (:43-123) + DeepLabV2Decoder (:136-198),
same module tree / state_dict keys; inference forward on storage tensors (BatchNorm folded, dropout = identity in eval)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops
from .resnetmulti_v2 import fold_bn


def find_target_size(opts, task):
    """climategan/utils.py:984-995."""
    new_size = opts.data.transforms[-1].new_size
    if isinstance(new_size, int):
        return new_size
    if task in new_size:
        return new_size[task]
    assert "default" in new_size
    return new_size["default"]


class _ASPPModule(nn.Module):
    def __init__(self, inplanes, planes, kernel_size, padding, dilation, BatchNorm, no_init):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                     dilation=dilation, bias=False)
        self.bn = BatchNorm(planes)
        self.relu = nn.ReLU()
        if not no_init:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    torch.nn.init.kaiming_normal_(m.weight)
                elif isinstance(m, nn.BatchNorm2d):
                    m.weight.data.fill_(1)
                    m.bias.data.zero_()

    def forward_storage(self, x):
        c = self.atrous_conv
        if self.training:
            return ops.conv_bn_act(x, c.weight, self.bn, dil=c.dilation[0], pad=c.padding[0], act=_lib.ACT_RELU)
        w, b = fold_bn(c, self.bn, x.dtype, cis=x.shape[-1])
        return ops.conv2d_infer(x, w, b, k=c.kernel_size[0], dil=c.dilation[0], pad=c.padding[0], act=_lib.ACT_RELU)


class ASPP(nn.Module):
    def __init__(self, backbone, output_stride, BatchNorm, no_init):
        super().__init__()
        inplanes = 320 if backbone == "mobilenet" else 2048
        if output_stride == 16:
            dilations = [1, 6, 12, 18]
        elif output_stride == 8:
            dilations = [1, 12, 24, 36]
        else:
            raise NotImplementedError
        self.aspp1 = _ASPPModule(inplanes, 256, 1, padding=0, dilation=dilations[0], BatchNorm=BatchNorm, no_init=no_init)
        self.aspp2 = _ASPPModule(inplanes, 256, 3, padding=dilations[1], dilation=dilations[1], BatchNorm=BatchNorm, no_init=no_init)
        self.aspp3 = _ASPPModule(inplanes, 256, 3, padding=dilations[2], dilation=dilations[2], BatchNorm=BatchNorm, no_init=no_init)
        self.aspp4 = _ASPPModule(inplanes, 256, 3, padding=dilations[3], dilation=dilations[3], BatchNorm=BatchNorm, no_init=no_init)
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                                             BatchNorm(256), nn.ReLU())
        self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
        self.bn1 = BatchNorm(256)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)
        if not no_init:
            for m in self.modules():
                if isinstance(m, nn.Conv2d):
                    torch.nn.init.kaiming_normal_(m.weight)
                elif isinstance(m, nn.BatchNorm2d):
                    m.weight.data.fill_(1)
                    m.bias.data.zero_()

    def forward_storage(self, x):
        n, h, w, _ = x.shape
        x1 = self.aspp1.forward_storage(x)
        x2 = self.aspp2.forward_storage(x)
        x3 = self.aspp3.forward_storage(x)
        x4 = self.aspp4.forward_storage(x)
        g = ops.global_mean(x)                                     # AdaptiveAvgPool2d(1)
        if self.training:
            x5 = ops.conv2d(g, self.global_avg_pool[1].weight, None)
            x5 = ops.batchnorm_act(x5, self.global_avg_pool[2], None, _lib.ACT_RELU)   # batch stats over N only (1x1 maps)
            x5 = ops.broadcast_hw(x5, h, w)                        # bilinear 1x1 -> h x w = broadcast (deeplab_v2.py:116)
            y = torch.cat((x1, x2, x3, x4, x5), dim=-1)
            y = ops.conv_bn_act(y, self.conv1.weight, self.bn1, act=_lib.ACT_RELU)
            return ops.dropout(y, self.dropout.p, True)
        wg, bg = fold_bn(self.global_avg_pool[1], self.global_avg_pool[2], x.dtype, cis=g.shape[-1])
        x5 = ops.conv2d_infer(g, wg, bg, k=1, act=_lib.ACT_RELU)
        x5 = ops.resize_bilinear(x5, h, w, align_corners=True)     # 1x1 -> h x w broadcast (deeplab_v2.py:116)
        y = torch.cat((x1, x2, x3, x4, x5), dim=-1)                # channel concat of NHWC tensors (a copy, no math)
        w1, b1 = fold_bn(self.conv1, self.bn1, y.dtype, cis=y.shape[-1])
        return ops.conv2d_infer(y, w1, b1, k=1, act=_lib.ACT_RELU)  # dropout: identity in eval


class DeepLabV2Decoder(nn.Module):
    def __init__(self, opts, no_init=False):
        super().__init__()
        self.aspp = ASPP("resnet", 16, nn.BatchNorm2d, no_init)
        self.use_dada = ("d" in opts.tasks) and opts.gen.s.use_dada
        conv_modules = [
            nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(), nn.Dropout(0.5),
            nn.Conv2d(256, 256, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(), nn.Dropout(0.1),
        ]
        self._c0 = 0   # index of the first conv in self.conv
        if opts.gen.s.upsample_featuremaps:   # deeplab_v2.py:154-155: nearest x2 in front of the head (state_dict keys shift by 1)
            from ..blocks import InterpolateNearest2d

            conv_modules = [InterpolateNearest2d(scale_factor=2)] + conv_modules
            self._c0 = 1
        conv_modules += [nn.Conv2d(256, opts.gen.s.output_dim, kernel_size=1, stride=1)]
        self.conv = nn.Sequential(*conv_modules)
        self.output_dim = opts.gen.s.output_dim
        self._target_size = find_target_size(opts, "s")

    def set_target_size(self, size):
        self._target_size = size[:2] if isinstance(size, (list, tuple)) else (size, size)

    def forward_storage(self, z, z_depth=None):
        if self._target_size is None:
            raise Exception("self._target_size should be set with self.set_target_size()")
        if z.shape[-1] != 2048:
            raise Exception("Segmentation decoder will only work with 2048 channels for z")
        if z_depth is not None and self.use_dada:
            z = ops.mul(z, z_depth)
        y = self.aspp.forward_storage(z)
        c0 = self._c0
        if c0:
            y = self.conv[0](y)
        last = self.conv[c0 + 8]
        if self.training:
            for i in (c0, c0 + 4):
                y = ops.conv_bn_act(y, self.conv[i].weight, self.conv[i + 1], pad=1, act=_lib.ACT_RELU)
                y = ops.dropout(y, self.conv[i + 3].p, True)
            y = ops.conv2d(y, last.weight, last.bias)
        else:
            for i in (c0, c0 + 4):
                w, b = fold_bn(self.conv[i], self.conv[i + 1], y.dtype, cis=y.shape[-1])
                y = ops.conv2d_infer(y, w, b, k=3, pad=1, act=_lib.ACT_RELU)
            wl = ops.pack_weight(last.weight, y.dtype, cis=y.shape[-1])
            y = ops.conv2d_infer(y, wl, ops.pad_bias(last.bias, wl.shape[0]), k=1)
        ts = self._target_size
        th, tw = (ts, ts) if isinstance(ts, int) else ts
        return ops.resize_bilinear(y, th, tw, align_corners=True)
