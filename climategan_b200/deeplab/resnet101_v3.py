"""DeepLab-v3+ ResNet-101 backbone (``climategan/deeplab/resnet101_v3.py``): same module tree and state_dict keys.  Unlike the
v2 encoder the stride sits on the 3x3 conv (:13-21), the stem pools with padding 1 (:75), layer4 is a multi-grid unit with
dilations blocks[i] * dilation (:131-170), BatchNorm parameters are trainable, and the forward returns
``(z, low_level_feat)`` = (layer4, layer1) (:176-187).  Forwards run on NHWC storage tensors: eval = every conv with its BatchNorm
folded (one launch per conv+BN+ReLU(+residual)); train = conv, then batch-statistics BatchNorm passes."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops
from .resnetmulti_v2 import fold_bn


def conv_bn(x, conv: nn.Conv2d, bn: nn.BatchNorm2d, act=_lib.ACT_NONE, residual=None, training=False):
    """act(bn(conv(x)) (+ residual)) — the unit every DeepLab-v3 block is made of."""
    k, s, d, p = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
    if training:
        return ops.conv_bn_act(x, conv.weight, bn, conv.bias, residual, stride=s, dil=d, pad=p, act=act)
    w, b = fold_bn(conv, bn, x.dtype, cis=x.shape[-1])
    return ops.conv2d_infer(x, w, b, residual, k=k, stride=s, dil=d, pad=p, act=act, res_before_act=1)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, BatchNorm=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = BatchNorm(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, dilation=dilation, padding=dilation, bias=False)
        self.bn2 = BatchNorm(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = BatchNorm(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward_storage(self, x):
        t = self.training
        out = conv_bn(x, self.conv1, self.bn1, _lib.ACT_RELU, training=t)
        out = conv_bn(out, self.conv2, self.bn2, _lib.ACT_RELU, training=t)
        residual = x
        if self.downsample is not None:
            residual = conv_bn(x, self.downsample[0], self.downsample[1], _lib.ACT_NONE, training=t)
        return conv_bn(out, self.conv3, self.bn3, _lib.ACT_RELU, residual=residual, training=t)


class ResNet(nn.Module):
    def __init__(self, block, layers, output_stride, BatchNorm, verbose=0, no_init=False):
        self.inplanes = 64
        self.verbose = verbose
        super().__init__()
        blocks = [1, 2, 4]
        if output_stride == 16:
            strides, dilations = [1, 2, 2, 1], [1, 1, 1, 2]
        elif output_stride == 8:
            strides, dilations = [1, 2, 1, 1], [1, 1, 2, 4]
        else:
            raise NotImplementedError
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = BatchNorm(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], stride=strides[0], dilation=dilations[0], BatchNorm=BatchNorm)
        self.layer2 = self._make_layer(block, 128, layers[1], stride=strides[1], dilation=dilations[1], BatchNorm=BatchNorm)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=strides[2], dilation=dilations[2], BatchNorm=BatchNorm)
        self.layer4 = self._make_MG_unit(block, 512, blocks=blocks, stride=strides[3], dilation=dilations[3], BatchNorm=BatchNorm)

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1, BatchNorm=None):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                                       BatchNorm(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, dilation, downsample, BatchNorm)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation, BatchNorm=BatchNorm))
        return nn.Sequential(*layers)

    def _make_MG_unit(self, block, planes, blocks, stride=1, dilation=1, BatchNorm=None):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                                       BatchNorm(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, dilation=blocks[0] * dilation, downsample=downsample, BatchNorm=BatchNorm)]
        self.inplanes = planes * block.expansion
        for i in range(1, len(blocks)):
            layers.append(block(self.inplanes, planes, stride=1, dilation=blocks[i] * dilation, BatchNorm=BatchNorm))
        return nn.Sequential(*layers)

    def forward_storage(self, x):
        """x: storage [N,H,W,8] (3 real channels) -> (z [N,H/8,W/8,2048], low_level_feat [N,H/4,W/4,256]) at output stride 8."""
        t = self.training
        c1 = self.conv1
        k, cin = c1.kernel_size[0], c1.in_channels
        xc = ops.im2col_strided(x, cin, k, c1.padding[0], 1, c1.stride[0])   # stem as one K=152 GEMM (see resnetmulti_v2)
        kk = k * k * cin

        def as_gemm(w):
            w2 = w.permute(0, 2, 3, 1).reshape(w.shape[0], kk)
            return torch.nn.functional.pad(w2, (0, xc.shape[-1] - kk)).view(w.shape[0], xc.shape[-1], 1, 1)

        if t:
            y = ops.conv_bn_act(xc, as_gemm(c1.weight), self.bn1, act=_lib.ACT_RELU)
        else:
            from .resnetmulti_v2 import eval_bn_fold

            w, b = eval_bn_fold(c1.weight.detach(), None if c1.bias is None else c1.bias.detach(), self.bn1)
            wp = ops.pack_weight(as_gemm(w), x.dtype, cis=xc.shape[-1])
            y = ops.conv2d_infer(xc, wp, ops.pad_bias(b, wp.shape[0]), k=1, act=_lib.ACT_RELU)
        y = ops.maxpool3s2_pad1(y)
        for blk in self.layer1:
            y = blk.forward_storage(y)
        low = y
        for layer in (self.layer2, self.layer3, self.layer4):
            for blk in layer:
                y = blk.forward_storage(y)
        return y, low


def ResNet101(output_stride=8, BatchNorm=nn.BatchNorm2d, verbose=0, no_init=False, layers=(3, 4, 23, 3)):
    """resnet101_v3.py:190-203 (``layers`` is exposed so the parity fixtures can use a shallow copy of the same architecture)."""
    return ResNet(Bottleneck, list(layers), output_stride, BatchNorm, verbose=verbose, no_init=no_init)
