"""Caffe-style ResNet-101 (``climategan/deeplab/resnetmulti_v2.py``): same module tree and state_dict keys.
Inference forward: every conv + its eval-mode BatchNorm (+ ReLU, + the bottleneck's residual add) is ONE libcgb200 conv
launch — BN folded into the packed weights/bias, ``relu(conv3 + residual)`` in the epilogue (``res_before_act``)."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops

affine_par = True


def eval_bn_fold(w, b, bn):
    """(w', b') of a conv weight / bias followed by eval-mode BatchNorm ``bn``: w' = w*g/sqrt(var+eps),
    b' = beta + (b - mean)*g/sqrt(var+eps).  ``bn`` may be the identity layer ``bn_fusion.bn_fuse`` leaves behind — the fold
    then already lives in the conv's own weight and bias."""
    from ..bn_fusion import is_fused

    if is_fused(bn):
        return w, b
    scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
    bb = bn.bias.detach() - bn.running_mean * scale
    if b is not None:
        bb = bb + b * scale
    return w * scale.view(-1, 1, 1, 1), bb


def fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d, dtype, cis=None):
    """Packed weights/bias of conv followed by eval-mode BatchNorm: w' = w*g/sqrt(var+eps), b' = beta + (b-mean)*g/sqrt(..)
    — what ``climategan/bn_fusion.py::fuse`` (:6-49) computes.  The folded packing is cached until any of the six tensors
    changes (ops.cached_pack), so repeated inference forwards do not re-fold."""
    from ..bn_fusion import is_fused

    def build():
        w, b = eval_bn_fold(conv.weight.detach(), None if conv.bias is None else conv.bias.detach(), bn)
        wp = ops.pack_weight(w, dtype, cis=cis)
        return wp, ops.pad_bias(b, wp.shape[0])

    deps = [conv.weight] + ([] if is_fused(bn) else [bn.weight, bn.bias, bn.running_mean, bn.running_var])
    deps += [conv.bias] if conv.bias is not None else []
    return ops.cached_pack(deps, ("fold_bn", dtype, cis), build, uses_running_stats=True)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, affine=affine_par)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=dilation, bias=False, dilation=dilation)
        self.bn2 = nn.BatchNorm2d(planes, affine=affine_par)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4, affine=affine_par)
        for bn in (self.bn1, self.bn2, self.bn3):
            for p in bn.parameters():
                p.requires_grad = False
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward_storage(self, x):
        if self.training:
            return self._forward_train(x)
        if isinstance(x, tuple):
            x = x[0]
        dt = x.dtype
        w1, b1 = fold_bn(self.conv1, self.bn1, dt, cis=x.shape[-1])
        out = ops.conv2d_infer(x, w1, b1, k=1, stride=self.stride, act=_lib.ACT_RELU)
        w2, b2 = fold_bn(self.conv2, self.bn2, dt, cis=out.shape[-1])
        out = ops.conv2d_infer(out, w2, b2, k=3, dil=self.dilation, pad=self.dilation, act=_lib.ACT_RELU)
        residual = x
        if self.downsample is not None:
            wd, bd = fold_bn(self.downsample[0], self.downsample[1], dt, cis=x.shape[-1])
            residual = ops.conv2d_infer(x, wd, bd, k=1, stride=self.downsample[0].stride[0])
        w3, b3 = fold_bn(self.conv3, self.bn3, dt, cis=out.shape[-1])
        return ops.conv2d_infer(out, w3, b3, residual, k=1, act=_lib.ACT_RELU, res_before_act=1)

    def _forward_train(self, x, dual=False):
        """resnetmulti_v2.py:40-56 in train mode: BatchNorm uses BATCH statistics (only its affine parameters are frozen,
        :16-18) and updates its running statistics; the conv's epilogue accumulates the statistics of its own output, so each
        conv -> BN -> ReLU is the conv launch + ONE normalise / ReLU (/ residual) pass (ops.conv_bn_act)."""
        # x may arrive as two aliases of the previous block's output (ops.batchnorm_act(dual=True)): one per consumer here
        x, x_skip = x if isinstance(x, tuple) else (x, x)
        if (self.downsample is None and self.stride == 1 and ops._CONV_SKIP and x.requires_grad and x.dtype != torch.float32
                and torch.is_grad_enabled()):
            # identity block: conv1 and the skip share x — the skip's gradient is added in conv1's dgrad epilogue (ops._Conv2dSkip)
            batch_stats = bool(self.bn1.training or self.bn1.running_mean is None)
            y1, partial, x_skip = ops.conv2d_skip(x, self.conv1.weight, want_stats=batch_stats)
            out = ops.batchnorm_act(y1, self.bn1, None, _lib.ACT_RELU, 0.2, partial=partial)
        else:
            out = ops.conv_bn_act(x, self.conv1.weight, self.bn1, stride=self.stride, act=_lib.ACT_RELU)
        out = ops.conv_bn_act(out, self.conv2.weight, self.bn2, dil=self.dilation, pad=self.dilation, act=_lib.ACT_RELU)
        residual = x_skip
        if self.downsample is not None:
            residual = ops.conv_bn_act(x_skip, self.downsample[0].weight, self.downsample[1], stride=self.downsample[0].stride[0])
        return ops.conv_bn_act(out, self.conv3.weight, self.bn3, residual=residual, act=_lib.ACT_RELU, dual=dual)


class ResNetMulti(nn.Module):
    def __init__(self, layers, n_res=4, res_norm="instance", activ="lrelu", pad_type="reflect"):
        super().__init__()
        self.inplanes = 64
        block = Bottleneck
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64, affine=affine_par)
        for p in self.bn1.parameters():
            p.requires_grad = False
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=0, ceil_mode=True)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=4)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                m.weight.data.normal_(0, 0.01)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        if n_res != 0:
            raise NotImplementedError("encoder.n_res > 0 (layer_res ResBlocks; 0 in defaults.yaml:105) is not built")
        self.layer_res = nn.Module()
        self.layer_res.model = nn.Sequential()  # ResBlocks(0, ...) holds an empty Sequential: no state

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion or dilation == 2 or dilation == 4:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm2d(planes * block.expansion, affine=affine_par),
            )
        for p in downsample._modules["1"].parameters():
            p.requires_grad = False
        layers = [block(self.inplanes, planes, stride, dilation=dilation, downsample=downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation))
        return nn.Sequential(*layers)

    def forward_storage(self, x):
        """x: storage [N,H,W,8] (3 real channels) -> z storage [N,H/8,W/8,2048]."""
        # stem: the 7x7 stride-2 conv on the 3-channel image runs as im2col (147 -> 152 channels) + ONE K=152 GEMM
        c1 = self.conv1
        k, cin = c1.kernel_size[0], c1.in_channels
        xc = ops.im2col_strided(x, cin, k, c1.padding[0], 1, c1.stride[0])
        kk = k * k * cin

        def as_gemm(w):   # [co, cin, k, k] -> [co, round8(k*k*cin), 1, 1] in the patches' tap-major order (autograd-native)
            w2 = w.permute(0, 2, 3, 1).reshape(w.shape[0], kk)
            return torch.nn.functional.pad(w2, (0, xc.shape[-1] - kk)).view(w.shape[0], xc.shape[-1], 1, 1)

        if self.training:
            x = ops.conv_bn_act(xc, as_gemm(c1.weight), self.bn1, act=_lib.ACT_RELU)
        else:
            w, b = eval_bn_fold(c1.weight.detach(), None if c1.bias is None else c1.bias.detach(), self.bn1)
            wp = ops.pack_weight(as_gemm(w), x.dtype, cis=xc.shape[-1])
            x = ops.conv2d_infer(xc, wp, ops.pad_bias(b, wp.shape[0]), k=1, act=_lib.ACT_RELU)
        x = ops.maxpool3s2_ceil(x)
        layers = (self.layer1, self.layer2, self.layer3, self.layer4)
        blocks = [blk for layer in layers for blk in layer]
        # every block's output but the last feeds exactly two consumers inside the next block (conv1 and the identity branch or
        # its down-sampling conv): hand it over as two aliases (ops.batchnorm_act(dual=True))
        dual_ok = self.training and ops._BN_DUAL and torch.is_grad_enabled() and x.dtype != torch.float32
        for i, blk in enumerate(blocks):
            if dual_ok and isinstance(blk, Bottleneck):
                x = blk._forward_train(x, dual=i + 1 < len(blocks))
            else:
                x = blk.forward_storage(x)
        return x
