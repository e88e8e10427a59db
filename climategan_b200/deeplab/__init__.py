"""DeepLab encoders / segmentation decoders — drop-in surface of ``climategan/deeplab/__init__.py`` for the v2 path the
north star names (``DeeplabV2Encoder`` = caffe-style ResNet-101 ``ResNetMulti``; ``DeepLabV2Decoder`` = ASPP head).
Forward (inference, eval-mode BatchNorm folded into the conv weights) runs on NHWC storage tensors through libcgb200.
The v3 / MobileNet variants are not built."""
from __future__ import annotations

import torch.nn as nn

from .deeplab_v2 import DeepLabV2Decoder
from .resnetmulti_v2 import ResNetMulti


def create_encoder(opts, no_init=False, verbose=0):
    if opts.gen.encoder.architecture == "deeplabv2":
        return DeeplabV2Encoder(opts, no_init, verbose)
    raise NotImplementedError("encoder architecture {} is not built (deeplabv2 only)".format(opts.gen.encoder.architecture))


def create_segmentation_decoder(opts, no_init=False, verbose=0):
    if opts.gen.s.architecture == "deeplabv2":
        return DeepLabV2Decoder(opts)
    raise NotImplementedError("segmentation architecture {} is not built (deeplabv2 only)".format(opts.gen.s.architecture))


class DeeplabV2Encoder(nn.Module):
    """deeplab/__init__.py:83-101 (pretrained-weight loading is out of scope: weights come from a checkpoint)."""

    def __init__(self, opts, no_init=False, verbose=0):
        super().__init__()
        self.model = ResNetMulti(opts.gen.deeplabv2.nblocks, opts.gen.encoder.n_res)

    def forward_storage(self, x):
        return self.model.forward_storage(x)
