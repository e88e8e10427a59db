"""DeepLab encoders / segmentation decoders — drop-in surface of ``climategan/deeplab/__init__.py``: the v2 path the north
star names (``DeeplabV2Encoder`` = caffe-style ResNet-101 ``ResNetMulti``; ``DeepLabV2Decoder`` = ASPP head) and the reference's
default v3 path with the ResNet backbone (``resnet101_v3.ResNet101``, ``DeepLabV3Decoder``).  Forwards run on NHWC storage
tensors through libcgb200 (eval: BatchNorm folded into the conv weights; train: batch statistics).  The MobileNet backbone is
not built."""
from __future__ import annotations

import torch.nn as nn

from .deeplab_v2 import DeepLabV2Decoder
from .deeplab_v3 import DeepLabV3Decoder
from .resnet101_v3 import ResNet101
from .resnetmulti_v2 import ResNetMulti


def create_encoder(opts, no_init=False, verbose=0):
    if opts.gen.encoder.architecture == "deeplabv2":
        return DeeplabV2Encoder(opts, no_init, verbose)
    if opts.gen.encoder.architecture == "deeplabv3":
        return build_v3_backbone(opts, no_init, verbose)
    raise NotImplementedError("Unknown encoder: {}".format(opts.gen.encoder.architecture))


def create_segmentation_decoder(opts, no_init=False, verbose=0):
    if opts.gen.s.architecture == "deeplabv2":
        return DeepLabV2Decoder(opts)
    if opts.gen.s.architecture == "deeplabv3":
        return DeepLabV3Decoder(opts, no_init=True)   # weights come from a checkpoint (pretrained files are out of scope)
    raise NotImplementedError("Unknown Segmentation architecture: {}".format(opts.gen.s.architecture))


def build_v3_backbone(opts, no_init, verbose=0):
    """deeplab/__init__.py:45-80 (ResNet backbone; the pretrained-file loading is out of scope).  ``gen.deeplabv3.nblocks``
    (not a reference option; default [3, 4, 23, 3]) lets the parity fixtures use a shallow copy of the architecture."""
    if opts.gen.deeplabv3.backbone != "resnet":
        raise NotImplementedError("deeplabv3 backbone '{}' is not built (resnet only)".format(opts.gen.deeplabv3.backbone))
    layers = tuple(opts.gen.deeplabv3.nblocks) if opts.gen.deeplabv3.nblocks else (3, 4, 23, 3)
    return ResNet101(output_stride=opts.gen.deeplabv3.output_stride, BatchNorm=nn.BatchNorm2d, verbose=verbose, no_init=no_init,
                     layers=layers)


class DeeplabV2Encoder(nn.Module):
    """deeplab/__init__.py:83-101 (pretrained-weight loading is out of scope: weights come from a checkpoint)."""

    def __init__(self, opts, no_init=False, verbose=0):
        super().__init__()
        self.model = ResNetMulti(opts.gen.deeplabv2.nblocks, opts.gen.encoder.n_res)

    def forward_storage(self, x):
        return self.model.forward_storage(x)
