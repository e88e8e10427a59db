"""Oracle for the reference's default masker architecture (deeplabv3, ResNet backbone): functional fp32 restatement of

  climategan/deeplab/resnet101_v3.py  Bottleneck.forward :31-50, ResNet.forward :176-187 (maxpool 3/s2/pad1 :75, strides and
                                      dilations at output stride 8 :61-68, multi-grid layer4 :131-170)
  climategan/deeplab/deeplab_v3.py    ConvBNReLU.forward :50-53 (no ReLU), ASPPv3Plus.forward :92-105 (conv_out padding 1 :84),
                                      Decoder.forward :126-136, DeepLabV3Decoder.forward :244-266 (decoder called as (z_high, z_low))
  climategan/blocks.py                BaseDecoder.forward with low-level features :291-318
  climategan/depth.py                 DADADepthDecoder.forward :128-155 (z = z[0])

on a reference-layout state_dict; BatchNorm follows masker_oracle.train_mode().  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import masker_oracle as mo
from oracle.painter_oracle import SNState


def bottleneck(sd, p, x, stride, dilation):
    out = F.relu(mo.bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = F.relu(mo.bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=dilation, dilation=dilation)))
    out = mo.bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    residual = x
    if p + ".downsample.0.weight" in sd:
        residual = mo.bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + residual)


def encoder(sd, x, prefix="encoder"):
    """output stride 8: strides [1,2,1,1], dilations [1,1,2,4], layer4 multi-grid [1,2,4] * 4."""
    x = F.relu(mo.bn(sd, prefix + ".bn1", F.conv2d(x, sd[prefix + ".conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    low = None
    for li, (stride, dil) in enumerate([(1, 1), (2, 1), (1, 2), (1, 4)], start=1):
        n_blocks = len({k.split(".")[2] for k in sd if k.startswith(f"{prefix}.layer{li}.")})
        for b in range(n_blocks):
            d = dil * [1, 2, 4][b] if li == 4 else dil
            x = bottleneck(sd, f"{prefix}.layer{li}.{b}", x, stride if b == 0 else 1, d)
        if li == 1:
            low = x
    return x, low


def conv_bn(sd, p, x, pad, dil=1):
    return mo.bn(sd, p + ".bn", F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=pad, dilation=dil))


def seg_decoder(sd, z, z_depth, target, use_dada=True, p="decoders.s"):
    z_high, z_low = z
    if z_depth is not None and use_dada:
        z_high = z_high * z_depth
    a = p + ".aspp"
    feat = torch.cat([conv_bn(sd, a + ".conv1", z_high, 0), conv_bn(sd, a + ".conv2", z_high, 6, 6),
                      conv_bn(sd, a + ".conv3", z_high, 12, 12), conv_bn(sd, a + ".conv4", z_high, 18, 18)], 1)
    feat = conv_bn(sd, a + ".conv_out", feat, 1)                       # 1x1 conv with padding 1: H+2 x W+2
    dcd = p + ".decoder"
    feat_low = conv_bn(sd, dcd + ".conv_low", feat, 0)                 # Decoder(feat_low=ASPP output, feat_aspp=z_low)
    up = F.interpolate(z_low, feat.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat([feat_low, up], 1)
    y = conv_bn(sd, dcd + ".conv_cat.0", y, 1)
    y = conv_bn(sd, dcd + ".conv_cat.1", y, 1)
    logits = F.conv2d(y, sd[dcd + ".conv_out.weight"])
    return F.interpolate(logits, size=target, mode="bilinear", align_corners=True)


def mask_decoder(sd, sn, z, n_res=3, n_upsample=3, p="decoders.m"):
    """BaseDecoder with low-level features (norm='spectral', lrelu, reflect pad)."""
    z_high, low = z
    low = mo.conv2d_block(sd, sn, p + ".low_level_conv", low, 3, 1, "reflect", "spectral", "lrelu")
    low = F.interpolate(low, size=z_high.shape[-2:], mode="bilinear")
    y = mo.conv2d_block(sd, sn, p + ".proj_conv", z_high, 1, 0, "zero", "spectral", "lrelu")
    y = mo.conv2d_block(sd, sn, p + ".merge_feats_conv", torch.cat([low, y], 1), 1, 0, "reflect", "spectral", "lrelu")
    for r in range(n_res):
        q = f"{p}.model.0.model.{r}.model"
        t = mo.conv2d_block(sd, sn, q + ".0", y, 3, 1, "reflect", "spectral", "lrelu")
        t = mo.conv2d_block(sd, sn, q + ".1", t, 3, 1, "reflect", "spectral", "none")
        y = t + y
    for u in range(n_upsample):
        y = F.interpolate(y, size=(y.shape[-2] * 2, y.shape[-1] * 2), mode="nearest")
        y = mo.conv2d_block(sd, sn, f"{p}.model.{2 + 2 * u}", y, 3, 1, "reflect", "spectral", "lrelu")
    return mo.conv2d_block(sd, sn, f"{p}.model.{1 + 2 * n_upsample}", y, 3, 1, "reflect", "none", "none")


def forward(sd, x, d_target, s_target, sn=None):
    """encode + the three decoders in the order Trainer.infer_all / get_masker_loss run them (d, s, m)."""
    sn = sn or SNState(sd)
    z = encoder(sd, x)
    d, z_depth = mo.depth_decoder(sd, sn, z[0], d_target)
    s = seg_decoder(sd, z, z_depth, s_target)
    logits = mask_decoder(sd, sn, z)
    return {"d": d, "s": s, "m_logits": logits, "m": torch.sigmoid(logits)}
