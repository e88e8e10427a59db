"""Oracle for the v2 masker inference path (eval mode): functional fp32 restatement of

  climategan/deeplab/resnetmulti_v2.py  Bottleneck.forward :40-56, ResNetMulti.forward :126-136 (maxpool 3/s2/ceil :76-78)
  climategan/deeplab/deeplab_v2.py      _ASPPModule :27-31, ASPP.forward :110-123, DeepLabV2Decoder.forward :181-198
  climategan/depth.py                   DADADepthDecoder.forward :128-155
  climategan/blocks.py                  Conv2dBlock.forward :138-144, ResBlock :191-197, BaseDecoder.forward :291-318
  climategan/generator.py               decode :120-176, make_m_cond :196-230, mask :232-277 ; tutils.normalize :567-575

on a reference-layout state_dict.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle.painter_oracle import SNState


BN_TRAIN = [False]  # set through train_mode(): nn.BatchNorm2d in train mode = batch statistics + running-stat update


class train_mode:
    """with train_mode(): every bn() below behaves as nn.BatchNorm2d.train() (momentum 0.1, running stats updated in place
    in the state_dict, num_batches_tracked incremented) — what G.train() gives the reference's masker."""

    def __init__(self, on=True):
        self.on = on

    def __enter__(self):
        self.prev = BN_TRAIN[0]
        BN_TRAIN[0] = self.on

    def __exit__(self, *a):
        BN_TRAIN[0] = self.prev


def bn(sd, p, x):
    if BN_TRAIN[0]:
        if p + ".num_batches_tracked" in sd:
            sd[p + ".num_batches_tracked"] += 1
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def bottleneck(sd, p, x, stride, dilation):
    out = F.relu(bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=stride)))
    out = F.relu(bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], padding=dilation, dilation=dilation)))
    out = bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    residual = x
    if p + ".downsample.0.weight" in sd:
        residual = bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return F.relu(out + residual)


def encoder(sd, x, prefix="encoder.model"):
    x = F.relu(bn(sd, prefix + ".bn1", F.conv2d(x, sd[prefix + ".conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=0, ceil_mode=True)
    for li, (stride, dil) in enumerate([(1, 1), (2, 1), (1, 2), (1, 4)], start=1):
        n_blocks = len({k.split(".")[3] for k in sd if k.startswith(f"{prefix}.layer{li}.")})
        for b in range(n_blocks):
            x = bottleneck(sd, f"{prefix}.layer{li}.{b}", x, stride if b == 0 else 1, dil)
    return x


def conv2d_block(sd, sn, p, x, k, pad, pad_type, norm, act):
    """blocks.py:138-144."""
    if pad > 0:
        x = F.pad(x, (pad,) * 4, mode="reflect" if pad_type == "reflect" else "constant")
    if norm == "spectral":
        w, b = sn.weight(p + ".conv.module"), sd.get(p + ".conv.module.bias")
    else:
        w, b = sd[p + ".conv.weight"], sd.get(p + ".conv.bias")
    x = F.conv2d(x, w, b)
    if norm == "batch":
        x = bn(sd, p + ".norm", x)
    if act == "lrelu":
        x = F.leaky_relu(x, 0.2)
    elif act == "relu":
        x = F.relu(x)
    return x


def depth_decoder(sd, sn, z, target, p="decoders.d"):
    z4 = conv2d_block(sd, sn, p + ".enc4_1", z, 1, 0, "reflect", "batch", "lrelu")
    z4 = conv2d_block(sd, sn, p + ".enc4_2", z4, 3, 1, "reflect", "batch", "lrelu")
    z4 = conv2d_block(sd, sn, p + ".enc4_3", z4, 1, 0, "reflect", "batch", "lrelu")
    z_depth = conv2d_block(sd, sn, p + ".dec4", z4, 1, 0, "zero", "none", "lrelu") if p + ".dec4.conv.weight" in sd else None
    if p + ".upsample.1.conv.weight" in sd:
        y = F.interpolate(z4, size=(z4.shape[-2] * 2, z4.shape[-1] * 2), mode="nearest")
        y = conv2d_block(sd, sn, p + ".upsample.1", y, 3, 1, "reflect", "batch", "lrelu")
        z4 = F.conv2d(y, sd[p + ".upsample.2.weight"], sd[p + ".upsample.2.bias"])
    depth = torch.mean(z4, dim=1, keepdim=True)
    if depth.shape[-1] != target:
        depth = F.interpolate(depth, size=(384, 384), mode="bicubic", align_corners=False)
        depth = F.interpolate(depth, (target, target), mode="nearest")
    return depth, z_depth


def aspp_module(sd, p, x, k, dil):
    return F.relu(bn(sd, p + ".bn", F.conv2d(x, sd[p + ".atrous_conv.weight"], padding=0 if k == 1 else dil, dilation=dil)))


def seg_decoder(sd, z, z_depth, target, use_dada=True, p="decoders.s"):
    if z_depth is not None and use_dada:
        z = z * z_depth
    a = p + ".aspp"
    x1 = aspp_module(sd, a + ".aspp1", z, 1, 1)
    x2 = aspp_module(sd, a + ".aspp2", z, 3, 6)
    x3 = aspp_module(sd, a + ".aspp3", z, 3, 12)
    x4 = aspp_module(sd, a + ".aspp4", z, 3, 18)
    x5 = F.adaptive_avg_pool2d(z, (1, 1))
    x5 = F.relu(bn(sd, a + ".global_avg_pool.2", F.conv2d(x5, sd[a + ".global_avg_pool.1.weight"])))
    x5 = F.interpolate(x5, size=x4.shape[2:], mode="bilinear", align_corners=True)
    y = torch.cat((x1, x2, x3, x4, x5), dim=1)
    y = F.relu(bn(sd, a + ".bn1", F.conv2d(y, sd[a + ".conv1.weight"])))
    y = F.relu(bn(sd, p + ".conv.1", F.conv2d(y, sd[p + ".conv.0.weight"], padding=1)))
    y = F.relu(bn(sd, p + ".conv.5", F.conv2d(y, sd[p + ".conv.4.weight"], padding=1)))
    y = F.conv2d(y, sd[p + ".conv.8.weight"], sd[p + ".conv.8.bias"])
    return F.interpolate(y, target, mode="bilinear", align_corners=True)


def mask_decoder(sd, sn, z, n_res=3, n_upsample=3, p="decoders.m"):
    """BaseDecoder with norm='spectral', activ lrelu, reflect pad, no low-level feats, no dada."""
    z = conv2d_block(sd, sn, p + ".proj_conv", z, 1, 0, "zero", "spectral", "lrelu")
    for r in range(n_res):
        q = f"{p}.model.0.model.{r}.model"
        y = conv2d_block(sd, sn, q + ".0", z, 3, 1, "reflect", "spectral", "lrelu")
        y = conv2d_block(sd, sn, q + ".1", y, 3, 1, "reflect", "spectral", "none")
        z = y + z
    for u in range(n_upsample):
        z = F.interpolate(z, size=(z.shape[-2] * 2, z.shape[-1] * 2), mode="nearest")
        z = conv2d_block(sd, sn, f"{p}.model.{2 + 2 * u}", z, 3, 1, "reflect", "spectral", "lrelu")
    return conv2d_block(sd, sn, f"{p}.model.{1 + 2 * n_upsample}", z, 3, 1, "reflect", "none", "none")


def normalize(t):
    b = t.shape[0]
    min_t = t.reshape(b, -1).min(1)[0].reshape(b, 1, 1, 1)
    t = t - min_t
    max_t = t.reshape(b, -1).max(1)[0].reshape(b, 1, 1, 1)
    return t / max_t


def make_m_cond(d, s, x):
    return torch.cat([normalize(d), torch.softmax(s, dim=1),
                      F.interpolate(x, s.shape[-2:], mode="bilinear", align_corners=True)], dim=1)


def decode(sd, x, d_target, s_target, sn=None):
    sn = sn or SNState(sd)
    z = encoder(sd, x)
    d, z_depth = depth_decoder(sd, sn, z, d_target)
    s = seg_decoder(sd, z, z_depth, s_target)
    logits = mask_decoder(sd, sn, z)
    return {"z": z, "z_depth": z_depth, "d": d, "s": s, "m_logits": logits, "m": torch.sigmoid(logits)}


# ---- MaskSpadeDecoder (masker.py:59-231), eval mode: norms.SPADE with a BatchNorm(affine=False) param-free norm ------------
def spade_batch(sd, p, x, seg):
    """norms.py:174-186 with param_free_norm = nn.BatchNorm2d(affine=False) read from running statistics (eval)."""
    normalized = F.batch_norm(x, sd[p + ".param_free_norm.running_mean"], sd[p + ".param_free_norm.running_var"], None, None,
                              False, 0.0, 1e-5)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = F.relu(F.conv2d(seg, sd[p + ".mlp_shared.0.weight"], sd[p + ".mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


def spade_resblock_batch(sd, sn, p, x, seg):
    """blocks.py:369-392 with spectral-norm convs and last_activation lrelu; shortcut first (power-iteration order)."""
    if p + ".conv_s.module.weight_bar" in sd:
        x_s = F.conv2d(spade_batch(sd, p + ".norm_s", x, seg), sn.weight(p + ".conv_s.module"))
    else:
        x_s = x
    dx = F.conv2d(F.leaky_relu(spade_batch(sd, p + ".norm_0", x, seg), 0.2), sn.weight(p + ".conv_0.module"),
                  sd[p + ".conv_0.module.bias"], padding=1)
    dx = F.conv2d(F.leaky_relu(spade_batch(sd, p + ".norm_1", dx, seg), 0.2), sn.weight(p + ".conv_1.module"),
                  sd[p + ".conv_1.module.bias"], padding=1)
    return F.leaky_relu(x_s + dx, 0.2)


def mask_spade_decoder(sd, sn, z, cond, num_layers=3, p="decoders.m"):
    y = F.pad(z, (1, 1, 1, 1), mode="reflect")
    y = F.conv2d(y, sn.weight(p + ".fc_conv.conv.module"), sd[p + ".fc_conv.conv.module.bias"])
    y = F.leaky_relu(bn(sd, p + ".fc_conv.norm", y), 0.2)
    for i in range(num_layers):
        y = spade_resblock_batch(sd, sn, f"{p}.spade_blocks.{i}", y, cond)
        y = F.interpolate(y, size=(y.shape[-2] * 2, y.shape[-1] * 2), mode="nearest")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(y, sn.weight(p + ".mask_conv.conv.module"), sd[p + ".mask_conv.conv.module.bias"])


def decode_spade(sd, x, d_target, s_target, sn=None):
    """OmniGenerator.decode with gen.m.use_spade (generator.py:120-176)."""
    sn = sn or SNState(sd)
    z = encoder(sd, x)
    d, z_depth = depth_decoder(sd, sn, z, d_target)
    s = seg_decoder(sd, z, z_depth, s_target)
    cond = make_m_cond(d, s, x)
    logits = mask_spade_decoder(sd, sn, z, cond)
    return {"d": d, "s": s, "m": torch.sigmoid(logits)}
