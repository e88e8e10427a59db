"""OmniDiscriminator — drop-in for ``climategan/discriminator.py``: same classes, constructor arguments, child-module
names (``discriminator_%d`` / ``model%d``) and state_dict keys; forward keeps the reference contract (NCHW fp32 in,
``list[num_D]`` of ``list[n_layers+2]`` NCHW fp32 feature maps out) and computes on NHWC storage tensors through
libcgb200: 4x4 spectral-norm convs on the tcgen05 engine (stride 2 through TMA element strides), leaky-relu in the conv
epilogue, InstanceNorm+LeakyReLU as one kernel, the 3x3/s2 average pool between scales.
"""
from __future__ import annotations

import functools

import torch
import torch.nn as nn
from torch.nn import init

from . import _lib, ops
from .norms import SpectralNorm, conv_weight_bias


def init_weights(net, init_type="normal", init_gain=0.02, verbose=0, caller=""):
    """Same selection rule as ``climategan/tutils.py:26-85``: modules whose class name contains Conv/Linear *and* that
    expose ``.weight`` are initialised — spectral-norm-wrapped convs expose ``weight_bar`` instead and are therefore
    left at PyTorch's default init, exactly as in the reference."""
    init_type = init_type or "normal"
    init_gain = init_gain or 0.02

    def init_func(m):
        classname = m.__class__.__name__
        if classname.find("BatchNorm2d") != -1:
            if hasattr(m, "weight") and m.weight is not None:
                init.normal_(m.weight.data, 1.0, init_gain)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "xavier_uniform":
                init.xavier_uniform_(m.weight.data, gain=1.0)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            elif init_type == "none":
                m.reset_parameters()
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)

    net.apply(init_func)


def create_discriminator(opts, device, no_init=False, verbose=0, storage_dtype=torch.bfloat16):
    """discriminator.py:16-39."""
    disc = OmniDiscriminator(opts, storage_dtype=storage_dtype)
    if no_init:
        return disc
    for task, model in disc.items():
        if isinstance(model, nn.ModuleDict):
            for domain, domain_model in model.items():
                init_weights(domain_model, init_type=opts.dis[task].init_type, init_gain=opts.dis[task].init_gain)
        else:
            init_weights(model, init_type=opts.dis[task].init_type, init_gain=opts.dis[task].init_gain)
    return disc.to(device)


def get_norm_layer(norm_type="instance"):
    """discriminator.py:63-78."""
    if not norm_type:
        norm_type = "instance"
    if norm_type == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == "none":
        return None
    raise NotImplementedError("normalization layer [%s] is not built (only instance/none)" % norm_type)


def define_D(input_nc, ndf, n_layers=3, norm="batch", use_sigmoid=False, get_intermediate_features=False, num_D=1,
             storage_dtype=torch.bfloat16):
    """discriminator.py:42-60."""
    return MultiscaleDiscriminator(input_nc, ndf, n_layers=n_layers, norm_layer=get_norm_layer(norm),
                                   use_sigmoid=use_sigmoid, get_intermediate_features=get_intermediate_features,
                                   num_D=num_D, storage_dtype=storage_dtype)


def _run_group(group: nn.Sequential, x):
    """One ``model%d`` group on a storage tensor: SpectralNorm(conv) [-> InstanceNorm2d] [-> LeakyReLU] [-> Sigmoid]."""
    conv = group[0]
    inner = conv.module if isinstance(conv, SpectralNorm) else conv
    has_norm = any(isinstance(m, nn.InstanceNorm2d) for m in group)
    has_lrelu = any(isinstance(m, nn.LeakyReLU) for m in group)
    has_sigmoid = any(isinstance(m, nn.Sigmoid) for m in group)
    w, b = conv_weight_bias(conv)
    act = _lib.ACT_NONE
    if not has_norm:
        act = _lib.ACT_LRELU if has_lrelu else (_lib.ACT_SIGMOID if has_sigmoid else _lib.ACT_NONE)
    if inner.in_channels <= 4:   # model0 on the image (+ mask): im2col + one short-K GEMM (ops.conv2d_first_layer)
        y = ops.conv2d_first_layer(x, w, b, stride=inner.stride[0], pad=inner.padding[0], act=act, slope=0.2)
    else:
        y = ops.conv2d(x, w, b, stride=inner.stride[0], pad=inner.padding[0], act=act, slope=0.2)
    if has_norm:
        y = ops.instnorm_act(y, _lib.ACT_LRELU if has_lrelu else _lib.ACT_NONE, 0.2)
    return y


class NLayerDiscriminator(nn.Module):
    """discriminator.py:82-182 (PatchGAN; every conv spectrally normalised; groups expose intermediate features)."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False,
                 get_intermediate_features=True):
        super().__init__()
        if norm_layer is None:
            raise NotImplementedError("NLayerDiscriminator without a norm layer is not built")
        if type(norm_layer) == functools.partial:
            use_bias = norm_layer.func == nn.InstanceNorm2d
        else:
            use_bias = norm_layer == nn.InstanceNorm2d
        self.get_intermediate_features = get_intermediate_features
        self.input_nc = input_nc
        kw, padw = 4, 1
        sequence = [[SpectralNorm(nn.Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw)), nn.LeakyReLU(0.2, True)]]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev = nf_mult
            nf_mult = min(2 ** n, 8)
            sequence += [[SpectralNorm(nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw,
                                                 bias=use_bias)), norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]]
        nf_mult_prev = nf_mult
        nf_mult = min(2 ** n_layers, 8)
        sequence += [[SpectralNorm(nn.Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw,
                                             bias=use_bias)), norm_layer(ndf * nf_mult), nn.LeakyReLU(0.2, True)]]
        sequence += [[SpectralNorm(nn.Conv2d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw))]]
        if use_sigmoid:
            sequence += [[nn.Sigmoid()]]
        for n in range(len(sequence)):
            self.add_module("model" + str(n), nn.Sequential(*sequence[n]))

    def out_channels(self):
        return [(g[0].module if isinstance(g[0], SpectralNorm) else g[0]).out_channels for g in self.children()
                if isinstance(g[0], (SpectralNorm, nn.Conv2d))]

    def forward_storage(self, x):
        results = []
        for group in self.children():
            if not isinstance(group[0], (SpectralNorm, nn.Conv2d)):
                raise NotImplementedError("stand-alone sigmoid group (use_sigmoid=True) is not built")
            x = _run_group(group, x)
            results.append(x)
        return results if self.get_intermediate_features else results[-1]


class MultiscaleDiscriminator(nn.Module):
    """discriminator.py:190-239."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False,
                 get_intermediate_features=True, num_D=3, storage_dtype=torch.bfloat16):
        super().__init__()
        self.n_layers, self.ndf, self.norm_layer = n_layers, ndf, norm_layer
        self.use_sigmoid, self.get_intermediate_features, self.num_D = use_sigmoid, get_intermediate_features, num_D
        self.storage_dtype = storage_dtype
        for i in range(self.num_D):
            self.add_module("discriminator_%d" % i,
                            NLayerDiscriminator(input_nc=input_nc, ndf=ndf, n_layers=n_layers, norm_layer=norm_layer,
                                                use_sigmoid=use_sigmoid, get_intermediate_features=get_intermediate_features))
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)  # holds no state

    def forward_storage(self, x):
        """x: storage [N,H,W,round8(input_nc)] -> list[num_D] of lists of storage feature maps."""
        result = []
        for name, D in self.named_children():
            if "discriminator" not in name:
                continue
            out = D.forward_storage(x)
            result.append(out if self.get_intermediate_features else [out])
            x = ops.avgpool3s2(x)
        return result

    def forward(self, input):
        feats = self.forward_storage(ops.to_storage(input, self.storage_dtype))
        out = []
        for (name, D), fl in zip([(n, d) for n, d in self.named_children() if "discriminator" in n], feats):
            chans = D.out_channels() if self.get_intermediate_features else [D.out_channels()[-1]]
            out.append([ops.from_storage(f, c) for f, c in zip(fl, chans)])
        return out


def get_fc_discriminator(num_classes=2, ndf=64, use_norm=False):
    """discriminator.py:327-361 (AdvEnt D for the masker's m / s heads): parameter container with the reference's
    Sequential layout; run it with :func:`fc_discriminator_forward`."""
    def conv(i, o):
        c = nn.Conv2d(i, o, kernel_size=4, stride=2, padding=1)
        return SpectralNorm(c) if use_norm else c

    return nn.Sequential(
        conv(num_classes, ndf), nn.LeakyReLU(0.2, inplace=True),
        conv(ndf, ndf * 2), nn.LeakyReLU(0.2, inplace=True),
        conv(ndf * 2, ndf * 4), nn.LeakyReLU(0.2, inplace=True),
        conv(ndf * 4, ndf * 8), nn.LeakyReLU(0.2, inplace=True),
        conv(ndf * 8, 1),
    )


def fc_discriminator_forward(net: nn.Sequential, x_nchw, storage_dtype=torch.bfloat16):
    """Forward of :func:`get_fc_discriminator` through libcgb200 (NCHW fp32 in / out)."""
    x = ops.to_storage(x_nchw, storage_dtype)
    mods = list(net)
    for i, m in enumerate(mods):
        if isinstance(m, (SpectralNorm, nn.Conv2d)):
            w, b = conv_weight_bias(m)
            lrelu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.LeakyReLU)
            inner = m.module if isinstance(m, SpectralNorm) else m
            conv = ops.conv2d_first_layer if inner.in_channels <= 4 else ops.conv2d   # the entropy-map input: im2col + one GEMM
            x = conv(x, w, b, stride=2, pad=1, act=_lib.ACT_LRELU if lrelu else _lib.ACT_NONE, slope=0.2)
    return ops.from_storage(x, 1)


class OmniDiscriminator(nn.ModuleDict):
    """discriminator.py:242-324."""

    def __init__(self, opts, storage_dtype=torch.bfloat16):
        super().__init__()
        if "p" in opts.tasks:
            if opts.dis.p.use_local_discriminator:   # discriminator.py:246-270: a global D on the image, a local one on image*mask
                kw = dict(input_nc=3, ndf=opts.dis.p.ndf, n_layers=opts.dis.p.n_layers, norm=opts.dis.p.norm,
                          use_sigmoid=opts.dis.p.use_sigmoid, get_intermediate_features=opts.dis.p.get_intermediate_features,
                          num_D=opts.dis.p.num_D, storage_dtype=storage_dtype)
                self["p"] = nn.ModuleDict({"global": define_D(**kw), "local": define_D(**kw)})
            else:
                self["p"] = define_D(input_nc=4, ndf=opts.dis.p.ndf, n_layers=opts.dis.p.n_layers, norm=opts.dis.p.norm,
                                     use_sigmoid=opts.dis.p.use_sigmoid,
                                     get_intermediate_features=opts.dis.p.get_intermediate_features, num_D=opts.dis.p.num_D,
                                     storage_dtype=storage_dtype)
        if "m" in opts.tasks and opts.gen.m.use_advent:
            if opts.dis.m.architecture != "base":
                raise NotImplementedError("dis.m.architecture=OmniDiscriminator is not built")
            self["m"] = nn.ModuleDict({"Advent": get_fc_discriminator(num_classes=2, use_norm=opts.dis.m.gan_type == "WGAN_norm")})
        if "s" in opts.tasks and opts.gen.s.use_advent:
            self["s"] = nn.ModuleDict({"Advent": get_fc_discriminator(num_classes=11, use_norm=opts.dis.s.gan_type == "WGAN_norm")})
