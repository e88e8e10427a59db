"""Minimal option container: attribute-access nested dict with addict semantics (the reference
passes ``addict.Dict`` opts everywhere; a missing key yields an empty, falsy Dict)."""
from __future__ import annotations

import copy


class Dict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if a:
                for k, v in dict(a).items():
                    self[k] = self._wrap(v)
        for k, v in kwargs.items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, Dict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(i) for i in v)
        return v

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self[k]

    def __setattr__(self, k, v):
        self[k] = self._wrap(v)

    def __missing__(self, k):
        # addict: a missing key is an empty Dict that attaches itself on first assignment
        child = _Pending(self, k)
        return child

    def __deepcopy__(self, memo):
        return Dict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Dict) else v) for k, v in self.items()}


class _Pending(Dict):
    """Empty child returned for a missing key; writes through to the parent on first set."""

    def __init__(self, parent, key):
        dict.__init__(self)
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_key", key)

    def _attach(self):
        parent = object.__getattribute__(self, "_parent")
        key = object.__getattribute__(self, "_key")
        if isinstance(parent, _Pending):
            parent._attach()
        if key not in parent:
            dict.__setitem__(parent, key, self)

    def __setitem__(self, k, v):
        self._attach()
        dict.__setitem__(self, k, v)


def default_painter_opts(latent_dim=640, spade_n_up=7, tasks=("p",), ndf=64, n_layers=4, num_D=3):
    """The painter-related sections of shared/trainer/defaults.yaml (reference values): gen.p :141-160, gen.opt :73-90,
    dis :193-228, train.lambdas.G.p :293-300."""
    return Dict(
        tasks=list(tasks),
        dis=dict(
            soft_shift=0.2, flip_prob=0.05,
            opt=dict(optimizer="ExtraAdam", beta1=0.5, lr=dict(default=0.00002), lr_policy="step", lr_step_size=15, lr_gamma=0.5),
            p=dict(input_nc=3, ndf=ndf, n_layers=n_layers, norm="instance", init_type="xavier", init_gain=0.02,
                   use_sigmoid=False, num_D=num_D, get_intermediate_features=True, use_local_discriminator=False),
        ),
        train=dict(lambdas=dict(G=dict(p=dict(context=0, dm=1, featmatch=10, gan=1, reconstruction=0, tv=0, vgg=10)))),
        gen=dict(
            opt=dict(optimizer="ExtraAdam", beta1=0.9, lr=dict(default=0.00005), lr_policy="step", lr_step_size=5, lr_gamma=0.5),
            p=dict(
                latent_dim=latent_dim,
                loss="gan",
                no_z=True,
                output_dim=3,
                pad_type="reflect",
                paste_original_content=True,
                spade_kernel_size=3,
                spade_n_up=spade_n_up,
                spade_param_free_norm="instance",
                spade_use_spectral_norm=True,
                use_final_shortcut=False,
                diff_aug=dict(use=False),
            )
        ),
    )


def default_masker_opts(tasks=("m", "s", "d"), nblocks=(3, 4, 23, 3), size=640, with_painter=False, **painter_kw):
    """gen.{encoder,deeplabv2,d,s,m} of shared/trainer/defaults.yaml with the deeplabv2 architecture the north star names
    (defaults.yaml:100-192) and data.transforms[-1].new_size (:62-67: x 640, d/s 160 -> here size and size/4)."""
    o = default_painter_opts(tasks=tuple(tasks) + (("p",) if with_painter else ()), **painter_kw)
    o.data = Dict(transforms=[Dict(name="resize", new_size=Dict(default=size, d=size // 4, s=size // 4))])
    o.gen.encoder = Dict(architecture="deeplabv2", n_res=0, norm="spectral", activ="lrelu", pad_type="reflect")
    o.gen.deeplabv2 = Dict(nblocks=list(nblocks), use_pretrained=False)
    o.gen.deeplabv3 = Dict(backbone="resnet", output_stride=8)
    o.gen.d = Dict(architecture="dada", upsample_featuremaps=True, output_dim=1, norm="batch", loss="sigm",
                   classify=Dict(enable=False))
    o.gen.s = Dict(architecture="deeplabv2", num_classes=11, output_dim=11, use_advent=True, use_minent=True,
                   upsample_featuremaps=False, use_dada=True, depth_feat_fusion=False, depth_dada_fusion=False)
    o.gen.m = Dict(use_spade=False, output_dim=1, n_res=3, n_upsample=3, proj_dim=64, norm="spectral", activ="lrelu",
                   pad_type="reflect", use_low_level_feats=True, use_dada=False, use_advent=True,
                   spade=Dict(latent_dim=128, detach=False, cond_nc=15, spade_use_spectral_norm=True,
                              spade_param_free_norm="batch", num_layers=3))
    return o
