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
            return Dict(v)
        if isinstance(v, (list, tuple)):
            return type(v)(Dict._wrap(i) for i in v)
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

    def update(self, *args, **kwargs):
        """addict.Dict.update: recursive merge of nested dicts."""
        other = {}
        if args:
            other.update(args[0])
        other.update(kwargs)
        for k, v in other.items():
            if k in self and isinstance(self[k], dict) and isinstance(v, dict):
                self[k].update(v)
            else:
                self[k] = self._wrap(v)

    def to_dict(self):
        def conv(v):
            if isinstance(v, Dict):
                return v.to_dict()
            if isinstance(v, (list, tuple)):
                return [conv(i) for i in v]
            return v

        return {k: conv(v) for k, v in self.items()}


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

    def update(self, *args, **kwargs):
        self._attach()
        Dict.update(self, *args, **kwargs)


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
                   activ="lrelu", n_res=1, proj_dim=32, pad_type="reflect",    # default-gen anchor (defaults.yaml:89-100)
                   classify=Dict(enable=False, linspace=Dict(min=0.35, max=6.95, buckets=256)))
    o.gen.s = Dict(architecture="deeplabv2", num_classes=11, output_dim=11, use_advent=True, use_minent=True,
                   upsample_featuremaps=False, use_dada=True, depth_feat_fusion=False, depth_dada_fusion=False)
    o.gen.m = Dict(use_spade=False, output_dim=1, n_res=3, n_upsample=3, proj_dim=64, norm="spectral", activ="lrelu",
                   pad_type="reflect", use_low_level_feats=True, use_dada=False, use_advent=True,
                   spade=Dict(latent_dim=128, detach=False, cond_nc=15, spade_use_spectral_norm=True,
                              spade_param_free_norm="batch", num_layers=3))
    return o


def full_opts(nblocks=(2, 2, 3, 2), size=128, latent=16, n_up=4, ndf=8, n_layers=3, num_d=2, tasks=("d", "s", "m", "p"),
              use_spade=False, overrides=None):
    """shared/trainer/defaults.yaml values on a small network (deeplabv2 encoder, as the north star names).  use_spade: the
    paper / release masker (MaskSpadeDecoder conditioned on make_m_cond(d, s, x), defaults.yaml:166-186)."""
    with_p = "p" in tasks
    o = default_masker_opts(tasks=tuple(t for t in tasks if t != "p"), nblocks=nblocks, size=size, with_painter=with_p,
                            latent_dim=latent, spade_n_up=n_up, ndf=ndf, n_layers=n_layers, num_D=num_d)
    o.tasks = list(tasks)
    o.domains = (["r", "s"] if any(t in tasks for t in "msd") else []) + (["rf"] if with_p else [])   # painter alone: rf only
    o.gen.default = Dict(init_type="xavier", init_gain=0.02)
    for k in ("encoder", "d", "s", "m", "p"):
        o.gen[k].init_type = "xavier"
        o.gen[k].init_gain = 0.02
    o.gen.m.use_minent = True
    o.gen.m.use_minent_var = True
    o.gen.m.use_ground_intersection = True
    o.gen.m.use_pl4m = False
    o.gen.m.use_proj = True
    if use_spade:
        o.gen.m.use_spade = True
        o.gen.m.spade.activations = Dict(all_lrelu=True)
    o.gen.p.pl4m_epoch = 49
    o.gen.opt.lr = Dict(default=0.00005)
    o.dis.soft_shift = 0.0
    o.dis.flip_prob = 0.0
    base = dict(input_nc=3, ndf=64, n_layers=4, norm="instance", init_type="xavier", init_gain=0.02, use_sigmoid=False,
                num_D=1, get_intermediate_features=False)
    o.dis.m = Dict(base, multi_level=False, architecture="base", gan_type="WGAN_norm", wgan_clamp_lower=-0.01,
                   wgan_clamp_upper=0.01)
    o.dis.s = Dict(base, gan_type="WGAN_norm", wgan_clamp_lower=-0.01, wgan_clamp_upper=0.01)
    # shared/trainer/events.yaml
    o.events = Dict(fire=Dict(kernel_size=281, kernel_sigma=140.5, transparency=200, sky_inc_factor=0.12, contrast_factor=1.5,
                              brightness_factor=0.95, crop_bottom_sky_mask=True),
                    smog=Dict(airlight=0.76, beta=2, vr=1, yellow_color=[224, 192, 29], alpha=20))
    o.train = Dict(
        amp=False, kitti=Dict(pretrain=False), pseudo=Dict(tasks=[], epochs=10), log_level=0, latent_domain_adaptation=False,
        lambdas=Dict(
            G=Dict(d=Dict(main=1, gml=0.5), s=Dict(crossent=1, crossent_pseudo=0.001, minent=0.001, advent=0.001),
                   m=Dict(bce=1, tv=1, gi=0.05, pl4m=1),
                   p=Dict(context=0, dm=1, featmatch=10, gan=1, reconstruction=0, tv=0, vgg=10)),
            advent=Dict(ent_main=0.5, ent_aux=0.0, ent_var=0.1, adv_main=1.0, adv_aux=0.0, dis_main=1.0, dis_aux=0.0, WGAN_gp=10)))
    for dotted, value in (overrides or {}).items():   # "gen.d.architecture": "base" — the reference's test-scenario notation
        node = o
        *path, leaf = dotted.split(".")
        for k in path:
            node = node[k]
        node[leaf] = value
    return o


def synth_batch(opts, batch, size, seed):
    """Synthetic multi_domain_batch of the shape Trainer.update_G/update_D consume: x ~ U(-1,1), m in {0,1},
    s int64 labels at size/4, d ~ U(0,1) at size/4."""
    import numpy as np
    import torch

    rs = np.random.RandomState(seed)
    out = {}
    q = size // 4
    for dom in opts.domains:
        data = {"x": torch.from_numpy((rs.random_sample((batch, 3, size, size)) * 2 - 1).astype(np.float32)),
                "m": torch.from_numpy((rs.random_sample((batch, 1, size // 8, size // 8)) > 0.5).astype(np.float32))
                .repeat_interleave(8, 2).repeat_interleave(8, 3)}
        if dom != "rf":
            data["s"] = torch.from_numpy(rs.randint(0, 11, size=(batch, 1, q, q)).astype(np.int64))
            data["d"] = torch.from_numpy(rs.random_sample((batch, 1, q, q)).astype(np.float32))
            if opts.gen.d.classify.enable and dom == "s":   # transforms.BucketizeDepth (:264-291): bucket indices on the sim domain
                data["d"] = torch.from_numpy(rs.randint(0, opts.gen.d.classify.linspace.buckets, size=(batch, 1, q, q)).astype(np.int64))
            for task in ("s", "d", "m"):   # a loader only yields the tasks being trained (data.py:446-484); the draws above
                if task not in opts.tasks:      # keep the random stream identical whatever the task list
                    del data[task]
        out[dom] = {"data": data, "domain": [dom] * batch, "mode": ["train"] * batch, "paths": {}}
    return out
