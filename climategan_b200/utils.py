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
