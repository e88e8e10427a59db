"""Import the UNMODIFIED reference modules from /root/reference on CPU (build container only).

The reference's ``climategan/__init__.py`` eagerly imports every submodule (and with them
comet_ml, addict, kornia, skimage, hydra ... none installed here), so we pre-seed
``sys.modules['climategan']`` with a bare namespace package pointing at the reference directory
and import only the hot-path submodules.  Nothing from the reference is copied: the modules are
executed where they lie.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CLIMATEGAN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "climategan"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    """Empty stand-ins for third-party modules the reference imports at module scope but the hot
    path never calls: addict.Dict (re-stated), comet_ml, skimage.io."""
    from climategan_b200.utils import Dict

    _stub("addict", Dict=Dict)
    _stub("comet_ml", Experiment=type("Experiment", (), {}), ExistingExperiment=type("ExistingExperiment", (), {}))
    _stub("torch_optimizer", NovoGrad=type("NovoGrad", (), {}), RAdam=type("RAdam", (), {}))
    sk = _stub("skimage")
    sk.io = _stub("skimage.io")
    sk.color = _stub("skimage.color")
    sk.transform = _stub("skimage.transform")
    sk.filters = _stub("skimage.filters")
    sk.io.imread = sk.color.rgba2rgb = sk.transform.resize = None
    # needed only to IMPORT climategan.trainer / data / eval_metrics / fire (never called on the oracle's paths, except
    # kornia's two Gaussian-blur helpers, restated from their published definition — SURVEY.md §8c row (ii))
    _stub("imageio", imread=None)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    _stub("seaborn")
    _install_kornia_stub()


def _install_kornia_stub():
    """kornia 0.5.10 ``filters.filter2D`` / ``filters.kernels.get_gaussian_kernel2d`` as used by climategan/fire.py:9-12,
    108-111 — restated: kernel = outer product of two normalised 1-D Gaussians exp(-(i - k//2)^2 / (2 sigma^2)) (even sizes
    shifted by half a pixel as kornia's ``gaussian`` does); filter2D = reflect-pad by k//2 + per-channel cross-correlation
    (``normalized=False``)."""
    import torch
    import torch.nn.functional as F

    def gaussian(window_size, sigma):
        x = torch.arange(window_size, dtype=torch.float32) - window_size // 2
        if window_size % 2 == 0:
            x = x + 0.5
        g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
        return g / g.sum()

    def get_gaussian_kernel2d(kernel_size, sigma, force_even=False):
        kx, ky = gaussian(kernel_size[1], sigma[1]), gaussian(kernel_size[0], sigma[0])
        return torch.matmul(ky.unsqueeze(-1), kx.unsqueeze(-1).t())

    def filter2D(input, kernel, border_type="reflect", normalized=False):
        b, c, h, w = input.shape
        k = kernel.to(input)
        kh, kw = k.shape[-2:]
        x = F.pad(input, (kw // 2, kw // 2, kh // 2, kh // 2), mode=border_type)
        out = F.conv2d(x, k.expand(c, 1, kh, kw).contiguous(), groups=c)
        return out[..., :h, :w]

    k = _stub("kornia")
    k.filters = _stub("kornia.filters", filter2D=filter2D, filter2d=filter2D)
    k.filters.kernels = _stub("kornia.filters.kernels", get_gaussian_kernel2d=get_gaussian_kernel2d)


def load(*submodules: str):
    """Return the requested reference submodules, e.g. load('painter', 'blocks', 'norms')."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if "climategan" not in sys.modules or not getattr(sys.modules["climategan"], "_cgb_shim", False):
        pkg = types.ModuleType("climategan")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "climategan")]
        pkg._cgb_shim = True
        sys.modules["climategan"] = pkg
    out = [importlib.import_module("climategan." + s) for s in submodules]
    return out[0] if len(out) == 1 else out
