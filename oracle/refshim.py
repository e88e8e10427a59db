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
