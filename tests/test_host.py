"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, option
container semantics, weight packing, module/state_dict surface."""
import os
import re

import pytest
import torch

from climategan_b200 import _lib, ops
from climategan_b200.painter import PainterSpadeDecoder
from climategan_b200.utils import Dict, default_painter_opts
from tests.helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cgb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()  # builds with nvcc if missing; loading needs no GPU
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cgb200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and include/cgb200.h disagree"
    assert b"sm_100a" in lib.cgb_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    """Without an sm_100 device every compute entry point refuses; the Python ops raise."""
    lib = _lib.lib()
    assert lib.cgb_device_ok() == 0
    d = _lib.ConvDesc(1, 8, 8, 8, 8, 8, 8, 3, 3, 1, 1, 1, 0, 0, 0, 0.2, 0, 0)
    import ctypes as C

    assert lib.cgb_conv2d_fwd(C.byref(d), None, None, None, None, None, None) == -4  # CGB_UNSUPPORTED_ARCH
    with pytest.raises(_lib.CgbError):
        ops.to_storage(torch.zeros(1, 3, 8, 8), torch.float32)
    with pytest.raises(_lib.CgbError):
        _lib.require_device()


def test_dict_semantics():
    d = Dict({"gen": {"p": {"latent_dim": 640}}})
    assert d.gen.p.latent_dim == 640
    assert not d.task  # addict: missing key -> empty falsy Dict (depth.py:12 relies on it)
    assert "task" not in d
    d.a.b.c = 3
    assert d.a.b.c == 3 and d.to_dict()["a"]["b"]["c"] == 3
    import copy

    e = copy.deepcopy(d)
    e.gen.p.latent_dim = 1
    assert d.gen.p.latent_dim == 640


def test_pack_unpack_roundtrip():
    w = torch.randn(20, 3, 3, 3)
    wp = ops.pack_weight(w, torch.float32)
    assert wp.shape == (24, 9, 8)
    assert float(wp[20:].abs().max()) == 0 and float(wp[:, :, 3:].abs().max()) == 0
    back = ops.unpack_weight_grad(wp, w.shape)
    assert torch.equal(back, w)
    assert ops.round8(20) == 24 and ops.round8(640) == 640 and ops.round8(3) == 8


def test_conv_geom():
    g = ops.ConvGeom(4, 4, stride=2, pad=1)
    assert g.out_hw(640, 640) == (320, 320)
    g = ops.ConvGeom(3, 3, dil=6, pad=6)
    assert g.out_hw(80, 80) == (80, 80)


def test_state_dict_surface_matches_reference():
    meta, _, sd, _ = load_golden()
    opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"])
    p = PainterSpadeDecoder(opts)
    mine = [(k, tuple(v.shape)) for k, v in p.state_dict().items()]
    assert mine == [(k, tuple(s)) for k, s in meta["shapes"]]
    p.load_state_dict(sd, strict=True)
    assert not p.head_0.conv_0.module.weight_u.requires_grad
    assert p.head_0.conv_0.module.weight_bar.requires_grad
    p.set_latent_shape(640, True)
    assert (p.z_h, p.z_w) == (640 // 2 ** meta["spade_n_up"],) * 2


def test_full_size_param_count():
    p = PainterSpadeDecoder(default_painter_opts())
    assert sum(v.numel() for v in p.state_dict().values()) == 42_341_735  # SURVEY.md §8a [probe]


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under climategan_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "climategan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), f"{f} references the oracle"
