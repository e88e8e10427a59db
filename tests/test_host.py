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
    wp = ops.pack_weight(w, torch.float32, kernel=False)   # the torch statement of the layout (the kernel path needs the device)
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


def _names_shapes(module):
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]


def test_generator_and_discriminator_surfaces_match_the_reference_for_every_built_configuration():
    """state_dict keys / shapes / order and the parameter order of OmniGenerator and OmniDiscriminator against the lists the
    REFERENCE modules produced (stored with the goldens), for every configuration that is built: v2 masker + painter (base mask
    decoder), the SPADE mask decoder (paper / release configuration), the base depth decoder with depth classification and the up-sampled segmentation
    head, the reference-default v3 masker with the base and with
    the SPADE mask decoder, and the painter with use_final_shortcut.  Construction needs no GPU."""
    import json

    from climategan_b200.discriminator import OmniDiscriminator
    from climategan_b200.generator import OmniGenerator
    from climategan_b200.utils import Dict, default_masker_opts, full_opts
    from tests.helpers import GOLDEN

    def meta_of(name):
        return json.load(open(os.path.join(GOLDEN, name + ".json")))

    # full step (tasks d, s, m, p), base and pl4m fixtures share the architecture; SPADE-masker step (tasks d, s, m)
    for name in ("full_step", "full_step_pl4m", "masker_step_spade", "masker_step_base_depth_classify", "masker_step_v3",
                 "mask_only_step_v3"):
        meta = meta_of(name)
        opts = full_opts(size=meta["size"], tasks=tuple(meta.get("tasks", ("d", "s", "m", "p"))),
                         use_spade=meta.get("use_spade", False), overrides=meta.get("overrides"))
        G = OmniGenerator(opts, latent_shape=(meta["size"], meta["size"]))
        D = OmniDiscriminator(opts)
        assert _names_shapes(G) == [(k, tuple(s)) for k, s in meta["g_shapes"]], name
        assert _names_shapes(D) == [(k, tuple(s)) for k, s in meta["d_shapes"]], name
        assert [k for k, _ in G.named_parameters()] == meta["g_param_names"], name
        assert [k for k, _ in D.named_parameters()] == meta["d_param_names"], name
    # v3 masker, base and SPADE mask decoders
    for name in ("masker_v3", "masker_v3_spade"):
        meta = meta_of(name)
        opts = default_masker_opts(nblocks=tuple(meta["nblocks"]), size=meta["size"])
        opts.gen.encoder.architecture = "deeplabv3"
        opts.gen.s.architecture = "deeplabv3"
        opts.gen.deeplabv3.nblocks = list(meta["nblocks"])
        if meta.get("use_spade"):
            opts.gen.m.use_spade = True
            opts.gen.m.use_proj = True
            opts.gen.m.spade.activations = Dict(all_lrelu=True)
        G = OmniGenerator(opts)
        assert _names_shapes(G) == [(k, tuple(s)) for k, s in meta["shapes"]], name
        assert [k for k, _ in G.named_parameters()] == meta["param_names"], name
    # painter with the final shortcut (and an explicit latent)
    meta = meta_of("painter_z_shortcut")
    opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"])
    opts.gen.p.no_z = False
    opts.gen.p.use_final_shortcut = True
    assert _names_shapes(PainterSpadeDecoder(opts)) == [(k, tuple(s)) for k, s in meta["shapes"]]


def test_make_m_cond_needs_x_for_15_channels():
    """generator.py:220-225: cond_nc == 15 without x is a ValueError (checked before any kernel runs)."""
    import pytest

    from climategan_b200.generator import OmniGenerator
    from climategan_b200.utils import Dict, default_masker_opts

    opts = default_masker_opts(nblocks=(1, 1, 1, 1), size=64)
    opts.gen.m.use_spade = True
    opts.gen.m.spade.activations = Dict(all_lrelu=True)
    G = OmniGenerator(opts)
    d, s = torch.zeros(1, 1, 16, 16), torch.zeros(1, 11, 16, 16)
    with pytest.raises(ValueError, match="x MUST be provided"):
        G.make_m_cond(d, s, None)


def test_run_epoch_and_train_follow_the_reference_control_flow():
    """Trainer.run_epoch / train (trainer.py:888-987) with the compute stubbed out: one update_G + update_D per multi-batch
    tuple and a global_step increment; no update_D (and no scheduler step) while the VKITTI2 pre-training is on; pl4m switches on
    at gen.p.pl4m_epoch; kitti pre-training and pseudo-label training end at their epoch counts."""
    import torch.nn as nn

    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts

    opts = full_opts(size=64)
    opts.gen.m.use_pl4m = True
    opts.gen.p.pl4m_epoch = 1
    opts.train.kitti = {"pretrain": True, "epochs": 2}
    opts.train.pseudo = {"tasks": ["d"], "epochs": 3}
    opts.train.epochs = 4
    t = Trainer(opts, device=torch.device("cpu"))
    assert t.kitti_pretrain and t.pseudo_training_tasks == {"d"}
    t.is_setup = True
    t.G, t.D = nn.Module(), nn.Module()
    t.G.painter = nn.Linear(2, 2)
    t.d_opt = object()
    calls, sched = [], []
    t.update_G = lambda mdb: calls.append(("G", tuple(sorted(mdb)), t.use_pl4m, t.kitti_pretrain))
    t.update_D = lambda mdb: calls.append(("D", tuple(sorted(mdb))))
    t.batch_to_device = lambda b: b
    t.update_learning_rates = lambda: sched.append(t.logger.epoch)
    batches = [tuple({"domain": [d], "data": {}} for d in ("r", "s", "rf")) for _ in range(2)]
    seen = []
    t.train(lambda epoch: (seen.append(epoch), batches)[1], on_epoch_end=lambda tr: seen.append(("end", tr.logger.epoch)))
    assert seen == [0, ("end", 0), 1, ("end", 1), 2, ("end", 2), 3, ("end", 3)]
    assert t.logger.global_step == 8 and t.logger.epoch == 3
    g_calls = [c for c in calls if c[0] == "G"]
    assert len(g_calls) == 8 and all(c[1] == ("r", "rf", "s") for c in g_calls)
    assert [c[2] for c in g_calls] == [False, False] + [True] * 6           # pl4m from epoch 1 on
    assert [c[3] for c in g_calls] == [True] * 4 + [False] * 4             # kitti pre-training during epochs 0 and 1
    assert len([c for c in calls if c[0] == "D"]) == 4                      # no update_D while pre-training
    assert sched == [2, 3]                                                  # ... and no scheduler step
    assert t.pseudo_training_tasks == set()
