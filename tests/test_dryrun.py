"""Host logic of the train step on CPU: every configuration the GPU tests pin numerically is also driven end to end against
a stand-in for libcgb200 that type-checks each argument against the declared C signature and computes nothing
(tests/dryrun.py).  Catches what needs no GPU to catch: wrong argument counts / types at the C boundary, layout or shape
mistakes between layers (the storage-tensor checks stay in force), missing autograd edges, parameters that never receive a
gradient, and a call sequence that is not the same from one step to the next."""
import json
import os

import pytest
import torch

from climategan_b200.trainer import Trainer
from climategan_b200.utils import full_opts, synth_batch
from tests.dryrun import noop_library
from tests.helpers import GOLDEN

CASES = ["full_step", "full_step_pl4m", "masker_step_spade", "masker_step_base_depth_classify", "masker_step_v3",
         "mask_only_step_v3"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_train_step_host_path_runs_against_the_noop_library(case, dtype):
    meta = json.load(open(os.path.join(GOLDEN, case + ".json")))
    size, batch = meta["size"], meta["batch"]
    opts = full_opts(size=size, tasks=tuple(meta.get("tasks", ("d", "s", "m", "p"))), use_spade=meta.get("use_spade", False),
                     overrides=meta.get("overrides"))
    with noop_library() as lib:
        t = Trainer(opts, device=torch.device("cpu"), storage_dtype=dtype).setup(input_shape=(size, size))
        t.use_pl4m = bool(meta.get("pl4m", False))
        mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, batch, size, meta["seeds"]["inputs"]).items()}
        per_step = []
        for _ in range(3):
            before = dict(lib.calls)
            t.update_G(mdb)
            if _ == 0:
                # every parameter the reference gives a gradient to gets one here (the golden stores -1 where it does not)
                import numpy as np

                ref = np.load(os.path.join(GOLDEN, case + ".npz"))["G.gradnorm"]
                names = meta["g_param_names"]
                gp = dict(t.G.named_parameters())
                missing = [n for n, r in zip(names, ref) if r >= 0 and gp[n].requires_grad and gp[n].grad is None]
                assert not missing, missing[:10]
            t.update_D(mdb)
            t.logger.global_step += 1
            per_step.append({k: v - before.get(k, 0) for k, v in lib.calls.items() if v != before.get(k, 0)})
        # extrapolation step and update step enqueue the same work; so does every later step
        assert per_step[1] == per_step[2], {k: (per_step[1].get(k), per_step[2].get(k)) for k in set(per_step[1]) | set(per_step[2])
                                            if per_step[1].get(k) != per_step[2].get(k)}
        assert per_step[1]["cgb_conv2d_fwd"] > 50 and per_step[1]["cgb_extra_adam"] >= 2
        logs = t.losses_to_host()
        assert "total_loss" in logs["gen"] and "total_loss" in logs["disc"]
        assert set(k for k in meta["logs"][0] if k.startswith("gen.task")) <= set(_flat(logs))


def _flat(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flat(v, prefix + k + "."))
        else:
            out[prefix + k] = v
    return out


_V3 = {"gen.encoder.architecture": "deeplabv3", "gen.s.architecture": "deeplabv3", "gen.deeplabv3.nblocks": [2, 2, 3, 2]}
_BASE_D = {"gen.d.architecture": "base", "gen.m.use_dada": False, "gen.s.use_dada": False}
MORE = {
    # configurations without a golden of their own (the reference's scenario matrix, tests/test_trainer.py:205-308, and
    # option combinations around it): the host path must run and enqueue the same work every step
    "dada_ms": dict(tasks=("d", "s", "m"), overrides={"gen.m.use_dada": True}),
    "v3_spade_msdp": dict(tasks=("d", "s", "m", "p"), use_spade=True, overrides=dict(_V3)),
    "base_depth_regression": dict(tasks=("d", "s", "m"), overrides=dict(_BASE_D)),
    "v3_base_depth_low_level": dict(tasks=("d", "s", "m"), overrides=dict(_V3, **_BASE_D, **{"gen.d.use_low_level_feats": True})),
    "painter_only": dict(tasks=("p",)),
    "spade_cond12_step": dict(tasks=("d", "s", "m"), use_spade=True, overrides={"gen.m.spade.cond_nc": 12}),
    "spade_detached_cond": dict(tasks=("d", "s", "m"), use_spade=True, overrides={"gen.m.spade.detach": True}),
    "no_adversarial_losses": dict(tasks=("d", "s", "m"), overrides={"gen.m.use_advent": False, "gen.s.use_advent": False}),
    "depth_and_seg_only": dict(tasks=("d", "s")),
    "adam": dict(tasks=("d", "s", "m", "p"), overrides={"gen.opt.optimizer": "Adam", "dis.opt.optimizer": "Adam"}),
}


@pytest.mark.parametrize("name", sorted(MORE))
def test_other_configurations_run_the_host_path(name):
    opts = full_opts(size=128, **MORE[name])
    with noop_library() as lib:
        t = Trainer(opts, device=torch.device("cpu")).setup(input_shape=(128, 128))
        if name == "no_adversarial_losses":
            assert t.d_opt is None   # trainer.py:762-767
        mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, 2, 128, 3).items()}
        per_step = []
        for _ in range(3):
            before = dict(lib.calls)
            t.update_G(mdb)
            t.update_D(mdb)
            t.logger.global_step += 1
            per_step.append({k: v - before.get(k, 0) for k, v in lib.calls.items() if v != before.get(k, 0)})
        assert per_step[1] == per_step[2]
        trainable = [n for n, p in t.G.named_parameters() if p.requires_grad]
        no_grad = [n for n in trainable if dict(t.G.named_parameters())[n].grad is None]
        # every trainable generator parameter is reached by some loss (frozen BatchNorm affines are not trainable)
        assert not no_grad, no_grad[:10]


@pytest.mark.parametrize("name,kw", [("v2", {}), ("v3", dict(overrides=_V3)), ("v2_spade", dict(use_spade=True)),
                                     ("v3_spade", dict(use_spade=True, overrides=_V3))])
def test_infer_all_host_path(name, kw):
    """Trainer.infer_all (trainer.py:218-334) — masker, painter, flood / wildfire / smog compositing, with and without the
    cloudy sky — for the four masker configurations, tensors out (the uint8 edge needs pinned memory)."""
    opts = full_opts(size=128, **kw)
    with noop_library() as lib:
        t = Trainer(opts, device=torch.device("cpu")).setup(inference=True, input_shape=(128, 128))
        x = torch.rand(2, 3, 128, 128) * 2 - 1
        out = t.infer_all(x, numpy=False, bin_value=0.5)
        assert {k: tuple(v.shape) for k, v in out.items()} == {k: (2, 3, 128, 128) for k in ("flood", "wildfire", "smog")}
        n1 = sum(lib.calls.values())
        t.infer_all(x, numpy=False, bin_value=0.5)
        n2 = sum(lib.calls.values()) - n1
        t.infer_all(x, numpy=False, bin_value=0.5)
        n3 = sum(lib.calls.values()) - n1 - n2
        # the same work every call once the folded-BatchNorm packings are cached (the first call issues their cgb_pack_weight
        # launches; spectrally-normalised weights are re-packed on every call because their power iteration moves them)
        assert n2 == n3 and n2 <= n1 and lib.calls["cgb_pack_weight"] > 0
        out = t.infer_all(x, numpy=False, cloudy=True, ignore_event={"smog"})
        assert out["smog"] is None and out["flood"].shape == (2, 3, 128, 128)
