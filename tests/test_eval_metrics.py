"""Validation metrics (Trainer.eval_images trainer.py:1706-1799, eval_metrics.py:68-124) on CPU: the numpy
oracle against the fixture generated from the reference's own functions, and the product's host path (ops.argmax_confusion ->
eval_metrics -> Trainer.eval_images) against the same fixture through the emulated C ABI (tests/emulib.py)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from tests.golden.eval_cases import cases
from tests.helpers import GOLDEN

FIXTURE = json.load(open(os.path.join(GOLDEN, "eval_metrics.json")))["cases"]


def check_case(name, pred, label, kind, accuracy, miou):
    """Shared with the GPU suite.  The metrics are ratios of integer counts: equal to the last bit or two of a double."""
    ref = FIXTURE[name]
    prob = torch.cat([1 - pred, pred], dim=1) if kind == "mask" else pred
    assert accuracy(pred, label) == pytest.approx(ref["accuracy"], rel=1e-14, abs=0)
    assert miou(prob, label) == pytest.approx(ref["mIOU"], rel=1e-14, abs=0)
    assert miou(prob, label, "weighted") == pytest.approx(ref["mIOU_weighted"], rel=1e-14, abs=0)


@pytest.mark.parametrize("name", sorted(FIXTURE))
def test_oracle_matches_reference_fixture(name):
    from oracle import eval_metrics_oracle as o

    pred, label, kind = cases()[name]
    check_case(name, pred, label, kind, lambda p, g: o.accuracy(p.numpy(), g.numpy()),
               lambda p, g, average="macro": o.miou(p.numpy(), g.numpy(), average))


def test_fixture_regenerates_from_the_reference():
    from oracle import refshim

    if not refshim.available():
        pytest.skip("reference tree absent")
    ref = refshim.load("eval_metrics")
    for name, (pred, label, kind) in cases().items():
        prob = torch.cat([1 - pred, pred], dim=1) if kind == "mask" else pred
        assert float(ref.accuracy(pred, label)) == FIXTURE[name]["accuracy"]
        assert float(ref.mIOU(prob, label)) == FIXTURE[name]["mIOU"]


@pytest.mark.parametrize("name", sorted(FIXTURE))
def test_product_host_path_on_the_emulated_abi(name):
    from climategan_b200 import eval_metrics as em
    from tests.emulib import emulated_library

    pred, label, kind = cases()[name]
    with emulated_library() as lib:
        check_case(name, pred, label, kind, em.accuracy, em.mIOU)
        conf, lmax = em.confusion(pred, label)
        assert conf.sum() == label.numel() and lmax == int(label.max())
        assert lib.calls["cgb_argmax_confusion"] >= 4


def test_empty_classes_give_nan_and_bad_shapes_raise():
    from climategan_b200 import eval_metrics as em
    from tests.emulib import emulated_library

    conf = np.zeros((3, 4), np.int64)
    assert math.isnan(em.miou_from_confusion(conf, 0))
    with emulated_library():
        with pytest.raises(ValueError):
            em.mIOU(torch.zeros(1, 3, 4, 4), torch.zeros(1, 1, 4, 5, dtype=torch.int64))
        with pytest.raises(NotImplementedError):
            em.accuracy(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))


def test_trainer_eval_images_on_the_emulated_abi():
    from tests.emulib import emulated_library

    with emulated_library():
        check_trainer_eval_images(torch.device("cpu"))


def check_trainer_eval_images(device):
    """Trainer.eval_images against the oracle metrics of the same trainer's own predictions (one image at a time, d on the sim
    domain only, mask metrics -1 as in the reference — see the method's docstring), and the early returns (:1707-1711).
    Shared with the GPU suite."""
    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts, synth_batch
    from oracle import eval_metrics_oracle as o

    size = 64
    opts = full_opts(size=size, tasks=("d", "s", "m"), overrides={"gen.d.architecture": "base", "gen.d.classify.enable": True,
                                                                  "gen.d.classify.linspace.buckets": 16, "gen.m.use_dada": False,
                                                                  "gen.s.use_dada": False, "gen.s.upsample_featuremaps": True})
    if True:
        torch.manual_seed(0)
        t = Trainer(opts, device=device, storage_dtype=torch.float32).setup(inference=True, input_shape=(size, size))
        t.G.eval()
        batch = synth_batch(opts, 2, size, 5)["s"]["data"]
        sets = [{"data": {k: v[i] for k, v in batch.items()}} for i in range(2)]
        t.display_images = {"val": {"s": sets}}
        assert t.eval_images("val", "rf") is None and t.eval_images("val", "r") is None and t.eval_images("train", "s") is None
        assert t.eval_images("val", "s") == 0
        got = t.metrics["metrics_val_s"]
        acc = {"s": [], "d": []}
        iou = {"s": [], "d": []}
        with torch.no_grad():
            for im in sets:
                x = im["data"]["x"].unsqueeze(0).to(device)
                z = t.G.encode(x)
                d_pred, _ = t.G.decode_d(z)
                s_pred = t.G.decode_s(z, None)
                for task, pred in (("d", d_pred), ("s", s_pred)):
                    acc[task].append(o.accuracy(pred.cpu().numpy(), im["data"][task].unsqueeze(0).numpy()))
                    iou[task].append(o.miou(pred.cpu().numpy(), im["data"][task].unsqueeze(0).numpy()))
    for task in ("s", "d"):
        assert got[f"{task}.accuracy"] == pytest.approx(np.mean(acc[task]), rel=1e-12)
        assert got[f"{task}.mIOU"] == pytest.approx(np.mean(iou[task]), rel=1e-12)
    assert got["m.accuracy"] == -1 and got["m.mIOU"] == -1
