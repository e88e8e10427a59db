"""GPU parity of Trainer.infer_all (masker + painter inference, wildfire / smog / flood compositing, uint8 NHWC output edge)
against the reference's own Trainer.infer_all run on CPU (tests/golden/infer_all.*, make_golden.py::run_infer_all_case)."""
import json
import os
import random

import numpy as np
import pytest
import torch

from climategan_b200 import events
from climategan_b200.trainer import Trainer
from climategan_b200.utils import full_opts
from tests.golden.weights import fill_state_dict, synth_inputs
from tests.helpers import GOLDEN, rel_max

pytestmark = pytest.mark.gpu


def _trainer(cuda, dtype):
    meta = json.load(open(os.path.join(GOLDEN, "infer_all.json")))
    g = dict(np.load(os.path.join(GOLDEN, "infer_all.npz")))
    size = meta["size"]
    t = Trainer(full_opts(size=size), device=cuda, storage_dtype=dtype).setup(inference=True, input_shape=(size, size))
    sd = fill_state_dict([(k, tuple(s)) for k, s in meta["g_shapes"]], meta["seeds"]["G"])
    t.G.load_state_dict({k: v.to(cuda) for k, v in sd.items()}, strict=True)
    x = synth_inputs(meta["batch"], size, seed=meta["seeds"]["inputs"])[0]
    return meta, g, t, x


def test_infer_all_fp32_matches_reference(cuda):
    """fp32 storage.  Stated tolerances: float events (numpy=False) within 2e-3 of full scale on a stride-4 grid; uint8 events
    within 1 LSB on >= 99.5 % of the pixels (stride-2 grid); binarised mask mismatch <= 1e-3 of the pixels."""
    meta, g, t, x = _trainer(cuda, torch.float32)
    # same call order as the golden run: the spectral-norm power iteration advances u / v on EVERY forward, eval included
    # (norms.py:100-112), so the first and the second infer_all of a freshly loaded model differ
    random.seed(meta["seeds"]["random"])
    out = t.infer_all(x.permute(0, 2, 3, 1).numpy(), numpy=True, bin_value=0.5, return_masks=True)   # NHWC numpy input
    for k in ("flood", "wildfire", "smog"):
        assert out[k].dtype == np.uint8 and out[k].shape == (meta["batch"], meta["size"], meta["size"], 3)
        diff = np.abs(out[k][:, ::2, ::2].astype(np.int32) - g[k].astype(np.int32))
        assert (diff <= 1).mean() >= 0.995, (k, float((diff <= 1).mean()), int(diff.max()))
    mism = (out["mask"][:, :, ::2, ::2] != g["mask"]).mean()
    assert mism <= 1e-3, float(mism)
    random.seed(meta["seeds"]["random"])
    raw = t.infer_all(x.clone(), numpy=False)
    for k in ("flood", "wildfire", "smog"):
        got = raw[k][:, :, ::4, ::4].cpu()
        assert rel_max(got, torch.from_numpy(g["raw_" + k])) < 2e-3, (k, rel_max(got, torch.from_numpy(g["raw_" + k])))
    # third call, cloudy=True: the flood is painted through an intermediary image whose sky is Perlin noise; the noise's
    # gradient angles come from torch.rand on the CPU generator in both implementations (tutils.py:660)
    torch.manual_seed(0)
    random.seed(meta["seeds"]["random"])
    cl = t.infer_all(x.clone(), numpy=False, cloudy=True)
    got = cl["flood"][:, :, ::4, ::4].cpu()
    assert rel_max(got, torch.from_numpy(g["raw_flood_cloudy"])) < 3e-3, rel_max(got, torch.from_numpy(g["raw_flood_cloudy"]))
    assert rel_max(got, torch.from_numpy(g["raw_flood"])) > 1e-2   # and it does differ from the non-cloudy flood


def test_infer_all_bf16_close_to_reference(cuda):
    """bf16 storage: every event within 8 LSB on >= 97 % of the pixels and 3 LSB on average (the painter output is 2.5e-2 of full
    scale = 6 LSB off in bf16; smog divides by the per-sample depth range, which amplifies the bf16 rounding of d); mask mismatch
    <= 2 % (random-weight logits sit near 0)."""
    meta, g, t, x = _trainer(cuda, torch.bfloat16)
    random.seed(meta["seeds"]["random"])
    out = t.infer_all(x.clone(), numpy=True, bin_value=0.5, return_masks=True)
    for k in ("flood", "wildfire", "smog"):
        diff = np.abs(out[k][:, ::2, ::2].astype(np.int32) - g[k].astype(np.int32))
        assert (diff <= 8).mean() >= 0.97 and diff.mean() <= 3, (k, float((diff <= 8).mean()), float(diff.mean()), int(diff.max()))
    assert (out["mask"][:, :, ::2, ::2] != g["mask"]).mean() <= 2e-2


def test_infer_all_single_image_and_ignore(cuda):
    meta, g, t, x = _trainer(cuda, torch.bfloat16)
    out = t.infer_all(x[0], numpy=True, ignore_event={"smog"})
    assert out["smog"] is None and out["flood"].shape == (1, meta["size"], meta["size"], 3)


def _gauss2d(k, s):
    ax = torch.arange(k, dtype=torch.float32) - k // 2
    g1 = torch.exp(-ax ** 2 / (2 * s * s))
    g1 = g1 / g1.sum()
    return g1[:, None] * g1[None, :]


def test_event_kernels_against_torch(cuda):
    """The compositing kernels one by one against plain PyTorch restatements of the reference lines (small sizes)."""
    import ctypes as C

    import torch.nn.functional as F

    from climategan_b200 import _lib
    from climategan_b200._lib import check

    torch.manual_seed(0)
    n, h, w = 2, 40, 52
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L = _lib.lib()
    mask = (torch.rand(n, h, w) > 0.97).float()
    # increase_sky_mask: shifted sums then clamp (fire.py:15-47)
    def ref_inc(m, p_w, p_h):
        m = m[:, None]
        n_lines, n_cols = int(p_h * h), int(p_w * w)
        tmp = m.clone()
        for i in range(1, n_cols):
            tmp[:, :, :, i:] += m[:, :, :, :-i]
            tmp[:, :, :, :-i] += m[:, :, :, i:]
        new = tmp.clone()
        for i in range(1, n_lines):
            new[:, :, i:, :] += tmp[:, :, :-i, :]
            new[:, :, :-i, :] += tmp[:, :, i:, :]
        new[new >= 1] = 1
        return new[:, 0]
    md = mask.to(cuda)
    tmp, out = torch.empty_like(md), torch.empty_like(md)
    check(L.cgb_box_dilate(P(md), P(tmp), P(out), n, h, w, int(0.18 * w) - 1, int(0.18 * h) - 1, st))
    assert torch.equal(out.cpu(), ref_inc(mask, 0.18, 0.18))
    # gaussian blur: dense 2-D kernel with reflect padding (kornia filter2d restated)
    k, s = 21, 10.5
    src = torch.rand(n, h, w)
    ref = F.conv2d(F.pad(src[:, None], (k // 2,) * 4, mode="reflect"), _gauss2d(k, s)[None, None])[:, 0]
    sd = src.to(cuda)
    check(L.cgb_gauss_blur(P(sd), P(tmp), P(out), n, h, w, k, s, st))
    assert rel_max(out, ref) < 1e-5
    # normalize -> uint8 NHWC
    img = torch.randn(n, 3, h, w)
    mn = img.reshape(n, -1).min(1)[0].reshape(n, 1, 1, 1)
    tt = img - mn
    tt = tt / tt.reshape(n, -1).max(1)[0].reshape(n, 1, 1, 1)
    ref8 = (tt.permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
    got8 = events.to_uint8_nhwc(img.to(cuda)).cpu().numpy()
    assert (np.abs(got8.astype(int) - ref8.astype(int)) <= 1).all() and (got8 == ref8).mean() > 0.999


def test_save_resume_and_apply_events_cli(cuda, tmp_path):
    """Trainer.save -> Trainer.resume_from_path (reference checkpoint layout: opts.yaml + checkpoints/latest_ckpt.pth with keys
    G / g_opt / D / d_opt / epoch / step) and the apply_events.py CLI: events written as {stem}_{event}_{width}{suffix}.png."""
    import subprocess
    import sys

    import yaml
    from PIL import Image

    size = 256
    opts = full_opts(size=size)
    opts.output_path = str(tmp_path / "run")
    t = Trainer(opts, device=cuda, storage_dtype=torch.bfloat16).setup(input_shape=(size, size))
    (tmp_path / "run").mkdir()
    with open(tmp_path / "run" / "opts.yaml", "w") as f:
        yaml.safe_dump(opts.to_dict(), f)
    ckpt = t.save()
    saved = torch.load(ckpt, map_location="cpu")
    assert {"G", "g_opt", "D", "d_opt", "epoch", "step"} <= set(saved)
    t2 = Trainer.resume_from_path(tmp_path / "run", inference=True, device=cuda, input_shape=(size, size))
    for (k, a), (_, b) in zip(t.G.state_dict().items(), t2.G.state_dict().items()):
        assert torch.equal(a, b), k
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    rs = np.random.RandomState(0)
    for i, (h, w) in enumerate([(300, 400), (512, 512), (260, 700)]):
        Image.fromarray(rs.randint(0, 255, size=(h, w, 3), dtype=np.uint8)).save(imgs / f"im{i}.png")
    out = tmp_path / "out"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "apply_events.py"), "-i", str(imgs), "-o", str(out), "-r",
                          str(tmp_path / "run"), "-t", str(size), "-b", "2", "--save_masks", "--no_cloudy", "--fuse"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    for i in range(3):
        for ev in ("flood", "wildfire", "smog", "mask"):
            f = out / f"im{i}_{ev}_{size}_no_cloudy.png"
            assert f.exists(), (f, sorted(p.name for p in out.iterdir()))
            im = np.asarray(Image.open(f))
            assert im.shape[:2] == (size, size)
