"""CPU parity of the product's Python layer.  The parity tests of the GPU suite (tests/test_gpu_*.py) are run HERE, unchanged
— same fixtures, same assertions, same tolerances — on CPU tensors, with the C ABI emulated in plain PyTorch
(tests/emulib.py: one naive function per entry point, following the contracts in include/cgb200.h).

What this pins without a GPU, against goldens produced by the unmodified reference: module wiring, weight packing and the
fused gamma||beta packing, the im2col forms, shared statistics, residual-in-epilogue wiring, every autograd Function's backward
formula (SPADE with instance / batch statistics, conv, spectral norm with its in-place power iteration and the u / v gradients
of the discriminators, BatchNorm, resizes, the differentiable make_m_cond, paste / mask adjoints), every loss assembly of
Trainer.get_masker_loss / get_painter_loss / get_D_loss, the flat ExtraAdam — for the v2 and v3 maskers, the SPADE mask decoder,
pl4m, the base depth decoder with classification, the painter options.  What it cannot pin is the kernels themselves: that is
the GPU suite's job (the emulation replaces them).  The bf16 train-step fixtures are chaotic (see their docstrings) and stay
GPU-only."""
import pytest
import torch

import tests.test_gpu_discriminator as t_disc
import tests.test_gpu_full_step as t_step
import tests.test_gpu_masker as t_masker
import tests.test_gpu_masker_v3 as t_v3
import tests.test_gpu_painter as t_painter
import tests.test_gpu_trainer as t_trainer
from tests.emulib import emulated_library

F32, BF16 = torch.float32, torch.bfloat16
RUNS = [
    # painter: paint + L1 + backward, explicit latent + final shortcut, painter train step (D, VGG, GAN / featmatch, ExtraAdam)
    (t_painter.test_paint_matches_reference_golden, dict(dtype=F32)),
    (t_painter.test_paint_matches_reference_golden, dict(dtype=BF16)),
    (t_painter.test_no_paste_and_painter_forward, {}),
    (t_painter.test_painter_explicit_z_and_final_shortcut_match_reference_golden, dict(dtype=F32)),
    (t_painter.test_painter_explicit_z_and_final_shortcut_match_reference_golden, dict(dtype=BF16)),
    (t_trainer.test_train_steps_fp32_match_reference, {}),
    (t_trainer.test_train_steps_bf16_close_to_reference, {}),
    # discriminators
    (t_disc.test_discriminator_matches_reference_golden, dict(dtype=F32)),
    (t_disc.test_discriminator_matches_reference_golden, dict(dtype=BF16)),
    (t_disc.test_fc_discriminator, {}),
    (t_disc.test_avgpool_and_instnorm_act, {}),
    # maskers: eval decodes (v2, v2 + SPADE decoder with 15 / 12 conditioning channels, v3 + SPADE) and train-mode functionals
    (t_masker.test_masker_decode_matches_reference_golden, dict(dtype=F32)),
    (t_masker.test_masker_decode_matches_reference_golden, dict(dtype=BF16)),
    (t_masker.test_masker_train_mode_uses_batch_statistics, {}),
    (t_masker.test_masker_spade_decoder_matches_reference_golden, dict(dtype=F32, case="masker_spade")),
    (t_masker.test_masker_spade_decoder_matches_reference_golden, dict(dtype=F32, case="masker_spade12")),
    (t_v3.test_masker_v3_matches_reference_golden, dict(dtype=F32, case="masker_v3_spade")),
    (t_v3.test_masker_v3_matches_reference_golden, dict(dtype=BF16, case="masker_v3")),
    (t_v3.test_full_train_step_runs_with_the_v3_masker, {}),
    # two iterations of Trainer.update_G / update_D against the reference's own Trainer, every configuration
    (t_step.test_full_step_fp32_matches_reference_trainer, {}),
    (t_step.test_full_step_with_pl4m_fp32_matches_reference_trainer, {}),
    (t_step.test_spade_masker_step_fp32_matches_reference_trainer, {}),
    (t_step.test_base_depth_classify_step_fp32_matches_reference_trainer, {}),
    (t_step.test_base_depth_classify_step_bf16_runs_close, {}),
    (t_step.test_v3_masker_step_fp32_matches_reference_trainer, {}),
    (t_step.test_v3_mask_only_step_fp32_matches_reference_trainer, {}),
    (t_step.test_v3_masker_step_bf16_runs_close, {}),
]


import tests.test_gpu_infer_all as t_infer  # noqa: E402

import tests.test_gpu_zz_new_kernels as t_new  # noqa: E402

RUNS += [
    (t_infer.test_infer_all_bf16_close_to_reference, {}),
    (t_infer.test_infer_all_single_image_and_ignore, {}),
    # the entry points that have not run on hardware yet: the tests' own logic and tolerances, against the header contracts
    (t_new.test_dada_depth_loss, dict(n=2, h=16, w=12)),
    (t_new.test_dada_depth_loss, dict(n=3, h=33, w=47)),
    (t_new.test_pack_weight_kernel, dict(dtype=F32)),
    (t_new.test_pack_weight_kernel, dict(dtype=BF16)),
    (t_new.test_eval_metrics_match_reference_fixture_and_oracle, {}),
    (t_new.test_trainer_eval_images, {}),
    (t_new.test_diff_aug_kernels, dict(shape=(3, 3, 24, 40), cut=(12, 20), shift=(3, 5))),
    (t_new.test_diff_aug_kernels, dict(shape=(2, 3, 17, 23), cut=(5, 7), shift=(5, 7))),
    (t_new.test_diff_aug_kernels, dict(shape=(2, 4, 16, 16), cut=(0, 0), shift=(8, 8))),
]


def _id(run):
    fn, kw = run
    tail = "-".join(str(v).replace("torch.", "") for v in kw.values())
    return fn.__name__[5:] + ("-" + tail if tail else "")


@pytest.mark.parametrize("run", RUNS, ids=_id)
def test_gpu_parity_test_on_the_emulated_abi(run):
    fn, kw = run
    t_step.NORM_TOL_SCALE = 2.0   # (see tests/test_gpu_full_step.py)
    try:
        with emulated_library() as lib:
            fn(torch.device("cpu"), **kw)
            assert sum(lib.calls.values()) > 0
    finally:
        t_step.NORM_TOL_SCALE = 1.0


def test_infer_all_on_the_emulated_abi_matches_the_reference():
    """Trainer.infer_all (masker + painter + flood / wildfire / smog compositing + the uint8 edge, and the cloudy flood) against
    the reference's own Trainer.infer_all (tests/golden/infer_all.*), as tests/test_gpu_infer_all.py does on the GPU.  The float
    flood and smog match to 1e-5; the wildfire image is uint8-valued (torchvision's uint8 blends) and one sampled pixel in
    38 400 sits on a truncation boundary of this emulation's blur, so it is held to <= 1 LSB with >= 99.9 % exact."""
    import random

    import numpy as np

    from tests.helpers import rel_max

    with emulated_library():
        meta, g, t, x = t_infer._trainer(torch.device("cpu"), torch.float32)
        random.seed(meta["seeds"]["random"])
        out = t.infer_all(x.permute(0, 2, 3, 1).numpy(), numpy=True, bin_value=0.5, return_masks=True)
        for k in ("flood", "wildfire", "smog"):
            assert out[k].dtype == np.uint8 and out[k].shape == (meta["batch"], meta["size"], meta["size"], 3)
            diff = np.abs(out[k][:, ::2, ::2].astype(np.int32) - g[k].astype(np.int32))
            assert (diff <= 1).mean() >= 0.995, (k, float((diff <= 1).mean()), int(diff.max()))
        assert (out["mask"][:, :, ::2, ::2] != g["mask"]).mean() <= 1e-3
        random.seed(meta["seeds"]["random"])
        raw = t.infer_all(x.clone(), numpy=False)
        for k in ("flood", "smog"):
            assert rel_max(raw[k][:, :, ::4, ::4], torch.from_numpy(g["raw_" + k])) < 1e-5, k
        d = (raw["wildfire"][:, :, ::4, ::4] - torch.from_numpy(g["raw_wildfire"])).abs()
        assert float(d.max()) <= 1.0 and float((d == 0).float().mean()) >= 0.999
        torch.manual_seed(0)
        random.seed(meta["seeds"]["random"])
        cl = t.infer_all(x.clone(), numpy=False, cloudy=True)
        got = cl["flood"][:, :, ::4, ::4]
        assert rel_max(got, torch.from_numpy(g["raw_flood_cloudy"])) < 3e-3
        assert rel_max(got, torch.from_numpy(g["raw_flood"])) > 1e-2


def _sweep():
    import json
    import os

    from tests.helpers import GOLDEN

    return json.load(open(os.path.join(GOLDEN, "config_sweep.json")))


@pytest.mark.parametrize("case", ["dada_ms", "base_depth_regression", "v3_spade_msdp", "spade_detached_cond", "adam", "pseudo_labels",
                                  "minent_v1_no_gi", "depth_and_seg_only", "dada_depth_loss", "painter_local_d", "painter_local_d_pl4m",
                                  "painter_aux_losses"])
def test_option_sweep_on_the_emulated_abi_matches_the_reference_trainer(case):
    with emulated_library():
        run_sweep_case(case, torch.device("cpu"))


def test_painter_step_with_diff_aug_on_the_emulated_abi_matches_the_reference_trainer():
    """gen.p.diff_aug (jitter + translation + cutout) in front of the painter discriminator: two iterations of update_G / update_D
    against the reference's own Trainer under the same generator seeds (tests/golden/config_sweep_diffaug.*), same tolerances as
    the option sweep."""
    with emulated_library() as lib:
        run_sweep_case("painter_diff_aug", torch.device("cpu"), fixture="config_sweep_diffaug", reseed=True)
        assert lib.calls["cgb_diff_aug_fwd"] == 8 and lib.calls["cgb_diff_aug_bwd"] == 2


def run_sweep_case(case, device, g_norm_rtol=1e-2, fixture="config_sweep", reseed=False):
    """Option combinations around the reference's scenario matrix that have no full fixture (DADA on the mask decoder, base depth
    regression, the reverse-Huber depth loss, v3 encoder + SPADE mask decoder + painter, detached SPADE conditioning, plain Adam, pseudo labels on the real
    domain, MinEnt v1 without the ground-intersection loss, tasks d + s alone, the global + local painter discriminators with
    and without the painter loss for the masker, the painter's tv / context / reconstruction losses): two iterations of update_G / update_D against the
    reference's own Trainer (tests/golden/config_sweep.*, from make_golden.py::run_config_sweep) — every logged loss of the first
    iteration within 1e-4, of the second within 3e-3, every gradient norm of the first backward within 1e-2 (G) / 1.5e-1 (D)."""
    import os

    import numpy as np

    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts, synth_batch
    from tests.golden.weights import fill_state_dict
    from tests.helpers import GOLDEN

    import json

    sweep = json.load(open(os.path.join(GOLDEN, fixture + ".json")))
    meta = sweep["cases"][case]
    arrays = np.load(os.path.join(GOLDEN, fixture + ".npz"))
    size, batch = sweep["size"], sweep["batch"]
    kw = dict(meta["kw"])
    kw["tasks"] = tuple(kw["tasks"])
    opts = full_opts(size=size, **kw)
    if True:
        t = Trainer(opts, device=device, storage_dtype=torch.float32).setup(input_shape=(size, size))
        mk = lambda shapes, seed: {k: v.to(device) for k, v in fill_state_dict([(k, tuple(s)) for k, s in shapes], seed).items()}  # noqa: E731
        t.G.load_state_dict(mk(meta["g_shapes"], sweep["seeds"]["G"]), strict=True)
        t.D.load_state_dict(mk(meta["d_shapes"], sweep["seeds"]["D"]), strict=True)
        if meta["v_shapes"]:
            t.losses["G"]["p"]["vgg"].vgg.load_state_dict(mk(meta["v_shapes"], sweep["seeds"]["vgg"]), strict=True)
        t.use_pl4m = bool(meta.get("pl4m", False))
        for mod in t.G.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, batch, size, sweep["seeds"]["inputs"]).items()}
        dn = None
        for it in range(2):
            if reseed:   # the fixture seeds torch's generator before each step side (make_golden.py::SWEEP_DIFFAUG)
                torch.manual_seed(1000 + it)
            t.update_G(mdb)
            if it == 0:
                gn = np.array([float(p.grad.norm()) if p.grad is not None and p.requires_grad else -1.0 for p in t.G.parameters()])
            if reseed:
                torch.manual_seed(2000 + it)
            t.update_D(mdb)
            if it == 0 and t.d_opt is not None:
                dn = np.array([float(p.grad.norm()) if p.grad is not None and p.requires_grad else -1.0 for p in t.D.parameters()])
            t.logger.global_step += 1
            logs = t_step._flatten(t.losses_to_host())
            tol, atol = (1e-4, 2e-6) if it == 0 else (3e-3, 2e-4)
            bad = [(k, logs.get(k), r) for k, r in meta["logs"][it].items() if k not in logs or abs(logs[k] - r) > tol * abs(r) + atol]
            assert not bad, (it, bad[:6])
        for got, key, rtol in ((gn, case + "::G.gradnorm", g_norm_rtol), (dn, case + "::D.gradnorm", 1.5e-1)):
            if got is None:
                continue
            ref = arrays[key]
            scale = ref[ref >= 0].max()
            names = [n for n, _ in (t.G if key.endswith("G.gradnorm") else t.D).named_parameters()]
            bad = [(n, a, b) for n, a, b in zip(names, got, ref)
                   if b >= 0 and not n.endswith(("weight_u", "weight_v")) and (a < 0 or abs(a - b) > rtol * b + 1e-6 * scale)]
            assert not bad, bad[:6]
