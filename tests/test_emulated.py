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
GPU-only; the events kernels of infer_all are not emulated."""
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


def _id(run):
    fn, kw = run
    tail = "-".join(str(v).replace("torch.", "") for v in kw.values())
    return fn.__name__[5:] + ("-" + tail if tail else "")


@pytest.mark.parametrize("run", RUNS, ids=_id)
def test_gpu_parity_test_on_the_emulated_abi(run):
    fn, kw = run
    with emulated_library() as lib:
        fn(torch.device("cpu"), **kw)
        assert sum(lib.calls.values()) > 0
