"""The option sweep of tests/test_emulated.py (12 option combinations around the reference's scenario matrix, pinned to two
iterations of the reference's own Trainer: tests/golden/config_sweep.*) on the GPU, fp32 storage.  These combinations compose
kernels that are each pinned by the rest of the GPU suite; the sweep itself was written after round 1's GPU budget was spent and
has only run on the CPU emulation so far, so it is opt-in until it has been seen green on hardware:

    CGB_RUN_SWEEP=1 python -m pytest tests/test_gpu_sweep.py -q -m gpu
"""
import os

import pytest

from tests.test_emulated import run_sweep_case

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CGB_RUN_SWEEP") != "1", reason="opt-in until first run on hardware (CGB_RUN_SWEEP=1)")]

CASES = ["dada_ms", "base_depth_regression", "v3_spade_msdp", "spade_detached_cond", "adam", "pseudo_labels", "minent_v1_no_gi",
         "depth_and_seg_only", "dada_depth_loss", "painter_local_d", "painter_local_d_pl4m", "painter_aux_losses"]


@pytest.mark.parametrize("case", CASES)
def test_option_sweep_fp32_matches_reference_trainer(cuda, case):
    run_sweep_case(case, cuda, g_norm_rtol=2e-2)
