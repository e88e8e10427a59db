"""The option sweep of tests/test_emulated.py (12 option combinations around the reference's scenario matrix, pinned to two
iterations of the reference's own Trainer: tests/golden/config_sweep.*) on the GPU, fp32 storage.  These combinations compose
kernels that are each pinned by the rest of the GPU suite.  Green on B200 since the first GPU call of round 2
(profiles/r02_pytest_gpu.log), so it runs with the rest of the suite.
"""
import pytest

from tests.test_emulated import run_sweep_case

pytestmark = pytest.mark.gpu

CASES = ["dada_ms", "base_depth_regression", "v3_spade_msdp", "spade_detached_cond", "adam", "pseudo_labels", "minent_v1_no_gi",
         "depth_and_seg_only", "dada_depth_loss", "painter_local_d", "painter_local_d_pl4m", "painter_aux_losses"]


@pytest.mark.parametrize("case", CASES)
def test_option_sweep_fp32_matches_reference_trainer(cuda, case):
    run_sweep_case(case, cuda, g_norm_rtol=2e-2)
