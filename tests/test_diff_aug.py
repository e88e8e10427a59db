"""gen.p.diff_aug (DiffTransforms, transforms.py:493-626) on CPU: the product's host path (same generator
consumption as the reference, packed draws, autograd Function) through the emulated C ABI, and the plain-PyTorch statement the GPU
suite compares the kernels with, both against the reference's own DiffTransforms under the same seed — value and gradient."""
import pytest
import torch

from climategan_b200.utils import Dict
from tests.helpers import rel_max, torch_diff_aug

COMBOS = [dict(do_color_jittering=True, do_cutout=False, do_translation=False),
          dict(do_color_jittering=False, do_cutout=False, do_translation=True),
          dict(do_color_jittering=False, do_cutout=True, do_translation=False),
          dict(do_color_jittering=True, do_cutout=True, do_translation=True),
          dict(do_color_jittering=False, do_cutout=False, do_translation=False)]


def _opts(combo, cutout_ratio=0.5, translation_ratio=0.125):
    return Dict(use=True, cutout_ratio=cutout_ratio, translation_ratio=translation_ratio, **combo)


@pytest.mark.parametrize("combo", COMBOS, ids=lambda c: "".join(k[3] for k, v in c.items() if v) or "none")
@pytest.mark.parametrize("shape,ratios", [((3, 3, 24, 40), (0.5, 0.125)), ((2, 3, 17, 23), (0.3, 0.3)), ((4, 4, 16, 16), (0.9, 0.5))])
def test_diff_transforms_match_the_reference(combo, shape, ratios):
    from oracle import refshim

    if not refshim.available():
        pytest.skip("reference tree absent")
    ref_t = refshim.load("transforms")
    from climategan_b200.transforms import DiffTransforms
    from tests.emulib import emulated_library

    o = _opts(combo, *ratios)
    for seed in range(4):
        g = torch.Generator().manual_seed(100 + seed)
        x = torch.randn(*shape, generator=g)
        wgt = torch.randn(*shape, generator=g)
        xr = x.clone().requires_grad_()
        torch.manual_seed(seed)
        want = ref_t.DiffTransforms(o)(xr)
        (want * wgt).sum().backward()
        end_state = torch.get_rng_state()
        with emulated_library() as lib:
            xo = x.clone().requires_grad_()
            torch.manual_seed(seed)
            t = DiffTransforms(o)
            got = t(xo)
            assert torch.equal(torch.get_rng_state(), end_state), "generator consumed differently from the reference"
            (got * wgt).sum().backward()
            assert lib.calls["cgb_diff_aug_fwd"] == 1 and lib.calls["cgb_diff_aug_bwd"] == 1
            torch.manual_seed(seed)
            params, cut_h, cut_w = t.draw(x)
        assert rel_max(got, want) < 2e-6 and rel_max(xo.grad, xr.grad) < 2e-6
        xh = x.clone().requires_grad_()
        plain = torch_diff_aug(xh, params, cut_h, cut_w)
        (plain * wgt).sum().backward()
        assert rel_max(plain, want) < 2e-6 and rel_max(xh.grad, xr.grad) < 2e-6


def test_painter_step_with_diff_aug_runs_on_the_emulated_abi():
    """Trainer.update_G / update_D with gen.p.diff_aug on (tasks = [p]): the augmentation sits in front of the painter
    discriminator in both steps (trainer.py:1079-1081, 1319-1321), two independent draws per step side; finite losses, a gradient
    on every painter weight."""
    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts, synth_batch
    from tests.emulib import emulated_library

    size = 64
    opts = full_opts(size=size, tasks=("p",), overrides={f"gen.p.diff_aug.{k}": v for k, v in _opts(COMBOS[3]).items()})
    with emulated_library() as lib:
        torch.manual_seed(0)
        t = Trainer(opts, device=torch.device("cpu"), storage_dtype=torch.float32).setup(input_shape=(size, size))
        mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, 2, size, 3).items()}
        t.update_G(mdb)
        assert lib.calls["cgb_diff_aug_fwd"] == 2 and lib.calls["cgb_diff_aug_bwd"] == 1   # x carries no gradient
        t.update_D(mdb)
        assert lib.calls["cgb_diff_aug_fwd"] == 4
        losses = t.losses_to_host()
    assert all(torch.isfinite(torch.tensor(float(v))) for v in _leaves(losses))


def _leaves(d):
    for v in d.values():
        if isinstance(v, dict):
            yield from _leaves(v)
        elif v is not None:
            yield v
