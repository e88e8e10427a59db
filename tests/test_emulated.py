"""CPU parity of the product's Python layer for the painter path: OmniGenerator.paint + L1 + backward with the C ABI
emulated in plain PyTorch (tests/emulib.py, one naive function per entry point, following include/cgb200.h), against the
golden the unmodified reference produced (tests/golden/painter_small.*).  What this pins without a GPU: the weight packing
and the fused gamma||beta packing, the im2col form of mlp_shared, the shared instance-norm statistics of norm_0 / norm_s, the
residual-in-epilogue wiring, and the backward formulas of every autograd Function on the path (SPADE, conv, spectral norm with
its in-place power iteration, nearest resize, paste, layout edges).  The kernels themselves are pinned on the GPU
(tests/test_gpu_*.py)."""
import pytest
import torch

from climategan_b200 import ops
from climategan_b200.generator import OmniGenerator
from climategan_b200.utils import default_painter_opts
from tests.emulib import emulated_library
from tests.helpers import cosine, load_golden, rel_l2, rel_max


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_paint_forward_backward_on_the_emulated_abi_matches_the_reference_golden(dtype):
    meta, g, sd, (x, m, t) = load_golden()
    fp32 = dtype == torch.float32
    with emulated_library() as lib:
        opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"])
        G = OmniGenerator(opts, latent_shape=meta["size"], storage_dtype=dtype)
        G.painter.load_state_dict(sd, strict=True)
        G.train()
        out = G.paint(m, x)
        loss = ops.l1_loss(out, t)
        loss.backward()
        assert rel_max(out, torch.from_numpy(g["out"])) < (1e-5 if fp32 else 5e-2)
        assert abs(float(loss.detach()) - float(g["loss"])) / float(g["loss"]) < (1e-6 if fp32 else 1e-2)
        params = dict(G.painter.named_parameters())
        # spectral-norm u / v advanced one power iteration, in place (norms.py:106-108)
        assert rel_max(params["head_0.conv_0.module.weight_u"], torch.from_numpy(g["u_after"])) < 1e-5
        assert rel_max(params["up_spades.0.conv_s.module.weight_v"], torch.from_numpy(g["v_after"])) < 1e-5
        for k, v in g.items():
            if not k.startswith("grad::"):
                continue
            gm, gr = params[k[6:]].grad, torch.from_numpy(v)
            if fp32:
                assert rel_max(gm, gr) < 1e-5, (k, rel_max(gm, gr))
            else:
                assert cosine(gm, gr) > 0.98 and rel_l2(gm, gr) < 0.25, (k, cosine(gm, gr), rel_l2(gm, gr))
        norms_ref = dict(zip(meta["grad_keys"], g["grad_norms"]))
        bad = [(k, float(params[k].grad.norm()), r) for k, r in norms_ref.items()
               if r >= 1e-6 and abs(float(params[k].grad.norm()) - r) / r > (1e-4 if fp32 else 0.1)]
        assert not bad, bad[:5]
        with torch.no_grad():   # second forward: the advanced u / v
            out2 = G.paint(m, x)
        assert rel_max(out2, torch.from_numpy(g["out_second_forward"])) < (1e-5 if fp32 else 5e-2)
        assert lib.calls["cgb_conv2d_fwd"] > 40 and lib.calls["cgb_spectral_power_iter"] > 20


def test_painter_train_steps_on_the_emulated_abi_match_the_reference():
    """Trainer.update_G / update_D x2 on the painter task (VGG + GAN + feature-matching losses, multi-scale discriminator,
    ExtraAdam extrapolation then step) against the reference modules' own four optimiser steps
    (tests/golden/painter_step.*): losses to 1e-4, parameters as in tests/test_gpu_trainer.py."""
    import json
    import os

    import numpy as np

    from climategan_b200.trainer import Trainer
    from tests.golden.weights import fill_state_dict, synth_inputs
    from tests.helpers import GOLDEN

    meta = json.load(open(os.path.join(GOLDEN, "painter_step.json")))
    g = dict(np.load(os.path.join(GOLDEN, "painter_step.npz")))
    opts = default_painter_opts(latent_dim=meta["latent_dim"], spade_n_up=meta["spade_n_up"], ndf=meta["ndf"],
                                n_layers=meta["n_layers"], num_D=meta["num_D"])
    opts.dis.soft_shift = 0.0
    opts.dis.flip_prob = 0.0
    with emulated_library():
        t = Trainer(opts, device=torch.device("cpu"), storage_dtype=torch.float32).setup(input_shape=(meta["size"], meta["size"]))
        mk = lambda shapes, seed: fill_state_dict([(k, tuple(s)) for k, s in shapes], seed)  # noqa: E731
        t.G.painter.load_state_dict(mk(meta["g_shapes"], 11), strict=True)
        t.D.load_state_dict(mk(meta["d_shapes"], 12), strict=True)
        t.losses["G"]["p"]["vgg"].vgg.load_state_dict(mk(meta["v_shapes"], 13), strict=True)
        x, m, _ = synth_inputs(meta["batch"], meta["size"], 5)
        batch = {"rf": t.batch_to_device({"data": {"x": x, "m": m}, "domain": "rf", "mode": "train"})}
        logs = []
        for _ in range(2):
            t.update_G(batch)
            L = t.losses_to_host()
            logs += [L["gen"]["p"]["vgg"], L["gen"]["p"]["gan"], L["gen"]["p"]["featmatch"]]
            t.update_D(batch)
            logs.append(t.losses_to_host()["disc"]["p"]["gan"])
            t.logger.global_step += 1
        np.testing.assert_allclose(np.array(logs), g["logs"], rtol=1e-4)
        gsd, dsd = t.G.painter.state_dict(), t.D.state_dict()
        for k, v in g.items():
            if k.startswith("G::"):
                mine, lr = gsd[k[3:]], 5e-5
            elif k.startswith("D::"):
                mine, lr = dsd[k[3:]], 2e-5
            else:
                continue
            delta = (mine - torch.from_numpy(v)).abs()
            assert float(delta.max()) <= 2.2 * lr, (k, float(delta.max()))   # (see tests/test_gpu_trainer.py for the criterion)
            if k != "G::fc.bias":
                assert float(delta.mean()) <= 0.05 * lr, (k, float(delta.mean()))
