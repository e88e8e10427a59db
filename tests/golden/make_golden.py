"""Generate tests/golden/painter_small.{npz,json} by running the UNMODIFIED reference modules
(imported from /root/reference through oracle/refshim.py) on CPU.  Run in the build container:

    python tests/golden/make_golden.py

The reference has no golden vectors of its own; these pin the oracle
(oracle/painter_oracle.py) and, through it and directly, the CUDA path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402
from tests.golden.weights import fill_state_dict, synth_inputs  # noqa: E402
from climategan_b200.utils import default_painter_opts  # noqa: E402

# reference test scenarios 2 + 12 (tests/test_trainer.py:208-260): base depth decoder classifying bucketised log-depth, no DADA
# fusion, nearest-x2 in front of the segmentation head; 16 buckets keep the fixture small
BASE_DEPTH_CLASSIFY = {"gen.d.architecture": "base", "gen.d.classify.enable": True, "gen.d.classify.linspace.buckets": 16,
                       "gen.m.use_dada": False, "gen.s.use_dada": False, "gen.s.upsample_featuremaps": True}

# the reference-DEFAULT masker (deeplabv3 encoder / decoder, mask decoder with low-level features; defaults.yaml:101,136)
V3_MASKER = {"gen.encoder.architecture": "deeplabv3", "gen.s.architecture": "deeplabv3", "gen.deeplabv3.nblocks": [2, 2, 3, 2]}

CASES = {
    # name: (latent_dim, spade_n_up, batch, size)   channels 40->40->40->20 exercise the %8 padding
    "painter_small": (40, 3, 2, 32),
}


def run_case(name, latent_dim, n_up, batch, size):
    painter_mod, generator_mod = refshim.load("painter", "generator")
    opts = default_painter_opts(latent_dim=latent_dim, spade_n_up=n_up)
    torch.manual_seed(0)
    G = generator_mod.OmniGenerator(opts)  # tasks = ['p'] -> painter only (generator.py:65-101)
    G.painter.set_latent_shape(size, True)
    shapes = [(k, tuple(v.shape)) for k, v in G.painter.state_dict().items()]
    sd = fill_state_dict(shapes, seed=1234)
    G.painter.load_state_dict(sd, strict=True)
    G.train()
    x, m, target = synth_inputs(batch, size, seed=99)
    out = G.paint(m, x)  # generator.py:279-297
    loss = torch.nn.L1Loss()(out, target)
    loss.backward()
    sd_after = {k: v.clone() for k, v in G.painter.state_dict().items()}  # after exactly one forward
    fake = None
    with torch.no_grad():
        # second forward: spectral-norm u/v have advanced one step (norms.py:106-108)
        out2 = G.paint(m, x)
        G2 = generator_mod.OmniGenerator(opts)
        G2.painter.set_latent_shape(size, True)
        G2.painter.load_state_dict(sd, strict=True)
        fake = G2.paint(m, x, no_paste=True)
    grads = {k: p.grad for k, p in G.painter.named_parameters() if p.grad is not None}
    full = ["fc.weight", "conv_img.weight", "conv_img.bias", "final_spade.norm_1.mlp_gamma.weight",
            "final_spade.norm_0.mlp_shared.0.weight", "up_spades.0.conv_s.module.weight_bar",
            "head_0.conv_0.module.weight_bar", "up_spades.0.norm_s.mlp_beta.bias"]
    arrays = {
        "out": out.detach().numpy(),
        "out_second_forward": out2.numpy(),
        "fake_no_paste": fake.numpy(),
        "loss": np.float32(loss.item()),
        "grad_norms": np.array([float(grads[k].norm()) for k, _ in shapes if k in grads], dtype=np.float64),
        "u_after": sd_after["head_0.conv_0.module.weight_u"].numpy(),
        "v_after": sd_after["up_spades.0.conv_s.module.weight_v"].numpy(),
    }
    for k in full:
        arrays["grad::" + k] = grads[k].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {
        "case": name, "latent_dim": latent_dim, "spade_n_up": n_up, "batch": batch, "size": size,
        "weight_seed": 1234, "input_seed": 99,
        "shapes": [[k, list(s)] for k, s in shapes],
        "grad_keys": [k for k, _ in shapes if k in grads],
        "reference": "cc-ai/climategan @ /root/reference (climategan/{painter,generator,blocks,norms}.py)",
        "torch": torch.__version__,
    }
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(name, "loss", float(loss), "out absmax", float(out.abs().max()), "npz bytes",
          os.path.getsize(os.path.join(HERE, name + ".npz")))


def run_painter_z_case(name="painter_z_shortcut", latent_dim=40, n_up=3, batch=2, size=32):
    """The painter's two non-default options together: an explicit latent z (gen.p.no_z = false, generator.py:179-194 /
    painter.py:149-152: fc(cond) is bypassed) and gen.p.use_final_shortcut (painter.py:101-109, 163-164: the last SPADE block is
    conditioned on lrelu(BatchNorm(SN conv1x1(y))), a learned map that receives a gradient through SPADE's mlp_shared).
    ``painter(z, cond)`` is called directly with a seeded z (paint() would draw z from the global RNG)."""
    painter_mod, generator_mod = refshim.load("painter", "generator")
    opts = default_painter_opts(latent_dim=latent_dim, spade_n_up=n_up)
    opts.gen.p.no_z = False
    opts.gen.p.use_final_shortcut = True
    torch.manual_seed(0)
    G = generator_mod.OmniGenerator(opts)
    G.painter.set_latent_shape(size, True)
    shapes = [(k, tuple(v.shape)) for k, v in G.painter.state_dict().items()]
    G.painter.load_state_dict(fill_state_dict(shapes, seed=4321), strict=True)
    G.train()
    x, m, target = synth_inputs(batch, size, seed=98)
    z = torch.randn(batch, latent_dim, G.painter.z_h, G.painter.z_w, generator=torch.Generator().manual_seed(5))
    out = G.painter(z, x * (1.0 - m))
    loss = torch.nn.L1Loss()(out, target)
    loss.backward()
    sd_after = {k: v.clone() for k, v in G.painter.state_dict().items()}
    G.eval()
    with torch.no_grad():
        out_eval = G.painter(z, x * (1.0 - m))
    grads = {k: p.grad for k, p in G.painter.named_parameters() if p.grad is not None}
    full = ["final_shortcut.0.module.weight_bar", "final_shortcut.1.weight", "final_shortcut.1.bias", "conv_img.weight",
            "final_spade.norm_0.mlp_shared.0.weight", "final_spade.norm_1.mlp_gamma.weight", "head_0.conv_0.module.weight_bar",
            "up_spades.0.norm_s.mlp_beta.bias"]
    arrays = {"z": z.numpy(), "out": out.detach().numpy(), "out_eval": out_eval.numpy(), "loss": np.float32(loss.item()),
              "grad_norms": np.array([float(grads[k].norm()) for k, _ in shapes if k in grads], dtype=np.float64),
              "running_mean": sd_after["final_shortcut.1.running_mean"].numpy(),
              "running_var": sd_after["final_shortcut.1.running_var"].numpy()}
    for k in full:
        arrays["grad::" + k] = grads[k].numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "latent_dim": latent_dim, "spade_n_up": n_up, "batch": batch, "size": size, "weight_seed": 4321,
            "input_seed": 98, "shapes": [[k, list(s_)] for k, s_ in shapes], "grad_keys": [k for k, _ in shapes if k in grads],
            "full": full, "reference": "cc-ai/climategan @ /root/reference (climategan/{painter,generator,blocks,norms}.py)",
            "torch": torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(name, "loss", float(loss), "fc grad present:", "fc.weight" in grads, "npz bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def run_disc_case(name="disc_small", ndf=8, n_layers=3, num_d=2, batch=2, size=64):
    """Painter discriminator D["p"] on cat(real, fake) + GANLoss(BCE) + FeatMatchLoss, as Trainer.get_painter_loss
    assembles them (climategan/trainer.py:1362-1383) — generator-side gradients w.r.t. the fake image and D's params."""
    disc_mod, losses_mod, tutils_mod = refshim.load("discriminator", "losses", "tutils")
    from climategan_b200.utils import Dict

    opts = Dict(tasks=["p"], dis=dict(p=dict(ndf=ndf, n_layers=n_layers, norm="instance", use_sigmoid=False, num_D=num_d,
                                             get_intermediate_features=True, use_local_discriminator=False,
                                             init_type="xavier", init_gain=0.02)))
    torch.manual_seed(0)
    D = disc_mod.OmniDiscriminator(opts)
    shapes = [(k, tuple(v.shape)) for k, v in D.state_dict().items()]
    sd = fill_state_dict(shapes, seed=4321)
    D.load_state_dict(sd, strict=True)
    rs = np.random.RandomState(7)
    real = torch.from_numpy((rs.random_sample((batch, 4, size, size)) * 2 - 1).astype(np.float32))
    fake = torch.from_numpy((rs.random_sample((batch, 4, size, size)) * 2 - 1).astype(np.float32)).requires_grad_(True)
    out = D["p"](torch.cat([real, fake], dim=0))
    pred_real, pred_fake = tutils_mod.divide_pred(out)
    g_gan = losses_mod.GANLoss(use_lsgan=False)(pred_fake, True)
    g_feat = losses_mod.FeatMatchLoss()(pred_real, pred_fake)
    d_hinge = losses_mod.HingeLoss()(pred_fake, False, True) + losses_mod.HingeLoss()(pred_real, True, True)
    loss = g_gan + 10.0 * g_feat
    loss.backward()
    grads = {k: p.grad for k, p in D.named_parameters() if p.grad is not None}
    arrays = {
        "real": real.numpy(), "fake": fake.detach().numpy(),
        "g_gan": np.float32(g_gan.item()), "g_feat": np.float32(g_feat.item()), "d_hinge": np.float32(d_hinge.item()),
        "fake_grad": fake.grad.numpy(),
        "grad_norms": np.array([float(grads[k].norm()) for k, _ in shapes if k in grads], dtype=np.float64),
        "grad::p.discriminator_0.model0.0.module.weight_bar": grads["p.discriminator_0.model0.0.module.weight_bar"].numpy(),
        "grad::p.discriminator_1.model2.0.module.weight_bar": grads["p.discriminator_1.model2.0.module.weight_bar"].numpy(),
    }
    for i, feats in enumerate(out):
        arrays[f"pred_{i}"] = feats[-1].detach().numpy()
        arrays[f"feat_{i}_1"] = feats[1].detach().numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "ndf": ndf, "n_layers": n_layers, "num_D": num_d, "batch": batch, "size": size,
            "weight_seed": 4321, "shapes": [[k, list(s)] for k, s in shapes],
            "grad_keys": [k for k, _ in shapes if k in grads],
            "reference": "cc-ai/climategan @ /root/reference (climategan/{discriminator,losses,tutils}.py)",
            "torch": torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(name, "g_gan", float(g_gan), "g_feat", float(g_feat), "d_hinge", float(d_hinge), "npz bytes",
          os.path.getsize(os.path.join(HERE, name + ".npz")))


def run_step_case(name="painter_step", latent=16, n_up=4, ndf=8, n_layers=3, num_d=2, batch=2, size=64):
    """Three consecutive optimiser steps of the painter task composed from the reference's OWN modules exactly as
    Trainer.update_G / update_D do (trainer.py:989-1032, :1256-1387, :1071-1107, :674-694): G step (ExtraAdam
    extrapolation), D step (extrapolation), then G step and D step again (ExtraAdam step).  soft_shift = flip_prob = 0."""
    generator_mod, disc_mod, losses_mod, tutils_mod, optim_mod = refshim.load("generator", "discriminator", "losses", "tutils", "optim")
    import torchvision
    from climategan_b200.losses import Vgg19 as KeyHolder

    # offline + CPU patches of hard-coded assumptions
    _orig_vgg19 = torchvision.models.vgg19
    losses_mod.models.vgg19 = lambda pretrained=True: _orig_vgg19(weights=None)

    def vgg_preprocess_cpu(batch):
        (r, g, b) = torch.chunk(batch, 3, dim=1)
        batch = torch.cat((b, g, r), dim=1)
        batch = (batch + 1) * 255 * 0.5
        mean = torch.zeros_like(batch)
        mean[:, 0], mean[:, 1], mean[:, 2] = 103.939, 116.779, 123.680
        return batch.sub(mean)

    opts = default_painter_opts(latent_dim=latent, spade_n_up=n_up, ndf=ndf, n_layers=n_layers, num_D=num_d)
    torch.manual_seed(0)
    G = generator_mod.OmniGenerator(opts)
    G.painter.set_latent_shape(size, True)
    D = disc_mod.OmniDiscriminator(opts)
    vggloss = losses_mod.VGGLoss("cpu")
    g_shapes = [(k, tuple(v.shape)) for k, v in G.painter.state_dict().items()]
    d_shapes = [(k, tuple(v.shape)) for k, v in D.state_dict().items()]
    v_shapes = [(k, tuple(v.shape)) for k, v in KeyHolder().state_dict().items()]
    G.painter.load_state_dict(fill_state_dict(g_shapes, seed=11), strict=True)
    D.load_state_dict(fill_state_dict(d_shapes, seed=12), strict=True)
    vggloss.vgg.load_state_dict(fill_state_dict(v_shapes, seed=13), strict=True)
    g_opt = optim_mod.ExtraAdam(G.parameters(), lr=opts.gen.opt.lr.default, betas=(opts.gen.opt.beta1, 0.999))
    d_opt = optim_mod.ExtraAdam(D.parameters(), lr=opts.dis.opt.lr.default, betas=(opts.dis.opt.beta1, 0.999))
    gan = losses_mod.GANLoss(use_lsgan=False, soft_shift=0.0, flip_prob=0.0)
    featmatch = losses_mod.FeatMatchLoss()
    x, m, _ = synth_inputs(batch, size, seed=5)
    lam = opts.train.lambdas.G.p
    logs = []
    for it in range(4):
        global_step = it // 2
        if it % 2 == 0:  # update_G
            for p in D.parameters():
                p.requires_grad_(False) if p.requires_grad else None
            tutils_mod.zero_grad(G)
            fake = G.paint(m, x)
            l_vgg = vggloss(vgg_preprocess_cpu(fake * m), vgg_preprocess_cpu(x * m)) * lam.vgg
            real_fake = torch.cat([torch.cat([m, x], 1), torch.cat([m, fake], 1)], 0)
            real_d, fake_d = tutils_mod.divide_pred(D["p"](real_fake))
            l_gan = gan(fake_d, True, False)
            l_feat = featmatch(real_d, fake_d) * lam.featmatch
            (l_vgg + l_gan + l_feat).backward()
            (g_opt.extrapolation if global_step % 2 == 0 else g_opt.step)()
            for p in D.parameters():   # trainer.py:971-973: every parameter, the spectral-norm u / v included
                p.requires_grad_(True)
            logs += [float(l_vgg), float(l_gan), float(l_feat)]
        else:            # update_D
            tutils_mod.zero_grad(D)
            with torch.no_grad():
                fake = G.paint(m, x)
            real_fake = torch.cat([torch.cat([m, x], 1), torch.cat([m, fake.detach()], 1)], 0)
            real_d, fake_d = tutils_mod.divide_pred(D["p"](real_fake))
            l_d = gan(fake_d, False, True) + gan(real_d, True, True)
            l_d.backward()
            (d_opt.extrapolation if global_step % 2 == 0 else d_opt.step)()
            logs += [float(l_d)]
    gsd, dsd = G.painter.state_dict(), D.state_dict()
    arrays = {"logs": np.array(logs, dtype=np.float64),
              "G::conv_img.weight": gsd["conv_img.weight"].numpy(), "G::fc.bias": gsd["fc.bias"].numpy(),
              "G::head_0.conv_0.module.weight_bar": gsd["head_0.conv_0.module.weight_bar"].numpy(),
              "G::final_spade.norm_1.mlp_gamma.weight": gsd["final_spade.norm_1.mlp_gamma.weight"].numpy(),
              "D::p.discriminator_0.model0.0.module.weight_bar": dsd["p.discriminator_0.model0.0.module.weight_bar"].numpy(),
              "D::p.discriminator_1.model3.0.module.bias": dsd["p.discriminator_1.model3.0.module.bias"].numpy()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "latent_dim": latent, "spade_n_up": n_up, "ndf": ndf, "n_layers": n_layers, "num_D": num_d,
            "batch": batch, "size": size, "seeds": {"G": 11, "D": 12, "vgg": 13, "inputs": 5},
            "g_shapes": [[k, list(s)] for k, s in g_shapes], "d_shapes": [[k, list(s)] for k, s in d_shapes],
            "v_shapes": [[k, list(s)] for k, s in v_shapes],
            "log_names": ["G0.vgg", "G0.gan", "G0.featmatch", "D0", "G1.vgg", "G1.gan", "G1.featmatch", "D1"],
            "reference": "cc-ai/climategan @ /root/reference (generator, discriminator, losses, tutils, optim modules)",
            "torch": torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(name, "logs", [round(v, 5) for v in logs], "npz bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def run_masker_case(name="masker_small", nblocks=(2, 2, 3, 2), batch=2, size=64):
    """Masker inference (eval mode) with the reference's OmniGenerator, deeplabv2 encoder + DADA depth + deeplabv2
    segmentation + base mask decoder: G.decode(x) as Trainer.infer_all drives it (trainer.py:272-287)."""
    generator_mod = refshim.load("generator")
    from climategan_b200.utils import default_masker_opts

    opts = default_masker_opts(nblocks=nblocks, size=size)
    opts.data.transforms[-1].new_size.d = 24   # != the decoder's native 16 -> exercises the bicubic(384)+nearest path
    opts.data.transforms[-1].new_size.s = 24   # make_m_cond needs d and s at the same resolution
    torch.manual_seed(0)
    G = generator_mod.OmniGenerator(opts)
    shapes = [(k, tuple(v.shape)) for k, v in G.state_dict().items()]
    G.load_state_dict(fill_state_dict(shapes, seed=77), strict=True)
    G.eval()
    x, _, _ = synth_inputs(batch, size, seed=3)
    with torch.no_grad():
        out = G.decode(x=x, return_z=True, return_z_depth=True)
        cond = G.make_m_cond(out["d"], out["s"], x)
        m_logits = G.mask(z=out["z"], z_depth=out["z_depth"], sigmoid=False)
    arrays = {"m_logits": m_logits.numpy(), "d": out["d"].numpy(), "s": out["s"].numpy(), "m": out["m"].numpy(), "cond": cond.numpy(),
              "z_mean_abs": np.float32(out["z"].abs().mean().item()), "z_sample": out["z"][:, ::97].numpy(),
              "z_depth_sample": out["z_depth"][:, ::97].numpy()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "nblocks": list(nblocks), "batch": batch, "size": size, "d_size": 24, "s_size": 24, "weight_seed": 77,
            "input_seed": 3, "shapes": [[k, list(s)] for k, s in shapes],
            "reference": "cc-ai/climategan @ /root/reference (generator, deeplab/*, depth, masker, blocks modules)",
            "torch": torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, {k: tuple(v.shape) for k, v in out.items() if hasattr(v, "shape")}, "m range", float(out["m"].min()),
          float(out["m"].max()), "npz bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def _flatten_logs(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten_logs(v, prefix + k + "."))
        else:
            out[prefix + k] = float(v)
    return out


def _sample(a, cap=8192):
    """Large tensors are stored as a strided sample of their flattened values (stride = ceil(numel / cap))."""
    a = np.asarray(a).reshape(-1)
    k = max(1, -(-a.size // cap))
    return a[::k].copy()


def run_full_step_case(name="full_step", batch=2, size=128, tasks=("d", "s", "m", "p"), use_spade=False, pl4m=False,
                       overrides=None):
    """Two iterations of the reference's OWN ``Trainer.update_G`` / ``update_D`` (trainer.py:989-1032) on tasks [d,s,m,p]
    — deeplabv2 masker (ResNet [2,2,3,2], train-mode BatchNorm, dropout p=0) + SPADE painter + all three discriminators —
    driven as ``run_epoch`` does (oracle/ref_trainer.py).  Stores every logged loss, the gradient norm of every parameter
    after the first G and D backward, a few full gradients, and a few parameters / BatchNorm running statistics after the
    second iteration (ExtraAdam extrapolation then step)."""
    from oracle import ref_trainer as rt

    opts = rt.full_opts(size=size, tasks=tasks, use_spade=use_spade, overrides=overrides)
    base_depth = opts.gen.d.architecture == "base"
    if use_spade:
        blocks_mod = refshim.load("blocks")
        blocks_mod.SPADEResnetBlock.cuda = lambda self, *a, **k: self   # masker.py:196 hard-codes .cuda() (SURVEY.md §8c patch 1)
    v3 = opts.gen.encoder.architecture == "deeplabv3"
    if v3:   # a shallow ResNet of the same architecture (the reference hard-codes [3,4,23,3], resnet101_v3.py:190-203)
        deeplab_mod, resnet_mod = refshim.load("deeplab", "deeplab.resnet101_v3")
        nb = list(opts.gen.deeplabv3.nblocks)
        deeplab_mod.ResNet101 = lambda output_stride=8, BatchNorm=None, verbose=0, no_init=False: resnet_mod.ResNet(
            resnet_mod.Bottleneck, nb, output_stride, BatchNorm, verbose=verbose, no_init=no_init)
    t = rt.build_reference_trainer(opts, size)
    g_shapes, d_shapes, v_shapes = rt.load_weights(t)
    t.use_pl4m = bool(pl4m)   # what Trainer.train() flips at epoch gen.p.pl4m_epoch (trainer.py:899-909)
    mdb = rt.synth_batch(opts, batch, size, seed=7)
    arrays = {}
    logs = []
    full_g = ["encoder.model.conv1.weight", "encoder.model.layer3.1.conv2.weight", "decoders.s.aspp.aspp3.atrous_conv.weight"]
    if base_depth:   # BaseDepthDecoder: projection, a ResBlock conv, the head
        full_g += ["decoders.d.proj_conv.conv.weight", "decoders.d.model.0.model.0.model.1.conv.weight", "decoders.d.model.0.model.0.model.0.norm.weight"]
    else:
        full_g += ["decoders.d.enc4_2.conv.weight", "decoders.d.enc4_2.norm.weight"]
    full_g += ["decoders.s.conv.%d.bias" % (9 if opts.gen.s.upsample_featuremaps else 8)]
    if use_spade:   # MaskSpadeDecoder: fc_conv, a SPADE layer's three convs, a spectral conv of a block, the mask head
        full_g += ["decoders.m.fc_conv.conv.module.weight_bar", "decoders.m.spade_blocks.0.norm_0.mlp_shared.0.weight",
                   "decoders.m.spade_blocks.1.norm_1.mlp_gamma.weight", "decoders.m.spade_blocks.2.norm_s.mlp_beta.bias",
                   "decoders.m.spade_blocks.1.conv_0.module.weight_bar", "decoders.m.mask_conv.conv.module.weight_bar"]
    else:
        full_g += ["decoders.m.proj_conv.conv.module.weight_bar", "decoders.m.model.6.conv.module.weight_bar"]
    full_d = ["m.Advent.0.module.weight_bar", "s.Advent.8.module.weight_bar"]
    if "p" in tasks:
        full_g += ["painter.conv_img.weight"]
        full_d += ["p.discriminator_0.model0.0.module.weight_bar"]
    if v3:   # deeplabv3 names: backbone without the .model prefix, ASPPv3Plus / Decoder, mask decoder with low-level features
        full_g += ["encoder.conv1.weight", "encoder.layer3.1.conv2.weight", "encoder.layer4.1.conv2.weight",
                   "decoders.s.aspp.conv_out.conv.weight", "decoders.s.decoder.conv_low.conv.bias", "decoders.s.decoder.conv_out.weight",
                   "decoders.m.low_level_conv.conv.module.weight_bar", "decoders.m.merge_feats_conv.conv.module.weight_bar"]
    # keep what this configuration actually has (task subsets, v2 / v3 names)
    gp_names, dp_names = {k for k, _ in t.G.named_parameters()}, {k for k, _ in t.D.named_parameters()}
    full_g = [k for k in full_g if k in gp_names]
    full_d = [k for k in full_d if k in dp_names]
    for it in range(2):
        for p_ in t.D.parameters():
            p_.requires_grad = False
        t.update_G(mdb)
        if it == 0:
            gp = dict(t.G.named_parameters())
            arrays["G.gradnorm"] = np.array([float(p_.grad.norm()) if p_.grad is not None else -1.0 for p_ in gp.values()], dtype=np.float64)
            for k in full_g:
                arrays["G.grad::" + k] = _sample(gp[k].grad.detach().numpy())
        for p_ in t.D.parameters():
            p_.requires_grad = True
        t.update_D(mdb)
        if it == 0:
            dp = dict(t.D.named_parameters())
            arrays["D.gradnorm"] = np.array([float(p_.grad.norm()) if p_.grad is not None else -1.0 for p_ in dp.values()], dtype=np.float64)
            for k in full_d:
                arrays["D.grad::" + k] = _sample(dp[k].grad.detach().numpy())
        t.logger.global_step += 1
        logs.append(_flatten_logs(t.logger.losses.to_dict()))
    gsd, dsd = t.G.state_dict(), t.D.state_dict()
    finals = ["encoder.model.bn1.running_mean", "encoder.model.layer4.0.bn2.running_var", "decoders.s.aspp.global_avg_pool.2.running_var",
              "decoders.d.proj_conv.norm.running_mean" if base_depth else "decoders.d.enc4_1.norm.running_mean"]
    if use_spade:
        finals += ["decoders.m.spade_blocks.0.norm_0.param_free_norm.running_mean", "decoders.m.spade_blocks.2.norm_1.param_free_norm.running_var",
                   "decoders.m.fc_conv.norm.running_var", "decoders.m.spade_blocks.1.conv_1.module.weight_u"]
    else:
        finals += ["decoders.m.model.0.model.1.model.0.conv.module.weight_u"]
    if v3:
        finals += ["encoder.bn1.running_mean", "encoder.layer4.1.bn2.running_var", "decoders.s.aspp.conv_out.bn.running_var"]
    for k in full_g + [f for f in finals if f in gsd]:
        arrays["G.final::" + k] = _sample(gsd[k].numpy())
    for k in full_d:
        arrays["D.final::" + k] = _sample(dsd[k].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "batch": batch, "size": size, "seeds": {"G": 21, "D": 22, "vgg": 23, "inputs": 7},
            "g_shapes": [[k, list(s_)] for k, s_ in g_shapes], "d_shapes": [[k, list(s_)] for k, s_ in d_shapes],
            "tasks": list(tasks), "use_spade": bool(use_spade), "pl4m": bool(pl4m), "overrides": overrides or {},
            "v_shapes": [[k, list(s_)] for k, s_ in (v_shapes or [])], "g_param_names": [k for k, _ in t.G.named_parameters()],
            "d_param_names": [k for k, _ in t.D.named_parameters()], "logs": logs, "full_g": full_g, "full_d": full_d,
            "reference": "cc-ai/climategan @ /root/reference: climategan.trainer.Trainer.update_G/update_D (unmodified), CPU, torch "
                         + torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, "logs[0]", {k: round(v, 5) for k, v in logs[0].items()}, "npz bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def run_infer_all_case(name="infer_all", batch=2, size=320):
    """The reference's OWN ``Trainer.infer_all`` (trainer.py:218-334) on CPU: eval-mode masker (ResNet [2,2,3,2]) + painter,
    then wildfire / smog / flood compositing and the normalize -> uint8 NHWC edge.  ``random.seed(0)`` fixes the wildfire
    filter's green value (fire.py:115); cloudy=False.  uint8 outputs are stored on a stride-2 pixel grid."""
    import random

    from oracle import ref_trainer as rt

    opts = rt.full_opts(size=size)
    t = rt.build_reference_trainer(opts, size, inference=True)
    g_shapes, _, _ = rt.load_weights(t, d=False)
    x = synth_inputs(batch, size, seed=9)[0]
    random.seed(0)
    out = t.infer_all(x.clone(), numpy=True, bin_value=0.5, return_masks=True)
    random.seed(0)
    raw = t.infer_all(x.clone(), numpy=False)
    torch.manual_seed(0)   # the Perlin angles of paint_cloudy come from torch.rand on the CPU generator (tutils.py:660)
    random.seed(0)
    cloudy = t.infer_all(x.clone(), numpy=False, cloudy=True)   # (ignore_event would hit an UnboundLocalError in the reference)
    arrays = {k: v[:, ::2, ::2].copy() for k, v in out.items() if k != "mask"}
    arrays["mask"] = out["mask"][:, :, ::2, ::2].copy()
    arrays["raw_flood_cloudy"] = cloudy["flood"].detach().numpy()[:, :, ::4, ::4].astype(np.float32)
    for k, v in raw.items():
        arrays["raw_" + k] = v.detach().numpy()[:, :, ::4, ::4].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "batch": batch, "size": size, "seeds": {"G": 21, "inputs": 9, "random": 0},
            "g_shapes": [[k, list(s_)] for k, s_ in g_shapes],
            "reference": "cc-ai/climategan @ /root/reference: climategan.trainer.Trainer.infer_all (unmodified), CPU, torch " + torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in arrays.items()}, "npz bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def run_masker_spade_case(name="masker_spade", nblocks=(2, 2, 3, 2), batch=2, size=64, cond_nc=15):
    """The paper / release masker configuration (gen.m.use_spade): reference OmniGenerator.decode in eval mode with the
    MaskSpadeDecoder conditioned on make_m_cond(d, s, x) (masker.py:59-231).  Two consecutive decodes (the spectral-norm
    power iteration advances on every forward)."""
    generator_mod, blocks_mod = refshim.load("generator", "blocks")
    from climategan_b200.utils import Dict, default_masker_opts

    blocks_mod.SPADEResnetBlock.cuda = lambda self, *a, **k: self   # masker.py:196 hard-codes .cuda() (SURVEY.md §8c patch 1)
    opts = default_masker_opts(nblocks=nblocks, size=size)
    opts.gen.m.use_spade = True
    opts.gen.m.spade.activations = Dict(all_lrelu=True)
    opts.gen.m.spade.cond_nc = cond_nc   # 15: normalize(d) | softmax(s) | x ; 12: without x (reference test scenario 14)
    torch.manual_seed(0)
    G = generator_mod.OmniGenerator(opts)
    shapes = [(k, tuple(v.shape)) for k, v in G.state_dict().items()]
    G.load_state_dict(fill_state_dict(shapes, seed=78), strict=True)
    G.eval()
    x, _, _ = synth_inputs(batch, size, seed=4)
    with torch.no_grad():
        out1 = G.decode(x=x)
        out2 = G.decode(x=x)
    arrays = {"m1": out1["m"].numpy(), "m2": out2["m"].numpy(), "d": out1["d"].numpy(), "s": out1["s"].numpy()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "nblocks": list(nblocks), "batch": batch, "size": size, "weight_seed": 78, "input_seed": 4,
            "shapes": [[k, list(s_)] for k, s_ in shapes], "cond_nc": cond_nc,
            "reference": "cc-ai/climategan @ /root/reference (generator, masker.MaskSpadeDecoder, norms.SPADE(batch), blocks)",
            "torch": torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, "m range", float(out1["m"].min()), float(out1["m"].max()), "m1 vs m2", float((out1["m"] - out2["m"]).abs().max()),
          "npz bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def v3_opts(size=128, nblocks=(2, 2, 3, 2)):
    from climategan_b200.utils import default_masker_opts

    opts = default_masker_opts(nblocks=nblocks, size=size)
    opts.gen.encoder.architecture = "deeplabv3"
    opts.gen.s.architecture = "deeplabv3"
    opts.gen.deeplabv3.nblocks = list(nblocks)   # read by climategan_b200 only; the reference is patched below
    return opts


def run_masker_v3_case(name="masker_v3", nblocks=(2, 2, 3, 2), batch=2, size=128, use_spade=False):
    """The reference's DEFAULT masker architecture (defaults.yaml:101,136: deeplabv3 encoder + decoder, ResNet backbone at
    output stride 8, mask decoder with low-level features), shallow copy [2,2,3,2] of the same architecture:
    (i) eval-mode OmniGenerator.decode; (ii) train-mode encode + the three decoders, a fixed random linear functional of
    (d, s, m) back-propagated: gradient norm of every parameter + sampled gradients + updated running statistics."""
    generator_mod, deeplab_mod, resnet_mod = refshim.load("generator", "deeplab", "deeplab.resnet101_v3")
    deeplab_mod.ResNet101 = lambda output_stride=8, BatchNorm=None, verbose=0, no_init=False: resnet_mod.ResNet(
        resnet_mod.Bottleneck, list(nblocks), output_stride, BatchNorm, verbose=verbose, no_init=no_init)
    opts = v3_opts(size, nblocks)
    if use_spade:   # deeplabv3 encoder + MaskSpadeDecoder (low-level / high-level / merge convs, masker.py:118-158, 212-224)
        from climategan_b200.utils import Dict

        refshim.load("blocks").SPADEResnetBlock.cuda = lambda self, *a, **k: self   # masker.py:196 (SURVEY.md §8c patch 1)
        opts.gen.m.use_spade = True
        opts.gen.m.use_proj = True
        opts.gen.m.spade.activations = Dict(all_lrelu=True)
    torch.manual_seed(0)
    G = generator_mod.OmniGenerator(opts, no_init=True)
    shapes = [(k, tuple(v.shape)) for k, v in G.state_dict().items()]
    G.load_state_dict(fill_state_dict(shapes, seed=79), strict=True)
    x, _, _ = synth_inputs(batch, size, seed=6)
    G.eval()
    with torch.no_grad():
        out = G.decode(x=x)
    arrays = {"d": out["d"].numpy(), "s": out["s"].numpy(), "m": out["m"].numpy()}
    G.train()
    for mod in G.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    rs = np.random.RandomState(123)
    z = G.encode(x)
    d, z_depth = G.decoders["d"](z)
    s_ = G.decoders["s"](z, z_depth)
    if use_spade:
        m = G.decoders["m"](z, G.make_m_cond(d, s_, x), z_depth)   # conditioning NOT detached (defaults.yaml:182)
    else:
        m = G.decoders["m"](z, z_depth=z_depth)
    wd, ws, wm = (torch.from_numpy(rs.standard_normal(size=t.shape).astype(np.float32)) for t in (d, s_, m))
    loss = (d * wd).mean() + (s_ * ws).mean() + (m * wm).mean()
    loss.backward()
    names = [k for k, _ in G.named_parameters()]
    arrays["train_loss"] = np.float64(loss.item())
    arrays["train_d"], arrays["train_s"], arrays["train_m"] = d.detach().numpy(), s_.detach().numpy(), m.detach().numpy()
    arrays["gradnorm"] = np.array([float(p_.grad.norm()) if p_.grad is not None else -1.0 for _, p_ in G.named_parameters()])
    gp = dict(G.named_parameters())
    full = ["encoder.conv1.weight", "encoder.layer2.0.conv2.weight", "encoder.layer4.2.conv2.weight", "encoder.layer3.0.bn1.weight",
            "decoders.s.aspp.conv_out.conv.weight", "decoders.s.decoder.conv_low.conv.bias", "decoders.s.decoder.conv_out.weight",
            "decoders.m.low_level_conv.conv.module.weight_bar", "decoders.m.merge_feats_conv.conv.module.weight_bar"]
    if use_spade:
        full += ["decoders.m.high_level_conv.conv.module.weight_bar", "decoders.m.spade_blocks.0.norm_s.mlp_gamma.weight",
                 "decoders.m.spade_blocks.2.conv_1.module.weight_bar", "decoders.d.upsample.2.weight"]
    for k in full:
        arrays["grad::" + k] = _sample(gp[k].grad.detach().numpy())
    sd = G.state_dict()
    finals = ["encoder.bn1.running_mean", "encoder.layer4.1.bn2.running_var", "decoders.s.aspp.conv_out.bn.running_var"]
    if use_spade:
        finals += ["decoders.m.spade_blocks.1.norm_0.param_free_norm.running_var", "decoders.m.merge_feats_conv.norm.running_mean"]
    for k in finals:
        arrays["final::" + k] = sd[k].numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    meta = {"case": name, "nblocks": list(nblocks), "batch": batch, "size": size, "weight_seed": 79, "input_seed": 6,
            "functional_seed": 123, "shapes": [[k, list(s__)] for k, s__ in shapes], "param_names": names, "full": full,
            "use_spade": bool(use_spade),
            "reference": "cc-ai/climategan @ /root/reference (generator, deeplab/resnet101_v3, deeplab/deeplab_v3, depth, masker, blocks)",
            "torch": torch.__version__}
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, {k: tuple(v.shape) for k, v in out.items() if hasattr(v, "shape")}, "loss", float(loss), "npz bytes",
          os.path.getsize(os.path.join(HERE, name + ".npz")))


# Option combinations around the reference's scenario matrix (tests/test_trainer.py:205-308) that have no full fixture: only the
# logged losses of two Trainer iterations and the first-iteration gradient norms are kept (a few KB each).
SWEEP = {
    "dada_ms": dict(tasks=("d", "s", "m"), overrides={"gen.m.use_dada": True}),
    "base_depth_regression": dict(tasks=("d", "s", "m"), overrides={"gen.d.architecture": "base", "gen.m.use_dada": False,
                                                                    "gen.s.use_dada": False}),
    "v3_spade_msdp": dict(tasks=("d", "s", "m", "p"), use_spade=True, overrides=dict(V3_MASKER)),
    "spade_detached_cond": dict(tasks=("d", "s", "m"), use_spade=True, overrides={"gen.m.spade.detach": True}),
    "adam": dict(tasks=("d", "s", "m", "p"), overrides={"gen.opt.optimizer": "Adam", "dis.opt.optimizer": "Adam"}),
    "pseudo_labels": dict(tasks=("d", "s", "m"), overrides={"train.pseudo.tasks": ["d", "s"]}),
    "minent_v1_no_gi": dict(tasks=("d", "s", "m"), overrides={"gen.m.use_minent_var": False, "gen.m.use_ground_intersection": False}),
    "depth_and_seg_only": dict(tasks=("d", "s")),
    "dada_depth_loss": dict(tasks=("d", "s", "m"), overrides={"gen.d.loss": "dada"}),
    # painter options off in defaults.yaml: the global + local discriminator pair (dis.p.use_local_discriminator), the same with
    # the painter loss for the masker on, and the tv / context / reconstruction losses with non-zero weights
    "painter_local_d": dict(tasks=("d", "s", "m", "p"), overrides={"dis.p.use_local_discriminator": True}),
    "painter_local_d_pl4m": dict(tasks=("d", "s", "m", "p"), overrides={"dis.p.use_local_discriminator": True}, pl4m=True),
    "painter_aux_losses": dict(tasks=("p",), overrides={"train.lambdas.G.p.tv": 1.0, "train.lambdas.G.p.context": 5.0,
                                                        "train.lambdas.G.p.reconstruction": 2.0}),
}


# gen.p.diff_aug draws from torch's global generator in the middle of the step (transforms.py:493-606): the fixture and the tests
# seed it before every update_G / update_D ("reseed"), after which both sides consume it identically
SWEEP_DIFFAUG = {
    "painter_diff_aug": dict(tasks=("p",), overrides={"gen.p.diff_aug.use": True, "gen.p.diff_aug.do_color_jittering": True,
                                                      "gen.p.diff_aug.do_cutout": True, "gen.p.diff_aug.cutout_ratio": 0.5,
                                                      "gen.p.diff_aug.do_translation": True,
                                                      "gen.p.diff_aug.translation_ratio": 0.125}),
}


def run_config_sweep(name="config_sweep", batch=2, size=128, sweep=None, reseed=False):
    from oracle import ref_trainer as rt

    refshim.load("blocks").SPADEResnetBlock.cuda = lambda self, *a, **k: self   # masker.py:196 (SURVEY.md §8c patch 1)
    meta, arrays = {"batch": batch, "size": size, "seeds": {"G": 21, "D": 22, "vgg": 23, "inputs": 7}, "cases": {}}, {}
    for case, kw in (sweep or SWEEP).items():
        kw = dict(kw)
        pl4m = kw.pop("pl4m", False)
        opts = rt.full_opts(size=size, **kw)
        if opts.gen.encoder.architecture == "deeplabv3":
            deeplab_mod, resnet_mod = refshim.load("deeplab", "deeplab.resnet101_v3")
            nb = list(opts.gen.deeplabv3.nblocks)
            deeplab_mod.ResNet101 = lambda output_stride=8, BatchNorm=None, verbose=0, no_init=False, nb=nb: resnet_mod.ResNet(
                resnet_mod.Bottleneck, nb, output_stride, BatchNorm, verbose=verbose, no_init=no_init)
        t = rt.build_reference_trainer(opts, size)
        g_shapes, d_shapes, v_shapes = rt.load_weights(t)
        t.use_pl4m = bool(pl4m)
        if "p" in opts.tasks and opts.gen.p.diff_aug.use:                            # Trainer.setup, trainer.py:772-773
            t.diff_transforms = refshim.load("transforms").DiffTransforms(opts.gen.p.diff_aug)
        mdb = rt.synth_batch(opts, batch, size, seed=7)
        logs = []
        for it in range(2):
            for p_ in t.D.parameters():
                p_.requires_grad = False
            if reseed:
                torch.manual_seed(1000 + it)
            t.update_G(mdb)
            if it == 0:
                arrays[case + "::G.gradnorm"] = np.array([float(p_.grad.norm()) if p_.grad is not None else -1.0
                                                          for p_ in t.G.parameters()])
            for p_ in t.D.parameters():
                p_.requires_grad = True
            if t.d_opt is not None:
                if reseed:
                    torch.manual_seed(2000 + it)
                t.update_D(mdb)
                if it == 0:
                    arrays[case + "::D.gradnorm"] = np.array([float(p_.grad.norm()) if p_.grad is not None else -1.0
                                                              for p_ in t.D.parameters()])
            t.logger.global_step += 1
            logs.append(_flatten_logs(t.logger.losses.to_dict()))
        meta["cases"][case] = {"kw": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}, "logs": logs, "pl4m": bool(pl4m),
                               "g_shapes": [[k, list(s_)] for k, s_ in g_shapes], "d_shapes": [[k, list(s_)] for k, s_ in d_shapes],
                               "v_shapes": [[k, list(s_)] for k, s_ in (v_shapes or [])]}
        print(case, {k: round(v, 5) for k, v in list(logs[0].items())[:6]})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")), os.path.getsize(os.path.join(HERE, name + ".json")))


if __name__ == "__main__":
    if not refshim.available():
        sys.exit("reference tree not available; goldens can only be regenerated in the build container")
    for name, cfg in CASES.items():
        run_case(name, *cfg)
    run_painter_z_case()
    run_disc_case()
    run_step_case()
    run_masker_case()
    run_full_step_case()
    run_full_step_case(name="masker_step_spade", tasks=("d", "s", "m"), use_spade=True)
    run_full_step_case(name="full_step_pl4m", pl4m=True)
    run_full_step_case(name="masker_step_base_depth_classify", tasks=("d", "s", "m"), overrides=BASE_DEPTH_CLASSIFY)
    run_full_step_case(name="masker_step_v3", tasks=("d", "s", "m"), overrides=V3_MASKER)
    run_full_step_case(name="mask_only_step_v3", tasks=("m",), overrides=V3_MASKER)
    run_infer_all_case()
    run_masker_spade_case()
    run_masker_spade_case(name="masker_spade12", cond_nc=12)
    run_masker_v3_case()
    run_masker_v3_case(name="masker_v3_spade", use_spade=True)
    run_config_sweep()
    run_config_sweep(name="config_sweep_diffaug", sweep=SWEEP_DIFFAUG, reseed=True)
