"""GAN-side losses of the painter step — ``climategan/losses.py`` GANLoss (:13-83), FeatMatchLoss (:86-103), HingeLoss
(:550-593) — with the reference call signatures, evaluated by libcgb200 reduce kernels (value + gradient in one pass).
"""
from __future__ import annotations

from random import random as rand

import torch
import torch.nn as nn

from . import _lib, ops


class GANLoss(nn.Module):
    """BCE-with-logits (use_lsgan=False) or MSE (use_lsgan=True) against a constant real/fake label, with the
    reference's label smoothing (``soft_shift``) and label flipping (``flip_prob``) draws in the same order."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, soft_shift=0.0, flip_prob=0.0,
                 verbose=0):
        super().__init__()
        self.soft_shift = soft_shift
        self.verbose = verbose
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.kind = ops.LOSS_MSE if use_lsgan else ops.LOSS_BCE_LOGITS
        self.flip_prob = flip_prob

    def get_target_value(self, target_is_real):
        soft_change = float(torch.FloatTensor(1).uniform_(0, self.soft_shift))  # same RNG draw as losses.py:57
        return float(self.real_label) - soft_change if target_is_real else float(self.fake_label) + soft_change

    def _draw_targets(self, n, target_is_real):
        """The host draws of one call, in the reference's order (losses.py:59-83): one ``random()`` for the flip decision, then
        per prediction the (cumulative) flip and one ``uniform_`` for the soft label."""
        r = rand()
        vals = []
        for _ in range(n):
            if r < self.flip_prob:
                target_is_real = not target_is_real
            vals.append(self.get_target_value(target_is_real))
        return vals

    def __call__(self, input, target_is_real, *args, **kwargs):
        from . import graphs

        preds = [p[-1] if isinstance(p, list) else p for p in input] if isinstance(input, list) else [input]
        n = len(preds)
        tape = graphs.current_tape()
        if tape is None:
            targets = self._draw_targets(n, target_is_real)
        else:   # capturing a CUDA graph: the labels live in device memory and are re-drawn before every replay
            targets = tape.floats(lambda: self._draw_targets(n, target_is_real), n)
        loss = 0
        for pred_i, t in zip(preds, targets):
            loss = loss + ops.const_target_loss(pred_i, self.kind, t)
        return loss / n if isinstance(input, list) else loss


class HingeLoss(nn.Module):
    def __init__(self, tensor=None):
        super().__init__()

    def loss(self, input, target_is_real, for_discriminator=True):
        if for_discriminator:
            return ops.const_target_loss(input, ops.LOSS_HINGE_D_REAL if target_is_real else ops.LOSS_HINGE_D_FAKE)
        assert target_is_real, "The generator's hinge loss must be aiming for real"
        return ops.const_target_loss(input, ops.LOSS_NEG_MEAN)

    def __call__(self, input, target_is_real, for_discriminator=True):
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                loss = loss + self.loss(pred_i, target_is_real, for_discriminator)
            return loss / len(input)
        return self.loss(input, target_is_real, for_discriminator)


class FeatMatchLoss(nn.Module):
    """L1 between the discriminator's intermediate features of the fake and (detached) real halves, summed over
    layers, divided by num_D.  Accepts NCHW fp32 feature lists (reference contract)."""

    def __call__(self, pred_real, pred_fake):
        num_D = len(pred_fake)
        total = 0.0
        for i in range(num_D):
            for j in range(len(pred_fake[i]) - 1):
                total = total + ops.l1_loss(pred_fake[i][j], pred_real[i][j].detach()) / num_D
        return total


_VGG19_CFG = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256), (12, 256, 256), (14, 256, 256),
              (16, 256, 256), "M", (19, 256, 512), (21, 512, 512), (23, 512, 512), (25, 512, 512), "M", (28, 512, 512)]
_VGG19_SLICES = [(0, 2), (2, 7), (7, 12), (12, 21), (21, 30)]


class Vgg19(nn.Module):
    """``climategan.losses.Vgg19`` (losses.py:304-336): torchvision vgg19.features[:30] cut into five slices ending at
    relu1_1, relu2_1, relu3_1, relu4_1, relu5_1.  Same state_dict keys (``slice{k}.{idx}.weight``).  The reference loads
    ImageNet weights (``pretrained=True``, a download); here the weights are whatever is loaded into the module
    (random init in the bench) — frozen, so only data gradients flow."""

    def __init__(self, requires_grad=False, storage_dtype=torch.bfloat16):
        super().__init__()
        self.storage_dtype = storage_dtype
        layers = {}
        for item in _VGG19_CFG:
            if item != "M":
                idx, ci, co = item
                layers[idx] = nn.Conv2d(ci, co, 3, padding=1)
        for k, (a, b) in enumerate(_VGG19_SLICES, start=1):
            seq = nn.Sequential()
            for idx in range(a, b):
                if idx in layers:
                    seq.add_module(str(idx), layers[idx])
                elif any(it == "M" for it in _VGG19_CFG) and idx in (4, 9, 18, 27):
                    seq.add_module(str(idx), nn.MaxPool2d(2, 2))
                else:
                    seq.add_module(str(idx), nn.ReLU(inplace=True))
            setattr(self, f"slice{k}", seq)
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def forward_storage(self, x):
        """x: storage [N,H,W,8] (vgg-preprocessed BGR).  Returns the five relu feature maps (storage tensors)."""
        outs = []
        for k in range(1, 6):
            for m in getattr(self, f"slice{k}"):
                if isinstance(m, nn.Conv2d):
                    if m.in_channels <= 4:   # conv1_1 on the image: im2col + one K = 32 GEMM
                        x = ops.conv2d_first_layer(x, m.weight, m.bias, pad=1, act=_lib.ACT_RELU)
                    else:
                        x = ops.conv2d(x, m.weight, m.bias, pad=1, act=_lib.ACT_RELU)  # conv + the ReLU that follows it
                elif isinstance(m, nn.MaxPool2d):
                    x = ops.maxpool2(x)
            outs.append(x)
        return outs


class VGGLoss(nn.Module):
    """``climategan.losses.VGGLoss`` (losses.py:338-350): sum_i w_i * L1(vgg(x)_i, vgg(y)_i.detach()), w = 1/32..1.
    ``forward(x, y, mask=None)`` takes the NCHW fp32 images *before* ``vgg_preprocess`` (the preprocess and the
    optional ``* m`` of trainer.py:1281-1283 run inside one kernel)."""

    def __init__(self, device=None, storage_dtype=torch.bfloat16, weights=None):
        """``weights``: a path to (or a state_dict of) torchvision's ``vgg19`` ImageNet weights — ``features.N.weight`` keys, as
        ``torchvision.models.vgg19(pretrained=True)`` saves them, or this module's own ``sliceK.N.weight`` keys.  The reference
        downloads them (losses.py:307); this package never touches the network, so WITHOUT ``weights`` (or the environment
        variable ``CGB_VGG19_WEIGHTS``, or an installed torchvision checkpoint in the torch hub cache) the features are those of
        a randomly initialised VGG and a loud warning says so — load real weights with :meth:`load_vgg19_weights` before
        training for real."""
        super().__init__()
        import os

        self.vgg = Vgg19(storage_dtype=storage_dtype).eval()
        self.pretrained = False
        weights = weights or os.environ.get("CGB_VGG19_WEIGHTS")
        if weights is None:
            hub = os.path.join(os.path.expanduser(os.environ.get("TORCH_HOME", "~/.cache/torch")), "hub", "checkpoints")
            if os.path.isdir(hub):
                cands = sorted(f for f in os.listdir(hub) if f.startswith("vgg19-") and f.endswith(".pth"))
                weights = os.path.join(hub, cands[-1]) if cands else None
        if weights is not None:
            self.load_vgg19_weights(weights)
        else:
            import warnings

            warnings.warn("VGGLoss: no ImageNet VGG19 weights were given (weights=..., CGB_VGG19_WEIGHTS, or torch hub cache): the "
                          "perceptual loss runs on a RANDOMLY INITIALISED VGG19. Fine for benchmarks and parity tests, wrong for "
                          "training (the reference uses torchvision's pretrained vgg19, losses.py:307).", stacklevel=2)
        if device is not None:
            self.vgg.to(device)
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
        self.channels = [64, 128, 256, 512, 512]

    def load_vgg19_weights(self, weights):
        """Load torchvision vgg19 weights (``features.N.*`` keys) or this module's ``sliceK.N.*`` keys."""
        sd = torch.load(weights, map_location="cpu") if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__") else weights
        own = self.vgg.state_dict()
        idx_to_key = {k.split(".")[1]: k.rsplit(".", 1)[0] for k in own}     # "0" -> "slice1.0"
        mapped = {}
        for k, v in sd.items():
            if k in own:
                mapped[k] = v
            elif k.startswith("features."):
                _, idx, leaf = k.split(".")
                if idx in idx_to_key:
                    mapped[f"{idx_to_key[idx]}.{leaf}"] = v
        missing = sorted(set(own) - set(mapped))
        if missing:
            raise KeyError(f"VGG19 weights: {len(missing)} tensors missing, e.g. {missing[:3]}")
        self.vgg.load_state_dict(mapped, strict=True)
        self.pretrained = True
        ops.invalidate_weight_cache()
        return self

    def forward(self, x, y, mask=None):
        dt = self.vgg.storage_dtype
        x_vgg = self.vgg.forward_storage(ops.vgg_preprocess(x, mask, dt))
        with torch.no_grad():
            y_vgg = self.vgg.forward_storage(ops.vgg_preprocess(y, mask, dt))
        loss = 0
        for w, a, b, c in zip(self.weights, x_vgg, y_vgg, self.channels):
            loss = loss + w * ops.l1_loss_storage(a, b, c)
        return loss


# ------------------------------------------------------------------------------------------------
# masker losses — climategan/losses.py CrossEntropy :106-112, TVLoss :140-171, MinentLoss :177-196, SIGMLoss :232-278,
# GroundIntersectionLoss :449-455, CustomBCELoss :466-477, ADVENTAdversarialLoss :480-524 — same call signatures, NCHW
# fp32 tensors in, each evaluated by one or two libcgb200 kernels that emit the scalar and its gradient together.
# ------------------------------------------------------------------------------------------------
class CrossEntropy(nn.Module):
    def __call__(self, logits, target):
        return ops.cross_entropy_nchw(logits, target.to(logits.device).long())


class DADADepthLoss:
    """``climategan.losses.DADADepthLoss`` (:596-620)."""

    def __call__(self, pred, label):
        return ops.dada_depth_loss(pred, label)


class TVLoss(nn.Module):
    def __init__(self, tvloss_weight=1):
        super().__init__()
        self.tvloss_weight = tvloss_weight

    def forward(self, x):
        return self.tvloss_weight * ops.tv_loss(x)


class MinentLoss(nn.Module):
    def __init__(self, version=1, lambda_var=0.1):
        super().__init__()
        self.version = version
        self.lambda_var = lambda_var

    def __call__(self, pred):
        assert pred.dim() == 4
        return ops.minent_loss(pred, self.version, self.lambda_var)


class SIGMLoss(nn.Module):
    def __init__(self, gmweight=0.5, scale=4, device="cuda"):
        super().__init__()
        self.gmweight = gmweight
        self.scale = scale

    def __call__(self, prediction, target):
        return ops.sigm_loss(prediction, target, self.gmweight, self.scale)


class BCEWithLogits(nn.Module):
    """nn.BCEWithLogitsLoss() with a tensor target (losses.py:419)."""

    def __call__(self, prediction, target):
        return ops.bce_logits_loss(prediction, target)


class GroundIntersectionLoss(nn.Module):
    def __call__(self, pred, pseudo_ground):
        return ops.ground_intersection_loss(pred, pseudo_ground)


class CustomBCELoss(nn.Module):
    """BCE-with-logits against an int domain label (losses.py:466-477)."""

    def __call__(self, prediction, target):
        return ops.const_target_loss(prediction, ops.LOSS_BCE_LOGITS, float(target))


class ADVENTAdversarialLoss(nn.Module):
    """losses.py:480-524.  gan_type "GAN" -> CustomBCELoss; anything else -> the WGAN form
    ``-mean(y*x + (1-y)*(1-x))`` (the reference's ``elif gan_type == "WGAN" or "WGAN_gp" or "WGAN_norm"`` is always true)."""

    def __init__(self, opts, gan_type="GAN"):
        super().__init__()
        self.opts = opts
        self.gan_type = gan_type
        self.bce = CustomBCELoss() if gan_type == "GAN" else None

    def loss(self, d_out, target):
        if self.bce is not None:
            return self.bce(d_out, target)
        y = float(target)
        neg_mean = ops.const_target_loss(d_out, ops.LOSS_NEG_MEAN)   # -mean(x)
        # -mean(y x + (1-y)(1-x)) = (2y-1)*(-mean x) - (1-y)
        return (2.0 * y - 1.0) * neg_mean - (1.0 - y)

    def __call__(self, prediction, target, discriminator, depth_preds=None):
        d_in = ops.prob_2_entropy(prediction, depth_preds)
        d_out = discriminator(d_in)
        if self.opts.dis.m.architecture == "OmniDiscriminator":
            raise NotImplementedError("dis.m.architecture=OmniDiscriminator (multiDiscriminatorAdapter) is not built")
        return self.loss(d_out, target)
