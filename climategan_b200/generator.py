"""OmniGenerator — drop-in surface for ``climategan/generator.py``: the painter path (``paint``; training and
inference) and the v2 masker path (``encode``, ``decode``, ``depth``, ``make_m_cond``, ``mask``: DeepLab-v2
ResNet-101 encoder, DADA depth decoder, DeepLab-v2 segmentation decoder, base mask decoder) in train mode (autograd tape,
batch-statistics BatchNorm) and eval mode (fused inference kernels).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .deeplab import create_encoder, create_segmentation_decoder
from .depth import create_depth_decoder
from .masker import create_mask_decoder
from .painter import create_painter


def create_generator(opts, device="cuda", latent_shape=None, no_init=False, verbose=0,
                     storage_dtype: torch.dtype = torch.bfloat16):
    """generator.py:24-61: every decoder except the segmentation one, and a "base" encoder, get ``init_weights`` with their
    ``opts.gen[model].init_type / init_gain`` unless ``no_init`` (the painter and the DeepLab encoders keep their constructors'
    initialisation, as in the reference)."""
    G = OmniGenerator(opts, latent_shape, verbose, no_init, storage_dtype=storage_dtype)
    if no_init:
        return G.to(device)
    from .discriminator import init_weights

    for model in G.decoders:
        if model == "s":
            continue
        init_weights(G.decoders[model], init_type=opts.gen[model].init_type, init_gain=opts.gen[model].init_gain, verbose=verbose,
                     caller=f"create_generator decoder {model}")
    if G.encoder is not None and opts.gen.encoder.architecture == "base":
        init_weights(G.encoder, init_type=opts.gen.encoder.init_type, init_gain=opts.gen.encoder.init_gain, verbose=verbose,
                     caller="create_generator encoder")
    return G.to(device)


class OmniGenerator(nn.Module):
    def __init__(self, opts, latent_shape=None, verbose=0, no_init=False, storage_dtype=torch.bfloat16):
        super().__init__()
        self.opts = opts
        self.verbose = verbose
        self.encoder = None
        self.storage_dtype = storage_dtype
        if any(t in opts.tasks for t in "msd"):
            self.encoder = create_encoder(opts, no_init, verbose)
        decoders = {}
        self.painter = nn.Module()   # registered before the decoders, as in the reference (fixes the parameter order)
        if "d" in opts.tasks:
            decoders["d"] = create_depth_decoder(opts, no_init, verbose)
        if "s" in opts.tasks:
            decoders["s"] = create_segmentation_decoder(opts, no_init, verbose)
        if "m" in opts.tasks:
            decoders["m"] = create_mask_decoder(opts, no_init, verbose)
        self.decoders = nn.ModuleDict(decoders)
        if "p" in self.opts.tasks:
            self.painter = create_painter(opts, no_init, verbose)
            self.painter.storage_dtype = storage_dtype
            if latent_shape is not None:
                self.painter.set_latent_shape(latent_shape, True)

    def sample_painter_z(self, batch_size, device, force_half=False):
        """generator.py (no_z=True in defaults.yaml:148 -> None)."""
        if self.opts.gen.p.no_z:
            return None
        z = torch.empty(batch_size, self.opts.gen.p.latent_dim, self.painter.z_h, self.painter.z_w,
                        device=device).normal_(mean=0, std=1.0)   # generator.py:183-194
        return z.half() if force_half else z

    def paint(self, m, x, no_paste=False):
        """generator.py:279-297: fake = painter(z, x*(1-m)); return x*(1-m) + fake*m."""
        z_paint = self.sample_painter_z(x.shape[0], x.device)
        p = self.painter
        cond = ops.mask_cond(x, m.to(x.dtype), p.storage_dtype)
        z_st = ops.to_storage(z_paint.float(), p.storage_dtype) if z_paint is not None else None
        fake = ops.from_storage(p.forward_storage(cond, z_st), 3)
        if self.opts.gen.p.paste_original_content and not no_paste:
            return ops.paste(x, m.to(x.dtype), fake)
        return fake

    def paint_cloudy(self, m, x, s, sky_idx=9, res=(8, 8), weight=0.8):
        """generator.py:299-328: paint through an intermediary image whose sky (argmax of the upsampled seg logits == sky_idx)
        is replaced by Perlin noise, then paste the original content back."""
        from . import events

        noised_x = events.cloudy_input(x, s, sky_idx, res, weight)
        fake = self.paint(m, noised_x, no_paste=True)
        return ops.paste(x, m.to(x.dtype), fake)

    # ------------------------------------------------------------------ masker
    # z and z_depth are NHWC storage tensors (they only ever travel between these methods); d, s, m are NCHW fp32 like the
    # reference's.  Train mode records the autograd tape (BatchNorm on batch statistics, dropout active); eval mode runs the
    # fused inference kernels under no_grad.
    def _grad_ctx(self):
        return torch.enable_grad() if (self.training and torch.is_grad_enabled()) else torch.no_grad()

    def encode(self, x):
        """generator.py:107-118.  NCHW fp32 image -> z [N,H/8,W/8,2048] (storage layout)."""
        assert self.encoder is not None
        with self._grad_ctx():
            return self.encoder.forward_storage(ops.to_storage(x, self.storage_dtype))

    def decode_d(self, z):
        """``self.decoders["d"](z)`` of the reference (depth.py:128-155): (d NCHW fp32 [N,1,T,T], z_depth storage)."""
        with self._grad_ctx():
            dec = self.decoders["d"]
            d, z_depth = dec.forward_storage(z)
            return ops.from_storage(d, getattr(dec, "output_dim", 1)), z_depth   # (> 1: bucket logits, gen.d.classify)

    def decode_s(self, z, z_depth=None):
        """``self.decoders["s"](z, z_depth)`` (deeplab_v2.py:181-198): seg logits NCHW fp32 [N,11,T,T]."""
        with self._grad_ctx():
            dec = self.decoders["s"]
            return ops.from_storage(dec.forward_storage(z, z_depth), dec.output_dim)

    def decode_m(self, z, cond=None, z_depth=None):
        """``self.decoders["m"](z, cond=cond, z_depth=z_depth)`` (blocks.py:291-318): mask logits NCHW fp32 [N,1,H,W]."""
        with self._grad_ctx():
            return ops.from_storage(self.decoders["m"].forward_storage(z, cond, z_depth), 1)

    def depth(self, x=None, z=None, return_z_depth=False):
        """generator.py:330-355."""
        assert x is not None or z is not None
        assert not (x is not None and z is not None)
        if z is None:
            z = self.encode(x)
        d, z_depth = self.decode_d(z)
        return (d, z_depth) if return_z_depth else d

    def make_m_cond(self, d, s, x=None):
        """generator.py:196-230: d, s NCHW fp32 predictions; returns the NCHW conditioning tensor (12 or 15 channels)."""
        if self.opts.gen.m.spade.cond_nc == 15 and x is None:
            raise ValueError("When using spade for the Masker with 15 channels, x MUST be provided")
        with self._grad_ctx():
            if self.opts.gen.m.spade.detach:
                d, s = d.detach(), s.detach()
            dt = self.storage_dtype
            ds, ss = ops.to_storage(d, dt), ops.to_storage(s, dt)
            xr = None
            if self.opts.gen.m.spade.cond_nc == 15:
                xr = ops.resize_bilinear(ops.to_storage(x, dt), s.shape[-2], s.shape[-1], align_corners=True)
            cond = ops.make_m_cond(ds, ss, xr, s.shape[1])
            return ops.from_storage(cond, 1 + s.shape[1] + (3 if xr is not None else 0))

    def mask(self, x=None, z=None, cond=None, z_depth=None, sigmoid=True):
        """generator.py:232-277 (base mask decoder: cond is accepted and unused, as in BaseDecoder.forward)."""
        assert x is not None or z is not None
        if z is None:
            z = self.encode(x)
        if cond is None and self.opts.gen.m.use_spade:
            # generator.py:257-262: the SPADE mask decoder's conditioning is built here, un-taped, from the depth and
            # segmentation predictions (x is needed when cond_nc == 15: make_m_cond raises otherwise, as the reference does)
            assert "s" in self.opts.tasks and "d" in self.opts.tasks
            with torch.no_grad():
                d_pred, z_d = self.decode_d(z)
                s_pred = self.decode_s(z, z_d)
                cond = self.make_m_cond(d_pred, s_pred, x)
        with self._grad_ctx():
            if z_depth is None and self.opts.gen.m.use_dada:
                with torch.no_grad():                                    # generator.py:263-266
                    _, z_depth = self.decoders["d"].forward_storage(z)
            logits = self.decoders["m"].forward_storage(z, cond, z_depth)
            if sigmoid:
                logits = ops.activation(logits, _lib.ACT_SIGMOID)
            return ops.from_storage(logits, 1)

    def decode(self, x=None, z=None, return_z=False, return_z_depth=False):
        """generator.py:120-176."""
        assert x is not None or z is not None
        out = {}
        z_depth = cond = d = s = None
        if z is None:
            z = self.encode(x)
        if return_z:
            out["z"] = z
        if "d" in self.decoders:
            d, z_depth = self.decode_d(z)
            out["d"] = d
        if return_z_depth:
            out["z_depth"] = z_depth
        if "s" in self.decoders:
            s = out["s"] = self.decode_s(z, z_depth)
        if "m" in self.decoders:
            if s is not None and d is not None and self.opts.gen.m.use_spade:
                cond = self.make_m_cond(d, s, x)
            out["m"] = self.mask(z=z, cond=cond, z_depth=z_depth)
        return out
