"""OmniGenerator — drop-in surface for ``climategan/generator.py``.  Built so far: the painter
path (``paint``, ``sample_painter_z``); the masker methods raise until their kernels land.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .painter import create_painter


def create_generator(opts, device="cuda", latent_shape=None, no_init=False, verbose=0,
                     storage_dtype: torch.dtype = torch.bfloat16):
    """generator.py:24-61 (the painter needs no init_weights pass in the reference either)."""
    G = OmniGenerator(opts, latent_shape, verbose, no_init, storage_dtype=storage_dtype)
    return G.to(device)


class OmniGenerator(nn.Module):
    def __init__(self, opts, latent_shape=None, verbose=0, no_init=False, storage_dtype=torch.bfloat16):
        super().__init__()
        self.opts = opts
        self.verbose = verbose
        self.encoder = None
        if any(t in opts.tasks for t in "msd"):
            raise NotImplementedError("masker tasks (m, s, d) are not built yet in climategan_b200")
        self.decoders = nn.ModuleDict({})
        self.painter = nn.Module()
        if "p" in self.opts.tasks:
            self.painter = create_painter(opts, no_init, verbose)
            self.painter.storage_dtype = storage_dtype
            if latent_shape is not None:
                self.painter.set_latent_shape(latent_shape, True)

    def sample_painter_z(self, batch_size, device, force_half=False):
        """generator.py (no_z=True in defaults.yaml:148 -> None)."""
        if self.opts.gen.p.no_z:
            return None
        raise NotImplementedError("gen.p.no_z=False is not built")

    def paint(self, m, x, no_paste=False):
        """generator.py:279-297: fake = painter(z, x*(1-m)); return x*(1-m) + fake*m."""
        z_paint = self.sample_painter_z(x.shape[0], x.device)
        assert z_paint is None
        p = self.painter
        cond = ops.mask_cond(x, m.to(x.dtype), p.storage_dtype)
        fake = ops.from_storage(p.forward_storage(cond), 3)
        if self.opts.gen.p.paste_original_content and not no_paste:
            return ops.paste(x, m.to(x.dtype), fake)
        return fake
