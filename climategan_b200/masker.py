"""Flood-mask decoders (``climategan/masker.py``): MaskBaseDecoder (:25-56) = BaseDecoder (``blocks.py:206-318``) with
the masker options.  The SPADE mask decoder (:59-231) is not built."""
from __future__ import annotations

from .blocks import BaseDecoder


def create_mask_decoder(opts, no_init=False, verbose=0):
    if opts.gen.m.use_spade:
        raise NotImplementedError("MaskSpadeDecoder (gen.m.use_spade) is not built")
    return MaskBaseDecoder(opts)


class MaskBaseDecoder(BaseDecoder):
    def __init__(self, opts):
        if opts.gen.encoder.architecture == "deeplabv3":
            raise NotImplementedError("the deeplabv3 encoder / low-level-feature branch is not built")
        use_dada = ("d" in opts.tasks) and opts.gen.m.use_dada
        super().__init__(n_upsample=opts.gen.m.n_upsample, n_res=opts.gen.m.n_res, input_dim=2048,
                         proj_dim=opts.gen.m.proj_dim, output_dim=opts.gen.m.output_dim, norm=opts.gen.m.norm,
                         activ=opts.gen.m.activ, pad_type=opts.gen.m.pad_type, output_activ="none",
                         low_level_feats_dim=-1, use_dada=use_dada)
