"""climategan_b200 — B200-native (sm_100a) implementation of the ClimateGAN conv-GAN hot path.

Host side: Python mirroring the reference's module surface (``painter``, ``generator``, ``blocks``,
``norms``); compute: hand-written CUDA kernels behind the C ABI in ``include/cgb200.h``
(``libcgb200.so``, built in-tree by :func:`climategan_b200._lib.build`).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
