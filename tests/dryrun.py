"""Dry-run harness for the HOST side of the product (tests only): replaces the loaded libcgb200 handle by a stand-in that
type-checks every argument against the declared ctypes signature (``_lib.SIGNATURES``) and returns success without computing, so
the whole Python path of a train step — module forwards, autograd Functions, weight packing / caching, the flat optimiser —
runs on CPU tensors.  Values are meaningless (outputs stay uninitialised); shapes, dtypes, layouts, argument marshalling and
the call sequence are real.  The pure host-side queries (workspace sizes) go to the real library."""
from __future__ import annotations

import collections
import contextlib
import types

import torch

from climategan_b200 import _lib, ops

_HOST_SIDE = {"cgb_instnorm_ws_doubles", "cgb_bn_bwd_ws_doubles", "cgb_version", "cgb_last_error"}


class NoopLib:
    def __init__(self, real):
        self._real = real
        self.calls = collections.Counter()

    def __getattr__(self, name):
        real_fn = getattr(self._real, name)   # AttributeError for a symbol the library does not export
        if name in _HOST_SIDE:
            return real_fn
        if name == "cgb_device_ok":
            return lambda: 1
        if name == "cgb_conv2d_uses_tcgen05":
            return lambda d, which: 0
        if name in ("cgb_launch_count",):
            return lambda: sum(self.calls.values())
        argtypes = real_fn.argtypes or []

        def call(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments for {len(argtypes)} parameters"
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    t.from_param(a)
                except Exception as e:  # noqa: BLE001
                    raise TypeError(f"{name}: argument {i} ({a!r}) does not convert to {t}") from e
            self.calls[name] += 1
            return 0

        return call


@contextlib.contextmanager
def noop_library():
    """with noop_library() as lib: ... ; lib.calls counts the launches that would have happened."""
    real = _lib.lib()
    fake = NoopLib(real)
    saved = (_lib._lib, ops._on_device, torch.cuda.current_stream)
    _lib._lib = fake
    ops._on_device = lambda x: True   # (the storage-layout checks of ops._chk_storage stay in force)
    torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0)
    try:
        yield fake
    finally:
        _lib._lib, ops._on_device, torch.cuda.current_stream = saved
        ops.invalidate_weight_cache()
