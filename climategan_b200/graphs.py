"""CUDA-graph replay of the train step's forward + backward (``Trainer.enable_cuda_graphs``).

``Trainer.update_G`` / ``update_D`` (climategan/trainer.py:989-1032) issue ~5 k library launches and ~4 k small tensor ops per
step from Python; captured once, the same work replays from one ``cudaGraphLaunch`` and the host leaves the critical path.
What a capture must not bake in are the values the reference draws on the HOST every step:

* ``GANLoss``'s label flip / soft-label draws (``random()`` and ``FloatTensor.uniform_``, losses.py:52-70),
* the dropout seeds (torch's CPU generator, ops.dropout).

:class:`StepTape` keeps them device-resident: while a tape is recording, each such draw is made by a *closure* that is
evaluated immediately (the capture's values) and remembered; the kernel is launched in its ``_dev`` form
(``cgb_const_target_loss_dev`` / ``cgb_dropout_dev``) reading slot ``i`` of the tape's device buffer.  Before every replay
:meth:`StepTape.refresh` re-evaluates the closures IN RECORDING ORDER — the same order, on the same generators, as the eager
step draws them, so a replayed step consumes the host RNG streams exactly like an eager one — and uploads them with one
asynchronous copy from pinned memory on the replay stream.

The optimiser update, the data-parallel all-reduce and the learning-rate bookkeeping stay outside the graph (a handful of
launches on flat buffers): the captured region is ``zero_grad`` + loss + ``backward``, which does not depend on the
extragradient parity of the step, so ONE graph per update kind serves both the extrapolation and the update step.
"""
from __future__ import annotations

import contextlib
from typing import Callable, List, Optional, Sequence

import torch

_CURRENT: List[Optional["StepTape"]] = [None]


def current_tape() -> Optional["StepTape"]:
    """The tape that is recording on this thread's capture, or None (eager execution)."""
    return _CURRENT[0]


class StepTape:
    CAP = 512

    def __init__(self, device, pin: bool = True):
        self.device = device
        self._hf = torch.zeros(self.CAP, dtype=torch.float32)
        self._hi = torch.zeros(self.CAP, dtype=torch.int64)
        if pin and torch.cuda.is_available():
            self._hf, self._hi = self._hf.pin_memory(), self._hi.pin_memory()
        self.df = torch.zeros(self.CAP, dtype=torch.float32, device=device)
        self.di = torch.zeros(self.CAP, dtype=torch.int64, device=device)
        self._actions = []          # (kind, first slot, count, closure) in recording order
        self._nf = self._ni = 0
        self.sealed = False

    # -- recording ---------------------------------------------------------------------------------
    def _record(self, kind: str, fn: Callable[[], Sequence], n: int):
        if self.sealed:
            raise RuntimeError("StepTape: a draw was requested after the capture ended")
        host, dev, first = (self._hf, self.df, self._nf) if kind == "f" else (self._hi, self.di, self._ni)
        if first + n > self.CAP:
            raise RuntimeError(f"StepTape: more than {self.CAP} host draws in one captured step")
        vals = list(fn())
        assert len(vals) == n, (len(vals), n)
        for j, v in enumerate(vals):
            host[first + j] = v
        self._actions.append((kind, first, n, fn))
        if kind == "f":
            self._nf += n
        else:
            self._ni += n
        return [dev[first + j: first + j + 1] for j in range(n)]

    def floats(self, fn: Callable[[], Sequence[float]], n: int):
        """n fp32 device slots filled by ``fn()`` (a list of n floats) now and before every replay."""
        return self._record("f", fn, n)

    def ints(self, fn: Callable[[], Sequence[int]], n: int):
        """n int64 device slots filled by ``fn()`` now and before every replay."""
        return self._record("i", fn, n)

    def upload(self):
        """Copy the host values to the device buffers (asynchronous on the current stream; pinned source)."""
        if self._nf:
            self.df[: self._nf].copy_(self._hf[: self._nf], non_blocking=True)
        if self._ni:
            self.di[: self._ni].copy_(self._hi[: self._ni], non_blocking=True)

    def refresh(self):
        """Re-draw every recorded value in recording order and upload — call before each replay."""
        for kind, first, n, fn in self._actions:
            vals = list(fn())
            host = self._hf if kind == "f" else self._hi
            for j in range(n):
                host[first + j] = vals[j]
        self.upload()

    @property
    def n_draws(self):
        return self._nf + self._ni


@contextlib.contextmanager
def recording(tape: StepTape):
    prev = _CURRENT[0]
    _CURRENT[0] = tape
    try:
        yield tape
    finally:
        _CURRENT[0] = prev
        tape.sealed = True


class GraphedStep:
    """One captured region with static inputs.  ``fn(static_inputs) -> outputs`` is captured on the first :meth:`__call__`
    (after the caller's eager warm-up); later calls copy the new inputs into the static buffers, refresh the tape and replay."""

    def __init__(self, fn, example_inputs, device, pool=None):
        self.fn = fn
        self.device = device
        self.static_in = _clone_tree(example_inputs)
        self.graph = None
        self.tape = None
        self.outputs = None
        self.pool = pool
        self.replays = 0

    def capture(self):
        from . import ops

        self.tape = StepTape(self.device)
        g = torch.cuda.CUDAGraph()
        ops.invalidate_weight_cache()   # every weight packing must be re-issued inside the capture
        ops._BEPOCH[0] += 1
        torch.cuda.synchronize()
        with recording(self.tape):
            with torch.cuda.graph(g, pool=self.pool):
                self.outputs = self.fn(self.static_in)
        self.graph = g
        self.tape.upload()
        return self

    def __call__(self, inputs):
        from . import ops

        _copy_tree(self.static_in, inputs)
        if self.graph is None:
            self.capture()
        else:
            self.tape.refresh()
        self.graph.replay()
        self.replays += 1
        # packed-weight / folded-BN caches filled during the capture hold graph-pool tensors whose contents the replay has
        # just rewritten from the CURRENT weights; anything cached by eager code before the replay is stale
        ops.invalidate_weight_cache()
        ops._BEPOCH[0] += 1
        return self.outputs


def _clone_tree(t):
    if isinstance(t, torch.Tensor):
        return t.detach().clone()
    if isinstance(t, dict):
        return {k: _clone_tree(v) for k, v in t.items()}
    if isinstance(t, (list, tuple)):
        return type(t)(_clone_tree(v) for v in t)
    return t


def _copy_tree(dst, src):
    if isinstance(dst, torch.Tensor):
        if dst.data_ptr() != src.data_ptr():
            dst.copy_(src, non_blocking=True)
        return
    if isinstance(dst, dict):
        for k in dst:
            _copy_tree(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for a, b in zip(dst, src):
            _copy_tree(a, b)


def tree_signature(t):
    """Hashable (shape, dtype) structure of a tensor tree — a graph is reused only for identical signatures."""
    if isinstance(t, torch.Tensor):
        return (tuple(t.shape), str(t.dtype))
    if isinstance(t, dict):
        return tuple((k, tree_signature(v)) for k, v in sorted(t.items(), key=lambda kv: str(kv[0])))
    if isinstance(t, (list, tuple)):
        return tuple(tree_signature(v) for v in t)
    return repr(t)
