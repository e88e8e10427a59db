"""Data parallelism for the hot path: one process per GPU, the per-image batch split across ranks, and ONE
collective per optimiser step — an all-reduce (mean) of the flat gradient bucket of G (and of D) over NCCL /
NVLink.  The reference has no distributed code at all; this is what wrapping its step in DDP would
compute: every rank runs the identical step on its shard, gradients are averaged, optimiser state is replicated.

The collective itself is torch.distributed's (NCCL on GPUs, gloo in the CPU tests): a pure bandwidth op with no math
fused around it, so it stays a library call.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Rank r takes x[r*B/n:(r+1)*B/n] of every domain batch."""
    b = t.shape[0]
    if b % world:
        raise ValueError(f"batch {b} is not divisible by world size {world}")
    per = b // world
    return t[rank * per:(rank + 1) * per]


def _allreduce_mean(flat: torch.Tensor, group, world: int) -> None:
    """Mean over the ranks in ONE pass: NCCL averages inside the collective (ReduceOp.AVG); gloo (the CPU tests) has no AVG, so
    it sums and scales."""
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)


class GradBucket:
    """Flat fp32 gradient bucket over a fixed parameter list.

    ``allreduce()`` packs every ``p.grad`` into one contiguous buffer, averages it over the process group with a
    single all-reduce and scatters the result back into the ``.grad`` tensors.  Parameters whose grad is None
    contribute zeros (and receive the mean of the other ranks' gradients)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one trainable parameter")
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.numel * 4

    def allreduce(self) -> None:
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        _allreduce_mean(self.flat, self.group, world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def allreduce_flat_grads(optimizer, group: Optional[dist.ProcessGroup] = None) -> int:
    """Average the optimiser's flat fp32 gradient buffers (``ExtraAdam.flat_grads``: parameters and ``.grad``s are views of one
    buffer per parameter group, so there is no pack / unpack copy) over the process group: one all-reduce per group.
    Returns the number of bytes reduced (0 when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    total = 0
    for flat in optimizer.flat_grads:
        _allreduce_mean(flat, group, world)
        total += flat.numel() * 4
    return total
