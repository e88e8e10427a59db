"""The loader -> device edge of the train step (``climategan/data.py:506-539`` ``get_loader`` /
``get_all_loaders``, ``trainer.batch_to_device`` :609-621, ``trainer.train_loaders`` :626-634).

The reference builds one ``DataLoader(shuffle=True, pin_memory=True, drop_last=True)`` per domain, zips them, and moves every
tensor of a batch to the GPU with a blocking ``tensor.to(device)`` at the top of the step.  What is built here is that edge
for the B200 data-parallel step — the datasets themselves (file lists, image decoding, augmentation: ``OmniListDataset``,
``transforms.py``) are out of the hot path's scope and are whatever ``torch.utils.data.Dataset`` the caller has:

* :func:`shard_indices` / :class:`ShardSampler` — the per-rank slice of a shuffled epoch (every rank draws the SAME
  permutation from ``seed + epoch`` and takes its interleaved share, padded or truncated to equal length): the
  ``DistributedSampler``-equivalent the reference never needed (it has no distributed code).
* :func:`get_loader` — the reference's ``DataLoader`` arguments (``batch_size`` = ``opts.data.loaders.batch_size`` PER RANK,
  ``pin_memory``, ``drop_last``) with that sampler.
* :class:`DevicePrefetcher` — wraps the zip of the per-domain loaders: batch i+1 is copied host -> device from pinned memory on a
  side stream while step i computes, two batches deep; ``__next__`` hands out device-resident ``multi_batch_tuple``s in the
  exact structure ``Trainer.run_epoch`` consumes and makes the compute stream wait on the copy's event (no host sync).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional

import torch
from torch.utils.data import DataLoader, Sampler


def shard_indices(n: int, rank: int, world: int, seed: int = 0, epoch: int = 0, shuffle: bool = True,
                  drop_last: bool = True) -> List[int]:
    """Indices of ``range(n)`` that rank ``rank`` of ``world`` visits in ``epoch``: one global permutation per (seed, epoch),
    identical on every rank, dealt round-robin.  ``drop_last``: truncate to a multiple of ``world`` (every rank gets
    ``n // world``); otherwise wrap around to ``ceil(n / world)`` each."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    if shuffle:
        g = torch.Generator().manual_seed(int(seed) + int(epoch))
        order = torch.randperm(n, generator=g).tolist()
    else:
        order = list(range(n))
    if drop_last:
        order = order[: n // world * world]
    elif n % world:
        order = order + order[: world - n % world]
    return order[rank::world]


class ShardSampler(Sampler):
    def __init__(self, n: int, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True, drop_last: bool = True):
        self.n, self.rank, self.world, self.seed, self.shuffle, self.drop_last = n, rank, world, seed, shuffle, drop_last
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def __iter__(self) -> Iterator[int]:
        return iter(shard_indices(self.n, self.rank, self.world, self.seed, self.epoch, self.shuffle, self.drop_last))

    def __len__(self) -> int:
        return self.n // self.world if self.drop_last else -(-self.n // self.world)


def get_loader(dataset, opts, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True) -> DataLoader:
    """data.py:506-528 for any map-style dataset: batch_size per rank, pinned host memory, drop_last, the rank's shard."""
    loaders = opts.data.loaders if "data" in opts and "loaders" in opts.data else {}
    return DataLoader(dataset, batch_size=int(loaders.get("batch_size", 4)),
                      sampler=ShardSampler(len(dataset), rank, world, seed, shuffle, drop_last=True),
                      num_workers=int(loaders.get("num_workers", 8)), pin_memory=torch.cuda.is_available(), drop_last=True)


def _to_device_tree(b, device, non_blocking):
    if isinstance(b, torch.Tensor):
        if not b.is_cuda and non_blocking and torch.cuda.is_available() and not b.is_pinned():
            b = b.pin_memory()          # a loader built without pin_memory: stage through pinned memory so the copy is asynchronous
        return b.to(device, non_blocking=non_blocking)
    if isinstance(b, dict):
        return {k: _to_device_tree(v, device, non_blocking) for k, v in b.items()}
    if isinstance(b, (list, tuple)) and b and isinstance(b[0], (torch.Tensor, dict)):
        return type(b)(_to_device_tree(v, device, non_blocking) for v in b)
    return b   # paths, domain names, modes


class DevicePrefetcher:
    """``for multi_batch_tuple in DevicePrefetcher(zip(*loaders), device): trainer.update_G(...)`` — double-buffered H2D."""

    def __init__(self, batches: Iterable, device, depth: int = 2):
        self.device = torch.device(device)
        self.depth = max(1, int(depth))
        self._it = iter(batches)
        self._cuda = self.device.type == "cuda"
        self._stream = torch.cuda.Stream(self.device) if self._cuda else None
        self._queue = []
        for _ in range(self.depth):
            self._fill()

    def _fill(self):
        try:
            b = next(self._it)
        except StopIteration:
            return
        if not self._cuda:
            self._queue.append((_to_device_tree(b, self.device, False), None))
            return
        with torch.cuda.stream(self._stream):
            d = _to_device_tree(b, self.device, True)
            ev = torch.cuda.Event()
            ev.record(self._stream)
        self._queue.append((d, ev))

    def __iter__(self):
        return self

    def __next__(self):
        if not self._queue:
            raise StopIteration
        d, ev = self._queue.pop(0)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)     # device-side dependency only
            _record_stream_tree(d, torch.cuda.current_stream(self.device))
        self._fill()
        return d


def _record_stream_tree(b, stream):
    if isinstance(b, torch.Tensor):
        if b.is_cuda:
            b.record_stream(stream)    # allocated on the copy stream, consumed on the compute stream
    elif isinstance(b, dict):
        for v in b.values():
            _record_stream_tree(v, stream)
    elif isinstance(b, (list, tuple)):
        for v in b:
            _record_stream_tree(v, stream)
