"""The loader -> device edge (climategan_b200/data.py): per-rank sharding and the double-buffered prefetcher, on CPU."""
import torch
from torch.utils.data import TensorDataset

from climategan_b200.data import DevicePrefetcher, ShardSampler, get_loader, shard_indices
from climategan_b200.utils import Dict


def test_shards_partition_the_epoch_and_reshuffle_per_epoch():
    n, world = 103, 4
    shards = [shard_indices(n, r, world, seed=7, epoch=3) for r in range(world)]
    assert all(len(s) == n // world for s in shards)
    flat = sorted(i for s in shards for i in s)
    assert len(set(flat)) == len(flat) == n // world * world and set(flat) <= set(range(n))
    assert shards != [shard_indices(n, r, world, seed=7, epoch=4) for r in range(world)]          # a new permutation per epoch
    assert shards == [shard_indices(n, r, world, seed=7, epoch=3) for r in range(world)]          # same on every rank / call
    padded = [shard_indices(n, r, world, seed=7, epoch=0, drop_last=False) for r in range(world)]
    assert all(len(s) == 26 for s in padded) and set(i for s in padded for i in s) == set(range(n))
    assert shard_indices(10, 1, 2, shuffle=False) == [1, 3, 5, 7, 9]


def test_get_loader_uses_the_reference_arguments_and_the_shard():
    ds = TensorDataset(torch.arange(40).float().view(40, 1))
    opts = Dict({"data": {"loaders": {"batch_size": 4, "num_workers": 0}}})
    seen = []
    for rank in range(2):
        dl = get_loader(ds, opts, rank=rank, world=2, seed=1)
        assert dl.batch_size == 4 and dl.drop_last and isinstance(dl.sampler, ShardSampler) and len(dl) == 5
        dl.sampler.set_epoch(2)
        seen.append(torch.cat([b[0].flatten() for b in dl]).tolist())
    assert not set(seen[0]) & set(seen[1]) and len(seen[0]) == len(seen[1]) == 20


def test_prefetcher_preserves_the_multi_batch_structure_and_order():
    def batches():
        for i in range(5):
            yield ({"data": {"x": torch.full((2, 3, 4, 4), float(i)), "m": torch.zeros(2, 1, 4, 4)}, "domain": ["r", "r"], "paths": {"x": ["a", "b"]}},
                   {"data": {"x": torch.full((2, 3, 4, 4), float(-i))}, "domain": ["rf", "rf"], "paths": {}})

    out = list(DevicePrefetcher(batches(), "cpu", depth=2))
    assert len(out) == 5
    for i, tup in enumerate(out):
        assert isinstance(tup, tuple) and len(tup) == 2
        assert float(tup[0]["data"]["x"][0, 0, 0, 0]) == i and float(tup[1]["data"]["x"][0, 0, 0, 0]) == -i
        assert tup[0]["domain"] == ["r", "r"] and tup[0]["paths"] == {"x": ["a", "b"]}
