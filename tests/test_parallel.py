"""CPU test of the N>1 path: world_size-2 gloo process group, flat gradient bucket all-reduce == what the reference
wrapped in DDP would produce (mean of the per-shard gradients)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from climategan_b200.parallel import GradBucket, allreduce_flat_grads, shard_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)  # identical initial weights on every rank
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1))
    frozen = torch.nn.Parameter(torch.ones(3), requires_grad=False)
    x = torch.arange(4 * 3 * 5 * 5, dtype=torch.float32).reshape(4, 3, 5, 5) / 100.0
    xs = shard_batch(x, rank, world)
    net(xs).pow(2).mean().backward()
    if rank == 1:
        net[2].bias.grad = None  # a rank with a missing grad still takes part
    bucket = GradBucket(list(net.parameters()) + [frozen])
    assert bucket.numel == sum(p.numel() for p in net.parameters())
    bucket.allreduce()
    out[rank] = [p.grad.clone() for p in net.parameters()]
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # reference: per-shard gradients averaged
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1))
    x = torch.arange(4 * 3 * 5 * 5, dtype=torch.float32).reshape(4, 3, 5, 5) / 100.0
    grads = []
    for r in range(world):
        net.zero_grad()
        net(x[r * 2:(r + 1) * 2]).pow(2).mean().backward()
        g = [p.grad.clone() for p in net.parameters()]
        if r == 1:
            g[3] = torch.zeros_like(g[3])
        grads.append(g)
    mean = [(a + b) / 2 for a, b in zip(*grads)]
    for r in range(world):
        for got, want in zip(out[r], mean):
            assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
    # both ranks hold identical gradients after the collective
    for a, b in zip(out[0], out[1]):
        assert torch.equal(a, b)


def test_shard_batch():
    x = torch.arange(12).reshape(6, 2)
    assert torch.equal(shard_batch(x, 1, 3), x[2:4])
    try:
        shard_batch(x, 0, 4)
    except ValueError:
        pass
    else:
        raise AssertionError("uneven shard must raise")


class _FlatOpt:
    """Stand-in for optim.ExtraAdam's flat-buffer surface (the optimiser itself needs the CUDA library)."""

    def __init__(self, *flats):
        self.flat_grads = list(flats)


def _worker_flat(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    d = torch.ones(4) * (10 * rank)
    nbytes = allreduce_flat_grads(_FlatOpt(g, d))
    out[rank] = (g.clone(), d.clone(), nbytes)
    dist.destroy_process_group()


def test_allreduce_flat_grads_gloo():
    """Trainer.enable_data_parallel's collective: the optimiser's flat G / D gradient buffers are averaged in place, one
    all-reduce per parameter group (what DDP would give the reference step)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_flat, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        g, d, nbytes = out[r]
        assert torch.allclose(g, torch.arange(10, dtype=torch.float32) * 1.5)
        assert torch.allclose(d, torch.ones(4) * 5.0)
        assert nbytes == 14 * 4
    assert allreduce_flat_grads(_FlatOpt(torch.ones(3))) == 0   # not distributed: a no-op


def _worker_trainer(rank, world, port, out):
    """The REAL Trainer / ExtraAdam / enable_data_parallel on each rank, kernels replaced by the type-checking no-op library
    (tests/dryrun.py): what is exercised is the N>1 host path — identical seeded weights, per-rank batch slices, one flat
    gradient all-reduce after each backward, the same number of collectives on every rank (no deadlock)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts, synth_batch
    from tests.dryrun import noop_library

    torch.manual_seed(0)
    opts = full_opts(size=128)
    n_coll = [0]
    real_all_reduce = dist.all_reduce

    def counting_all_reduce(*a, **k):
        n_coll[0] += 1
        return real_all_reduce(*a, **k)

    dist.all_reduce = counting_all_reduce
    with noop_library():
        t = Trainer(opts, device=torch.device("cpu")).setup(input_shape=(128, 128))
        t.enable_data_parallel()
        w0 = t.G.painter.conv_img.weight.detach().clone()
        mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, 2, 128, seed=100 + rank).items()}
        for _ in range(2):
            # stand-in gradients that differ per rank: the no-op kernels leave .grad untouched, so write them by hand between
            # the backward and the collective by wrapping _sync_grads
            sync = t._sync_grads

            def sync_with_fake_grads(opt, _sync=sync):
                for f in opt.flat_grads:
                    f.copy_(torch.arange(f.numel(), dtype=torch.float32) % 7 + rank)
                _sync(opt)

            t._sync_grads = sync_with_fake_grads
            t.update_G(mdb)
            t.update_D(mdb)
            t._sync_grads = sync
            t.logger.global_step += 1
        out[rank] = dict(w0=w0, g=[f.clone() for f in t.g_opt.flat_grads], d=[f.clone() for f in t.d_opt.flat_grads],
                         n_coll=n_coll[0], x=mdb["r"]["data"]["x"][0, 0, 0, :4].clone())
    dist.destroy_process_group()


def test_trainer_data_parallel_host_path_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_trainer, args=(world, _free_port(), out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert torch.equal(a["w0"], b["w0"])                       # identical initial weights on every rank
    assert not torch.equal(a["x"], b["x"])                     # each rank its own batch slice
    assert a["n_coll"] == b["n_coll"] and a["n_coll"] >= 4     # one all-reduce per parameter group after each backward
    for fa, fb in zip(a["g"] + a["d"], b["g"] + b["d"]):
        want = torch.arange(fa.numel(), dtype=torch.float32) % 7 + 0.5     # mean of rank 0's and rank 1's stand-in gradients
        assert torch.equal(fa, fb) and torch.equal(fa, want)
