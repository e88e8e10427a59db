"""Host-side profile of the full train step against the no-op library (tests/dryrun.py): the bench architecture (ResNet-101
masker, 640-channel / 7-level painter, 3-scale discriminator) at 256x256, so the launch count per step equals the real
workload's while no kernel runs — what is left is the Python / autograd / ctypes cost of enqueueing a step.
usage: PYTHONPATH=. python scripts/host_profile_dryrun.py [--profile]"""
import cProfile
import pstats
import sys
import time

import torch

from climategan_b200.trainer import Trainer
from climategan_b200.utils import full_opts, synth_batch
from tests.dryrun import noop_library

S, B = 256, 2
opts = full_opts(nblocks=(3, 4, 23, 3), size=S, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3)
opts.dis.soft_shift, opts.dis.flip_prob = 0.2, 0.05
torch.set_num_threads(1)
with noop_library() as lib:
    t = Trainer(opts, device=torch.device("cpu"), storage_dtype=torch.bfloat16).setup(input_shape=(S, S))
    mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, B, S, 3).items()}

    def step():
        t.update_G(mdb)
        t.update_D(mdb)
        t.logger.global_step += 1

    for _ in range(2):
        step()
    n0 = sum(lib.calls.values())
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    dt = (time.perf_counter() - t0) / 3
    print(f"host time per step {dt * 1e3:.1f} ms, library calls per step {(sum(lib.calls.values()) - n0) // 3}")
    if "--profile" in sys.argv:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(2):
            step()
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(35)
