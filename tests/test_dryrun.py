"""Host logic of the train step on CPU: every configuration the GPU tests pin numerically is also driven end to end against
a stand-in for libcgb200 that type-checks each argument against the declared C signature and computes nothing
(tests/dryrun.py).  Catches what needs no GPU to catch: wrong argument counts / types at the C boundary, layout or shape
mistakes between layers (the storage-tensor checks stay in force), missing autograd edges, parameters that never receive a
gradient, and a call sequence that is not the same from one step to the next."""
import json
import os

import pytest
import torch

from climategan_b200.trainer import Trainer
from climategan_b200.utils import full_opts, synth_batch
from tests.dryrun import noop_library
from tests.helpers import GOLDEN

CASES = ["full_step", "full_step_pl4m", "masker_step_spade", "masker_step_base_depth_classify", "masker_step_v3",
         "mask_only_step_v3"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_train_step_host_path_runs_against_the_noop_library(case, dtype):
    meta = json.load(open(os.path.join(GOLDEN, case + ".json")))
    size, batch = meta["size"], meta["batch"]
    opts = full_opts(size=size, tasks=tuple(meta.get("tasks", ("d", "s", "m", "p"))), use_spade=meta.get("use_spade", False),
                     overrides=meta.get("overrides"))
    with noop_library() as lib:
        t = Trainer(opts, device=torch.device("cpu"), storage_dtype=dtype).setup(input_shape=(size, size))
        t.use_pl4m = bool(meta.get("pl4m", False))
        mdb = {dom: t.batch_to_device(b) for dom, b in synth_batch(opts, batch, size, meta["seeds"]["inputs"]).items()}
        per_step = []
        for _ in range(3):
            before = dict(lib.calls)
            t.update_G(mdb)
            if _ == 0:
                # every parameter the reference gives a gradient to gets one here (the golden stores -1 where it does not)
                import numpy as np

                ref = np.load(os.path.join(GOLDEN, case + ".npz"))["G.gradnorm"]
                names = meta["g_param_names"]
                gp = dict(t.G.named_parameters())
                missing = [n for n, r in zip(names, ref) if r >= 0 and gp[n].requires_grad and gp[n].grad is None]
                assert not missing, missing[:10]
            t.update_D(mdb)
            t.logger.global_step += 1
            per_step.append({k: v - before.get(k, 0) for k, v in lib.calls.items() if v != before.get(k, 0)})
        # extrapolation step and update step enqueue the same work; so does every later step
        assert per_step[1] == per_step[2], {k: (per_step[1].get(k), per_step[2].get(k)) for k in set(per_step[1]) | set(per_step[2])
                                            if per_step[1].get(k) != per_step[2].get(k)}
        assert per_step[1]["cgb_conv2d_fwd"] > 50 and per_step[1]["cgb_extra_adam"] >= 2
        logs = t.losses_to_host()
        assert "total_loss" in logs["gen"] and "total_loss" in logs["disc"]
        assert set(k for k in meta["logs"][0] if k.startswith("gen.task")) <= set(_flat(logs))


def _flat(d, prefix=""):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flat(v, prefix + k + "."))
        else:
            out[prefix + k] = v
    return out
