"""Host logic of the CUDA-graph step (climategan_b200/graphs.py), on CPU: the StepTape re-makes the host draws of a captured
step in recording order, so a replayed step consumes the RNG streams exactly like an eager one."""
import random

import torch

from climategan_b200 import graphs, ops
from climategan_b200.losses import GANLoss


def test_step_tape_slots_and_refresh():
    tape = graphs.StepTape(torch.device("cpu"), pin=False)
    state = {"k": 0}

    def draw_f():
        state["k"] += 1
        return [state["k"] + 0.5, state["k"] + 0.25]

    def draw_i():
        state["k"] += 1
        return [2 ** 40 + state["k"]]

    with graphs.recording(tape):
        assert graphs.current_tape() is tape
        f = tape.floats(draw_f, 2)
        i = tape.ints(draw_i, 1)
    assert graphs.current_tape() is None and tape.sealed and tape.n_draws == 3
    tape.upload()
    assert [float(t) for t in f] == [1.5, 1.25] and int(i[0]) == 2 ** 40 + 2
    tape.refresh()   # same closures, same order
    assert [float(t) for t in f] == [3.5, 3.25] and int(i[0]) == 2 ** 40 + 4
    assert f[0].data_ptr() == tape.df.data_ptr()   # slots are views of the one device buffer the kernels read


def _targets_of(monkeypatch, fn):
    seen = []
    monkeypatch.setattr(ops, "const_target_loss", lambda x, kind, t: seen.append(float(t)) or torch.zeros(()))
    fn()
    return seen


def test_ganloss_replay_draws_match_eager(monkeypatch):
    """Two eager GANLoss steps and (one recorded step + one refresh) consume random() / the torch generator identically."""
    loss = GANLoss(use_lsgan=False, soft_shift=0.2, flip_prob=0.5)
    preds = [[torch.zeros(2, 1, 3, 3)] * 2, [torch.zeros(2, 1, 2, 2)] * 2, [torch.zeros(2, 1, 1, 1)] * 2]

    def one_step():
        loss(preds, True, False)
        loss(preds[0][0], False, True)

    random.seed(7)
    torch.manual_seed(7)
    eager = [_targets_of(monkeypatch, one_step), _targets_of(monkeypatch, one_step)]

    random.seed(7)
    torch.manual_seed(7)
    tape = graphs.StepTape(torch.device("cpu"), pin=False)
    with graphs.recording(tape):
        rec = _targets_of(monkeypatch, lambda: (one_step(), tape.upload()))
    first = tape.df[: tape.n_draws].tolist()
    tape.refresh()
    second = tape.df[: tape.n_draws].tolist()
    assert len(first) == 4 and tape.n_draws == 4
    assert first == [torch.tensor(v, dtype=torch.float32).item() for v in eager[0]]
    assert second == [torch.tensor(v, dtype=torch.float32).item() for v in eager[1]]
    assert eager[0] != eager[1]


def test_dropout_seed_goes_through_the_tape(monkeypatch):
    calls = []
    monkeypatch.setattr(ops._Dropout, "apply", staticmethod(lambda x, p, seed: calls.append(seed) or x))
    x = torch.zeros(1, 2, 2, 8)
    torch.manual_seed(3)
    ops.dropout(x, 0.5, True)
    eager_seed = calls[-1]
    assert isinstance(eager_seed, int)
    torch.manual_seed(3)
    tape = graphs.StepTape(torch.device("cpu"), pin=False)
    with graphs.recording(tape):
        ops.dropout(x, 0.5, True)
    tape.upload()
    assert isinstance(calls[-1], torch.Tensor) and int(calls[-1][0]) == eager_seed


def test_tree_signature_distinguishes_shapes():
    a = {"r": {"x": torch.zeros(2, 3, 4, 4)}, "s": {"x": torch.zeros(2, 3, 4, 4), "m": torch.zeros(2, 1, 4, 4)}}
    b = {"r": {"x": torch.zeros(4, 3, 4, 4)}, "s": {"x": torch.zeros(2, 3, 4, 4), "m": torch.zeros(2, 1, 4, 4)}}
    assert graphs.tree_signature(a) != graphs.tree_signature(b)
    assert graphs.tree_signature(a) == graphs.tree_signature({k: dict(v) for k, v in a.items()})
