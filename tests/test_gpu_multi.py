"""Data-parallel correctness on >= 2 real GPUs (skipped on a one-GPU box): scripts/dp_check.py under torchrun — the reduced
gradient is the mean of the per-shard gradients and the replicas stay bit-identical, eager and with the CUDA-graph step.
The log of the run on 2 B200s is committed as profiles/r02_dp_check_2gpu.json."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_data_parallel_gradients_and_replicas(cuda):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29511", os.path.join(ROOT, "scripts", "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, (res.stdout[-2000:], res.stderr[-2000:])
    rep = json.loads(lines[-1])
    assert rep["ok"] and rep["eager"]["replica_drift"] == 0.0 and rep["graphs"]["replica_drift"] == 0.0, rep
