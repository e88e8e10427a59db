"""Generate tests/golden/eval_metrics.json by calling the UNMODIFIED reference functions climategan.eval_metrics.accuracy / mIOU
(imported from /root/reference through oracle/refshim.py) on the seeded pairs of tests/golden/eval_cases.py.  Build container only:

    python tests/golden/make_eval_golden.py
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refshim  # noqa: E402
from tests.golden.eval_cases import cases  # noqa: E402


def main():
    ref = refshim.load("eval_metrics")
    out = {}
    for name, (pred, label, kind) in cases().items():
        if kind == "mask":
            prob = torch.cat([1 - pred, pred], dim=1)
            out[name] = {"accuracy": float(ref.accuracy(pred, label)), "mIOU": float(ref.mIOU(prob, label)),
                         "mIOU_weighted": float(ref.mIOU(prob, label, average="weighted"))}
        else:
            out[name] = {"accuracy": float(ref.accuracy(pred, label)), "mIOU": float(ref.mIOU(pred, label)),
                         "mIOU_weighted": float(ref.mIOU(pred, label, average="weighted"))}
    with open(os.path.join(HERE, "eval_metrics.json"), "w") as f:
        json.dump({"source": "climategan/eval_metrics.py:68-124 run on tests/golden/eval_cases.py", "cases": out}, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
