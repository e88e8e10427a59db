"""Seeded (prediction, label) pairs for the validation-metric fixtures (tests/golden/eval_metrics.json).  numpy RandomState streams
are stable across versions, so the fixture stores the seeds' RESULTS only; inputs are re-drawn here by the generator and by every
test.  Shapes follow Trainer.eval_images (trainer.py:1706-1799): one image, label [1,1,H,W]."""
import numpy as np
import torch


def _seg(seed, c, h, w, n=1, label_hi=None, dead=(), quantise=None, ignore=None):
    rs = np.random.RandomState(seed)
    pred = rs.standard_normal((n, c, h, w)).astype(np.float32)
    label = rs.randint(0, label_hi or c, size=(n, 1, h, w)).astype(np.int64)
    for k in dead:               # classes that are never predicted
        pred[:, k] -= 100.0
    if quantise:                 # exact ties between classes: the first maximum wins (torch.argmax)
        pred = np.round(pred / quantise) * quantise
    if ignore is not None:       # ignore index outside [0, c)
        label[rs.random_sample(label.shape) < 0.1] = ignore
    return torch.from_numpy(pred.astype(np.float32)), torch.from_numpy(label)


def _mask(seed, h, w, empty=False):
    rs = np.random.RandomState(seed)
    pm = (rs.random_sample((1, 1, h, w)) > 0.5).astype(np.float32)
    m = np.zeros((1, 1, h, w), np.float32) if empty else (rs.random_sample((1, 1, h, w)) > 0.6).astype(np.float32)
    return torch.from_numpy(pm), torch.from_numpy(m)


def cases():
    """name -> (pred, label, kind); kind "seg": accuracy + mIOU on the same pair, "mask": accuracy on the one-channel binarised
    mask, mIOU on cat[1 - mask, mask] (trainer.py:1766-1777)."""
    out = {
        "seg11": (*_seg(1, 11, 40, 40), "seg"),
        "seg11_ragged": (*_seg(2, 11, 33, 47), "seg"),
        "seg11_absent_classes": (*_seg(3, 11, 40, 40, label_hi=5, dead=(7, 8, 9, 10)), "seg"),
        "seg11_ties": (*_seg(4, 11, 40, 40, quantise=1.0), "seg"),
        "seg11_ignore255": (*_seg(5, 11, 40, 40, ignore=255), "seg"),
        "depth16_buckets": (*_seg(6, 16, 32, 32), "seg"),
        "seg3": (*_seg(7, 3, 24, 24), "seg"),
        "mask": (*_mask(8, 64, 64), "mask"),
        "mask_empty_label": (*_mask(9, 64, 64, empty=True), "mask"),
    }
    return out
