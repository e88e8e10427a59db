"""Validation metrics of ``Trainer.eval_images`` on the device (``climategan/eval_metrics.py:68-130``).

The reference moves every prediction to the host and loops over classes in Python (two masked reductions and three ``.item()``
per class).  Both metrics are integer functions of the confusion matrix of ``argmax(pred, 1)`` against the label, which
``cgb_argmax_confusion`` builds in one pass over the logits; one ``(C (C+1) + 1)``-element int64 copy then brings it to the host.
Labels are integer-valued (segmentation ids, depth buckets, the binarised mask of ``data.py:391-397``)."""
import numpy as np
import torch

from . import ops


def confusion(pred, label):
    """Host copy of ``ops.argmax_confusion``: (conf int64 ndarray [C, C+1], int label.max())."""
    conf, lmax = ops.argmax_confusion(pred, label)
    host = torch.cat([conf.reshape(-1), lmax.reshape(1)]).cpu().numpy()   # the only device->host transfer of a metric
    c = pred.shape[1]
    return host[:-1].reshape(c, c + 1), int(host[-1])


def accuracy_from_confusion(conf):
    return float(np.trace(conf[:, :-1])) / float(conf.sum())


def miou_from_confusion(conf, label_max, average="macro"):
    """eval_metrics.py:97-124: IoU of every class that is predicted or present; with <= 2 classes only class label.max()."""
    num_classes = conf.shape[0]
    classes = list(range(num_classes)) if num_classes > 2 else [label_max]
    weights, ious = [], []
    for k in classes:
        n_pred = int(conf[k].sum()) if 0 <= k < num_classes else 0
        n_target = int(conf[:, k].sum()) if 0 <= k < num_classes else 0
        if n_pred > 0 or n_target > 0:
            inter = int(conf[k, k])
            weights.append(n_pred)
            ious.append(float(inter) / float(n_pred + n_target - inter))
    if not ious:
        return float("nan")
    if average == "weighted":
        return np.sum(np.multiply(weights, ious) / np.sum(weights))
    return np.mean(ious)


def accuracy(pred_im, gt_im):
    """eval_metrics.py:68-77 for the forms ``eval_images`` uses: pred [N,C,H,W] against gt [N,1,H,W] or [N,H,W] — the argmax over
    the class axis is compared with the label (a one-channel prediction therefore scores the fraction of zero labels, as in the
    reference, where ``len(pred.shape) > len(gt_im.shape)`` holds after the label's channel axis is dropped)."""
    if gt_im.dim() == 4:
        assert gt_im.shape[1] == 1
    if pred_im.dim() != 4 or gt_im.dim() not in (3, 4):
        raise NotImplementedError("accuracy: prediction [N,C,H,W] against label [N,1,H,W] / [N,H,W] (the forms eval_images uses)")
    conf, _ = confusion(pred_im, gt_im)
    return accuracy_from_confusion(conf)


def mIOU(pred, label, average="macro"):
    """eval_metrics.py:80-124: pred [N,C,H,W] logits, label integer ids of N*H*W pixels."""
    conf, lmax = confusion(pred, label)
    return miou_from_confusion(conf, lmax, average)


def accuracy_and_mIOU(pred, label, average="macro"):
    """Both metrics from ONE pass over the logits (eval_images asks for both on every prediction)."""
    conf, lmax = confusion(pred, label)
    return accuracy_from_confusion(conf), miou_from_confusion(conf, lmax, average)
