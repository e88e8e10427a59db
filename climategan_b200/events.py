"""Climate events on the device — the compositing half of ``Trainer.infer_all`` (climategan/trainer.py:218-334):
``add_fire`` (climategan/fire.py:68-127), ``compute_smog`` (trainer.py:1879-1939), ``normalize`` -> uint8 NHWC output
(trainer.py:312-327).  NCHW fp32 tensors in and out, like the reference; every array op is a libcgb200 kernel."""
from __future__ import annotations

import ctypes as C
import random

import torch

from . import _lib, ops
from ._lib import check


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _L():
    return _lib.lib()


def _img(x):
    _lib.require_device()
    if not ops._on_device(x):
        raise _lib.CgbError("climategan_b200 tensors must live on a CUDA device (no CPU path)")
    return x.detach().contiguous().float()


def minmax_per_sample(x):
    """Per-sample (min, max) over all other dimensions -> fp32 [N, 2] (tutils.normalize :567-576)."""
    x = _img(x)
    n = x.shape[0]
    mm = torch.empty((n, 2), dtype=torch.float32, device=x.device)
    check(_L().cgb_minmax_per_sample(_p(x), _p(mm), n, x.numel() // n, _st()), "minmax_per_sample")
    return mm


def add_fire(x, seg_preds, fire_opts, green=None):
    """fire.add_fire(x, seg_preds, opts.events.fire): x NCHW in [-1,1], seg_preds logits [N,C,hs,ws] -> float image in [0,255].
    ``green``: the filter's G value; the reference draws random.randint(100, 150) per call (fire.py:115) — same draw here."""
    x, seg = _img(x), _img(seg_preds)
    n, c, h, w = x.shape
    assert c == 3
    hw = h * w
    mm = minmax_per_sample(x)
    toned = torch.empty_like(x)
    gray = torch.empty((n,), dtype=torch.float64, device=x.device)
    check(_L().cgb_fire_tone(_p(x), _p(mm), _p(toned), _p(gray), n, hw, 1.5, 0.73, _st()), "fire_tone")   # fire.py:90-91
    _, cs, hs, ws = seg.shape
    sky_small = torch.empty((n, hs, ws), dtype=torch.float32, device=x.device)
    check(_L().cgb_sky_mask(_p(seg), _p(sky_small), n, cs, hs, ws, 9, 1 if fire_opts.get("crop_bottom_sky_mask") else 0, _st()),
          "sky_mask")
    a = torch.empty((n, h, w), dtype=torch.float32, device=x.device)
    b = torch.empty_like(a)
    t = torch.empty_like(a)
    check(_L().cgb_plane_resize_nearest(_p(sky_small), _p(a), n, hs, ws, h, w, _st()), "plane_resize_nearest")
    n_lines, n_cols = int(0.18 * h), int(0.18 * w)                                                          # fire.py:103
    check(_L().cgb_box_dilate(_p(a), _p(t), _p(b), n, h, w, max(n_cols - 1, 0), max(n_lines - 1, 0), _st()), "box_dilate")
    ksize = int(fire_opts.get("kernel_size", 301) or 301)
    sigma = float(fire_opts.get("kernel_sigma", 150.5) or 150.5)
    check(_L().cgb_gauss_blur(_p(b), _p(t), _p(a), n, h, w, ksize, sigma, _st()), "gauss_blur")
    if green is None:
        green = random.randint(100, 150)
    out = torch.empty_like(x)
    check(_L().cgb_fire_paste(_p(toned), _p(a), _p(out), n, h, w, 255.0, float(green), 0.0, 200.0, 0.8, _st()), "fire_paste")
    return out


def add_smog(x, d, smog_opts):
    """Trainer.compute_smog given the depth prediction d [N,1,hd,wd] (trainer.py:1902-1939)."""
    x, d = _img(x), _img(d)
    n, c, h, w = x.shape
    assert c == 3 and d.shape[0] == n and d.shape[1] == 1
    mmx, mmd = minmax_per_sample(x), minmax_per_sample(d)
    out = torch.empty_like(x)
    yc = smog_opts.yellow_color
    alpha = smog_opts.alpha / 255
    check(_L().cgb_smog(_p(x), _p(mmx), _p(d), _p(mmd), _p(out), n, h, w, d.shape[2], d.shape[3], float(smog_opts.airlight),
                        float(smog_opts.beta / smog_opts.vr), float(alpha), yc[0] / 255, yc[1] / 255, yc[2] / 255, _st()), "smog")
    return out


def resize_and_crop_size(h, w, to):
    """apply_events.py:223-239: the resized shape (short side = ``to``, aspect ratio kept, ``int()`` truncation as there) and
    the centred crop offsets."""
    rh, rw = (to, int(to * w / h)) if h < w else (int(to * h / w), to)
    return rh, rw, (rh - to) // 2, (rw - to) // 2


class InputEdge:
    """uint8 HWC photographs -> one NCHW fp32 batch in [-1, 1] on the device: the pre-processing of ``apply_events.py``
    (:487-502: ``resize_and_crop`` + ``to_m1_p1``, or the ``--keep_ratio_128`` resize) as ONE kernel per image fed from pinned,
    double-buffered host staging (the H2D copy of image i+1 overlaps the kernel of image i on the same stream's copy engine)."""

    def __init__(self, device, quantize=True):
        self.device = device
        self.quantize = 1 if quantize else 0
        self._stage = [None, None]
        self._dev = [None, None]
        self._copied = [None, None]     # event recorded after the H2D copy out of each pinned buffer
        self._k = 0

    def _buffers(self, nbytes):
        k = self._k
        self._k ^= 1
        if self._copied[k] is not None:
            self._copied[k].synchronize()   # the copy that last read this pinned buffer has finished: the host may overwrite it
        if self._stage[k] is None or self._stage[k].numel() < nbytes:
            self._stage[k] = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            self._dev[k] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._stage[k], self._dev[k], k

    def __call__(self, images, target=640, keep_ratio_sizes=None):
        """images: list of uint8 [h, w, 3] numpy arrays / tensors.  ``keep_ratio_sizes``: per-image (H, W) for the
        ``--keep_ratio_128`` mode (plain resize, no crop; all images of a batch must share it); otherwise resize_and_crop to
        ``target``.  Returns fp32 [N, 3, T, T] (or [N, 3, H, W])."""
        import numpy as np

        if keep_ratio_sizes is not None:
            assert len(set(keep_ratio_sizes)) == 1, "a batch needs one output size"
            th, tw = keep_ratio_sizes[0]
        else:
            th = tw = target
        out = torch.empty((len(images), 3, th, tw), dtype=torch.float32, device=self.device)
        for i, im in enumerate(images):
            a = torch.from_numpy(np.ascontiguousarray(im)) if not isinstance(im, torch.Tensor) else im.contiguous()
            assert a.dtype == torch.uint8 and a.dim() == 3 and a.shape[2] == 3, (a.dtype, a.shape)
            h, w = int(a.shape[0]), int(a.shape[1])
            if keep_ratio_sizes is not None:
                rh, rw, top, left = th, tw, 0, 0
            else:
                rh, rw, top, left = resize_and_crop_size(h, w, target)
            host, dev, k = self._buffers(a.numel())
            host[: a.numel()].copy_(a.reshape(-1))
            dev[: a.numel()].copy_(host[: a.numel()], non_blocking=True)
            if torch.cuda.is_available():
                self._copied[k] = torch.cuda.Event()
                self._copied[k].record()
            check(_L().cgb_resize_crop_u8(_p(dev), _p(out[i]), h, w, rh, rw, top, left, th, tw, self.quantize, _st()), "resize_crop_u8")
        return out


def to_uint8_nhwc(t):
    """normalize(t) -> permute(0,2,3,1) -> (t*255).astype(uint8) (trainer.py:312-327) as one device kernel: uint8 [N,H,W,3]."""
    t = _img(t)
    n, c, h, w = t.shape
    assert c == 3
    mm = minmax_per_sample(t)
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=t.device)
    check(_L().cgb_to_uint8_nhwc(_p(t), _p(mm), _p(out), n, h * w, _st()), "to_uint8_nhwc")
    return out


def mask_to_uint8(mask, bin_value):
    """((mask > bin_value) * 255).astype(uint8) (trainer.py:330-332)."""
    m = _img(mask)
    out = torch.empty(m.shape, dtype=torch.uint8, device=m.device)
    check(_L().cgb_mask_to_uint8(_p(m), _p(out), float(bin_value), m.numel(), _st()), "mask_to_uint8")
    return out


def cloudy_input(x, s, sky_idx=9, res=(8, 8), weight=0.8):
    """The intermediary cloudy image of OmniGenerator.paint_cloudy (generator.py:318-325): Perlin noise (tutils.py:648-694)
    mixed into the sky region of x.  The (res+1)^2 gradient angles are drawn with ``torch.rand`` on the CPU generator — the
    same draw the reference makes — and evaluated per pixel on the device."""
    import math

    x, s = _img(x), _img(s)
    n, c3, h, w = x.shape
    assert c3 == 3
    angles = (2 * math.pi * torch.rand(res[0] + 1, res[1] + 1)).to(x.device)
    noise = torch.empty((1, h, w), dtype=torch.float32, device=x.device)
    check(_L().cgb_perlin_noise(_p(angles), _p(noise), h, w, res[0], res[1], _st()), "perlin_noise")
    mm = minmax_per_sample(noise)
    out = torch.empty_like(x)
    check(_L().cgb_cloudy_mix(_p(x), _p(s), _p(noise), _p(mm), _p(out), n, h, w, s.shape[1], s.shape[2], s.shape[3], sky_idx,
                              float(weight), _st()), "cloudy_mix")
    return out
