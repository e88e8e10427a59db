"""Autograd operators over libcgb200 (the C ABI in include/cgb200.h).

Tensors between operators are *storage tensors*: contiguous ``[N, H, W, Cs]`` (NHWC) with
``Cs = round_up(C, 8)`` and zero pad channels, dtype ``torch.float32``, ``torch.bfloat16`` or (inference) ``torch.float16``.
The reference's NCHW fp32 tensors exist only at the API edges (:func:`to_storage`,
:func:`from_storage`).  PyTorch is used for memory, streams and the autograd tape only; every
array operation on an activation goes through a ``cgb_*`` entry point.  Nothing here has a CPU
or eager fallback — ops raise :class:`CgbError` without an sm_100 device.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import ConvDesc, check

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}   # fp16: inference only (--half)


def round8(c: int) -> int:
    return (c + 7) // 8 * 8


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _L():
    return _lib.lib()


def _on_device(x: torch.Tensor) -> bool:
    """The one place that decides whether a tensor may be handed to the library (there is no CPU path)."""
    return x.is_cuda


def _chk_storage(x: torch.Tensor, name: str = "x") -> None:
    if x.dim() != 4 or not x.is_contiguous() or x.shape[-1] % 8 or x.dtype not in _DT or not _on_device(x):
        raise ValueError(
            f"{name}: expected a contiguous CUDA NHWC storage tensor with C%8==0 (fp32/bf16/fp16), got "
            f"shape={tuple(x.shape)} dtype={x.dtype} device={x.device} contiguous={x.is_contiguous()}"
        )


# ------------------------------------------------------------------------------------------------
# weight packing:  OIHW fp32  <->  [Os][kh*kw][Is] storage dtype
# ------------------------------------------------------------------------------------------------
# Weights are packed in ONE launch of cgb_pack_weight (bit-exact against the three-launch torch path:
# tests/test_gpu_zz_new_kernels.py::test_pack_weight_kernel, green on B200); CGB_PACK_KERNEL=0 restores the torch path.
_PACK_KERNEL = os.environ.get("CGB_PACK_KERNEL", "1") == "1"


def pack_weight(w: torch.Tensor, dtype: torch.dtype, cis: Optional[int] = None, cos: Optional[int] = None,
                kernel: Optional[bool] = None, with_dgrad: bool = False) -> torch.Tensor:
    """with_dgrad: also build the dgrad packing [cis][taps reversed][cos] in the same launch and hang it on the result
    (``wp._cgb_wt``, where conv_dgrad_raw looks for it)."""
    o, i, kh, kw = w.shape
    cos = cos or round8(o)
    cis = cis or round8(i)
    if (_PACK_KERNEL if kernel is None else kernel) and w.dtype == torch.float32:
        wd = w.detach()
        wd = wd if wd.is_contiguous() else wd.contiguous()
        wp = torch.empty(cos, kh * kw, cis, dtype=dtype, device=w.device)
        if with_dgrad and dtype != torch.float32:
            wt = torch.empty(cis, kh * kw, cos, dtype=dtype, device=w.device)
            check(_L().cgb_pack_weight_dual(_p(wd), _p(wp), _p(wt), _DT[dtype], o, i, kh * kw, cos, cis, _st()), "pack_weight_dual")
            wp._cgb_wt = wt
            return wp
        check(_L().cgb_pack_weight(_p(wd), _p(wp), _DT[dtype], o, i, kh * kw, cos, cis, _st()), "pack_weight")
        return wp
    wp = torch.zeros(cos, kh * kw, cis, dtype=dtype, device=w.device)
    wp[:o, :, :i] = w.detach().permute(0, 2, 3, 1).reshape(o, kh * kw, i)
    return wp


# Packed copies of nn.Parameter weights are cached between optimiser updates: the encoder runs 4 forwards + 2 backwards per
# train step on the same weights (r and s batches in update_G, again in update_D).  A cache entry is valid while the
# parameter object, its storage, its autograd version and the global epoch (bumped by every optimiser update, which writes
# parameters through raw kernels that autograd's version counter cannot see) are unchanged.  Spectrally-normalised weights
# are fresh tensors on every call and are never cached.
_WCACHE = {}
_WEPOCH = [0]


def invalidate_weight_cache() -> None:
    """Called by the optimisers after every parameter update."""
    _WEPOCH[0] += 1
    if len(_WCACHE) > 4096:
        _WCACHE.clear()


_DBG = set(os.environ.get("CGB_DEBUG_DISABLE", "").split(","))   # debugging switches: wcache, viewgrad, aliasparam


def pack_weight_cached(w: torch.Tensor, dtype: torch.dtype, cis: Optional[int] = None, with_dgrad: bool = False) -> torch.Tensor:
    if not isinstance(w, torch.nn.Parameter) or "wcache" in _DBG:
        return pack_weight(w, dtype, cis=cis, with_dgrad=with_dgrad)
    key = (id(w), dtype, cis)
    ent = _WCACHE.get(key)
    if ent is not None and ent[0] == _WEPOCH[0] and ent[1] == w._version and ent[2] == w.data_ptr() and ent[3]() is w:
        return ent[4]
    wp = pack_weight(w, dtype, cis=cis, with_dgrad=with_dgrad)
    import weakref

    _WCACHE[key] = (_WEPOCH[0], w._version, w.data_ptr(), weakref.ref(w), wp)
    return wp


_BEPOCH = [0]   # bumped by every train-mode BatchNorm forward: running statistics are updated through raw kernels too


def cached_pack(params, tag, builder, uses_running_stats=False):
    """Generic form of :func:`pack_weight_cached` for packings built from several nn.Parameters (SPADE's fused gamma||beta
    weights): ``builder()`` runs only when one of ``params`` changed since the cached copy was made."""
    if "wcache" in _DBG or not all(isinstance(p, torch.Tensor) and p.is_leaf for p in params):
        return builder()   # (leaf tensors only: parameters and buffers, whose identity persists between calls)
    import weakref

    key = (tag,) + tuple(id(p) for p in params)
    sig = (_WEPOCH[0], _BEPOCH[0] if uses_running_stats else 0) + tuple((p._version, p.data_ptr()) for p in params)
    ent = _WCACHE.get(key)
    if ent is not None and ent[0] == sig and all(r() is p for r, p in zip(ent[1], params)):
        return ent[2]
    out = builder()
    _WCACHE[key] = (sig, [weakref.ref(p) for p in params], out)
    return out


def unpack_weight_grad(gwp: torch.Tensor, shape) -> torch.Tensor:
    o, i, kh, kw = shape
    if gwp.shape[0] == o and gwp.shape[2] == i and "viewgrad" not in _DBG:
        return gwp.view(o, kh, kw, i).permute(0, 3, 1, 2)   # no channel padding: a strided view, no copy
    return gwp[:o, :, :i].reshape(o, kh, kw, i).permute(0, 3, 1, 2).contiguous()


def pad_bias(b: Optional[torch.Tensor], cos: int) -> Optional[torch.Tensor]:
    if b is None:
        return None
    if b.numel() == cos and b.dtype == torch.float32 and b.is_contiguous() and "aliasparam" not in _DBG:
        return b.detach()
    bp = torch.zeros(cos, dtype=torch.float32, device=b.device)
    bp[: b.numel()] = b.detach().float()
    return bp


# ------------------------------------------------------------------------------------------------
# raw (non-autograd) wrappers
# ------------------------------------------------------------------------------------------------
class ConvGeom:
    """kernel geometry + epilogue of one conv (forward view)."""

    __slots__ = ("kh", "kw", "stride", "dil", "pad", "pad_mode", "act", "slope", "engine", "res_before_act")

    def __init__(self, kh, kw, stride=1, dil=1, pad=0, pad_mode=_lib.PAD_ZERO, act=_lib.ACT_NONE,
                 slope=0.2, engine=_lib.ENGINE_AUTO, res_before_act=0):
        self.kh, self.kw, self.stride, self.dil, self.pad = kh, kw, stride, dil, pad
        self.pad_mode, self.act, self.slope, self.engine = pad_mode, act, slope, engine
        self.res_before_act = res_before_act

    def out_hw(self, hi, wi):
        ho = (hi + 2 * self.pad - self.dil * (self.kh - 1) - 1) // self.stride + 1
        wo = (wi + 2 * self.pad - self.dil * (self.kw - 1) - 1) // self.stride + 1
        return ho, wo

    def desc(self, n, hi, wi, ci, co, dtype) -> ConvDesc:
        ho, wo = self.out_hw(hi, wi)
        return ConvDesc(n, hi, wi, ci, ho, wo, co, self.kh, self.kw, self.stride, self.dil, self.pad,
                        self.pad_mode, _DT[dtype], self.act, self.slope, self.engine, self.res_before_act)


_STATS_ROWS = []


def _stats_rows() -> int:
    if not _STATS_ROWS:
        _STATS_ROWS.append(int(_L().cgb_conv2d_stats_rows()))
    return _STATS_ROWS[0]


_EPI_STATS = os.environ.get("CGB_EPILOGUE_STATS", "1") == "1"   # 0: BatchNorm statistics by a separate pass (A/B switch)


def conv_fwd_raw(x, wp, bias, residual, g: ConvGeom, want_stats: bool = False):
    """want_stats: also return the per-CTA partial sums [rows, 2, co] of the output's per-channel sum / sum of squares when the
    launch runs on the tcgen05 engine (cgb_conv2d_fwd_stats) — (y, partial) with partial None on the SIMT engine."""
    _chk_storage(x)
    n, hi, wi, ci = x.shape
    co = wp.shape[0]
    assert wp.shape[2] == ci and wp.shape[1] == g.kh * g.kw and wp.dtype == x.dtype, (wp.shape, x.shape)
    d = g.desc(n, hi, wi, ci, co, x.dtype)
    y = torch.empty((n, d.ho, d.wo, co), dtype=x.dtype, device=x.device)
    if want_stats:
        partial = None
        if _EPI_STATS and _L().cgb_conv2d_uses_tcgen05(C.byref(d), 0):
            partial = torch.empty((_stats_rows(), 2, co), dtype=torch.float32, device=x.device)
            check(_L().cgb_conv2d_fwd_stats(C.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(y), _p(partial), _st()),
                  "conv2d_fwd_stats")
            return y, partial
        check(_L().cgb_conv2d_fwd(C.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(y), _st()), "conv2d_fwd")
        return y, partial
    check(_L().cgb_conv2d_fwd(C.byref(d), _p(x), _p(wp), _p(bias), _p(residual), _p(y), _st()), "conv2d_fwd")
    return y


def conv_dgrad_raw(gy, wp, x_shape, g: ConvGeom, dact=_lib.ACT_NONE, mask_src=None):
    _chk_storage(gy, "gy")
    n, hi, wi, ci = x_shape
    co = wp.shape[0]
    d = g.desc(n, hi, wi, ci, co, gy.dtype)
    assert (d.ho, d.wo, co) == tuple(gy.shape[1:]), (d.ho, d.wo, co, gy.shape)
    gx = torch.empty(x_shape, dtype=gy.dtype, device=gy.device)
    wt = None
    if _L().cgb_conv2d_uses_tcgen05(C.byref(d), 1):
        wt = getattr(wp, "_cgb_wt", None)   # a cached packed weight (pack_weight_cached) keeps its dgrad packing alongside
        if wt is None or wt.shape != (ci, g.kh * g.kw, co):
            wt = torch.empty((ci, g.kh * g.kw, co), dtype=wp.dtype, device=wp.device)
            check(_L().cgb_conv2d_pack_dgrad_weight(C.byref(d), _p(wp), _p(wt), _st()), "conv2d_pack_dgrad_weight")
            wp._cgb_wt = wt
    check(_L().cgb_conv2d_dgrad(C.byref(d), _p(gy), _p(wp), _p(wt), dact, _p(mask_src), _p(gx), _st()), "conv2d_dgrad")
    return gx


def conv_wgrad_raw(x, gy, g: ConvGeom, want_bias: bool):
    _chk_storage(x)
    _chk_storage(gy, "gy")
    n, hi, wi, ci = x.shape
    co = gy.shape[-1]
    d = g.desc(n, hi, wi, ci, co, x.dtype)
    gw = torch.empty((co, g.kh * g.kw, ci), dtype=torch.float32, device=x.device)
    gb = torch.empty((co,), dtype=torch.float32, device=x.device) if want_bias else None
    check(_L().cgb_conv2d_wgrad(C.byref(d), _p(x), _p(gy), _p(gw), _p(gb), 0, _st()), "conv2d_wgrad")
    return gw, gb


_WS_CACHE = {}


def _stats_ws(n: int, hw: int, c: int) -> int:
    """doubles of scratch cgb_instnorm_stats needs for this shape (per-chunk partial sums)."""
    k = (n, hw, c)
    v = _WS_CACHE.get(k)
    if v is None:
        v = _WS_CACHE[k] = int(_L().cgb_instnorm_ws_doubles(n, hw, c))
    return v


def instnorm_stats(x: torch.Tensor, eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-(n,c) mean and 1/sqrt(var+eps) of a storage tensor (nn.InstanceNorm2d, norms.py:151)."""
    _chk_storage(x)
    n, h, w, c = x.shape
    ws = torch.empty((_stats_ws(n, h * w, c),), dtype=torch.float64, device=x.device)
    mean = torch.empty((n, c), dtype=torch.float32, device=x.device)
    rstd = torch.empty((n, c), dtype=torch.float32, device=x.device)
    check(_L().cgb_instnorm_stats(_p(x), _DT[x.dtype], n, h * w, c, eps, _p(ws), _p(mean), _p(rstd), _st()),
          "instnorm_stats")
    return mean, rstd


def im2col(x: torch.Tensor, c: int, k: int = 3, pad: int = 1, dil: int = 1) -> torch.Tensor:
    """[N,H,W,Cs] storage tensor with c logical channels -> [N,H,W,round8(k*k*c)] patches (tap-major).

    When the rounding leaves a spare channel, channel k*k*c is set to ONE (a "bias tap"): the packed weight of the GEMM that
    consumes the patches is zero there, so the forward is unchanged, and column k*k*c of that GEMM's weight gradient IS the bias
    gradient (sum over pixels of gy) — SPADE's mlp_shared gets its bias gradient out of the wgrad launch instead of a separate
    column-sum pass over the 128-channel gradient of actv (see _Spade.backward; has_bias_tap())."""
    _chk_storage(x)
    n, h, w, cs = x.shape
    cs_out = round8(k * k * c)
    y = torch.empty((n, h, w, cs_out), dtype=x.dtype, device=x.device)
    check(_L().cgb_im2col(_p(x), _p(y), _DT[x.dtype], n, h, w, cs, c, k, pad, dil, cs_out, _st()), "im2col")
    if has_bias_tap(c, k):
        y[..., k * k * c].fill_(1.0)
    return y


def has_bias_tap(c: int, k: int) -> bool:
    """True when :func:`im2col` patches of c channels and a k x k window carry the constant-one channel."""
    return k * k * c < round8(k * k * c)


def im2col_strided(x: torch.Tensor, c: int, k: int, pad: int, dil: int = 1, stride: int = 1) -> torch.Tensor:
    """[N,H,W,Cs] storage tensor with c logical channels -> [N,Ho,Wo,round8(k*k*c)] patches (tap-major) at an output stride.
    Not differentiable w.r.t. x (used on input images)."""
    _chk_storage(x)
    n, h, w, cs = x.shape
    cs_out = round8(k * k * c)
    ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    y = torch.empty((n, ho, wo, cs_out), dtype=x.dtype, device=x.device)
    check(_L().cgb_im2col_strided(_p(x.detach()), _p(y), _DT[x.dtype], n, h, w, cs, c, k, pad, dil, stride, cs_out, _st()),
          "im2col_strided")
    return y


class _Im2colStrided(Function):
    """Differentiable :func:`im2col_strided`: backward is the gather-form adjoint ``cgb_col2im_strided``."""

    @staticmethod
    def forward(ctx, x, c, k, pad, dil, stride):
        ctx.geom = (tuple(x.shape), c, k, pad, dil, stride)
        return im2col_strided(x, c, k, pad, dil, stride)

    @staticmethod
    def backward(ctx, g):
        (n, h, w, cs), c, k, pad, dil, stride = ctx.geom
        g = g.contiguous()
        gx = torch.empty((n, h, w, cs), dtype=g.dtype, device=g.device)
        check(_L().cgb_col2im_strided(_p(g), _p(gx), _DT[g.dtype], n, h, w, cs, c, k, pad, dil, stride, g.shape[-1], _st()),
              "col2im_strided")
        return gx, None, None, None, None, None


# first-layer convs on images (<= 4 logical channels in an 8-channel storage vector): as k*k taps of an 8-channel TMA box the
# tcgen05 engine runs them at a few % of either roofline (the discriminators' 4x4 stride-2 model0 at 640^2: 1.0-1.2 ms against a
# 24 us HBM bound); as im2col + ONE K = round8(k*k*c) GEMM they are a short-K 1x1 conv.  CGB_IM2COL_FIRST=0 turns the route off.
_IM2COL_FIRST = os.environ.get("CGB_IM2COL_FIRST", "1") != "0"


def conv2d_first_layer(x, w, bias=None, *, stride=1, pad=0, act=_lib.ACT_NONE, slope=0.2):
    """conv2d for a conv whose input is an image-like storage tensor with w.shape[1] <= 4 logical channels: im2col (differentiable
    when x needs a gradient) + a 1x1 conv on the patches; falls back to :func:`conv2d` for anything else."""
    co, cin, k, _ = w.shape
    n, h, w_, cs = x.shape
    if not (_IM2COL_FIRST and cin <= 4 and cs == 8 and k >= 3 and min(h, w_) >= 64 and x.dtype != torch.float32):
        return conv2d(x, w, bias, stride=stride, pad=pad, act=act, slope=slope)
    xc = _Im2colStrided.apply(x, cin, k, pad, 1, stride) if x.requires_grad else im2col_strided(x, cin, k, pad, 1, stride)
    kk = k * k * cin
    w2 = w.permute(0, 2, 3, 1).reshape(co, kk)                     # the patches' tap-major order (autograd-native)
    w2 = torch.nn.functional.pad(w2, (0, xc.shape[-1] - kk)).view(co, xc.shape[-1], 1, 1)
    return conv2d(xc, w2, bias, act=act, slope=slope)


def act_bwd_raw(gy, y, act, slope):
    gx = torch.empty_like(gy)
    check(_L().cgb_act_bwd(_p(gy), _p(y), _p(gx), _DT[gy.dtype], gy.numel(), act, slope, _st()), "act_bwd")
    return gx


def act_bwd_bias_raw(gy, y, act, slope):
    """(gx, gbias): :func:`act_bwd_raw` that also returns the per-channel sum of gx over pixels (fp32, storage channels) — the bias
    gradient of the conv whose fused activation is being differentiated."""
    gx = torch.empty_like(gy)
    c = gy.shape[-1]
    gbias = torch.empty((c,), dtype=torch.float32, device=gy.device)
    check(_L().cgb_act_bwd_bias(_p(gy), _p(y), _p(gx), _p(gbias), _DT[gy.dtype], gy.numel() // c, c, act, slope, _st()),
          "act_bwd_bias")
    return gx, gbias


_ACT_BIAS_FUSED = os.environ.get("CGB_ACT_BIAS_FUSED", "1") != "0"


# ------------------------------------------------------------------------------------------------
# autograd functions
# ------------------------------------------------------------------------------------------------
class _Conv2d(Function):
    """y = act(conv(x, w) + b) (+ residual); w is OIHW fp32 (master / spectrally normalised)."""

    @staticmethod
    def forward(ctx, x, w, bias, residual, g: ConvGeom, want_stats=False):
        # a conv whose input needs a gradient will run a dgrad: pack the weight both ways in one launch
        wp = pack_weight_cached(w, x.dtype, cis=x.shape[-1], with_dgrad=bool(ctx.needs_input_grad[0]))
        bp = pad_bias(bias, wp.shape[0])
        partial = None
        if want_stats:
            y, partial = conv_fwd_raw(x, wp, bp, residual, g, True)
        else:
            y = conv_fwd_raw(x, wp, bp, residual, g)
        ctx.want_stats = want_stats
        ctx.g = g
        ctx.w_shape = tuple(w.shape)
        ctx.has_bias = bias is not None
        ctx.nb = 0 if bias is None else bias.numel()
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, wp, y if g.act != _lib.ACT_NONE else None)
        if want_stats:
            if partial is None:
                partial = torch.empty(0, dtype=torch.float32, device=x.device)   # "no statistics": the SIMT engine ran
            ctx.mark_non_differentiable(partial)
            return y, partial
        return y

    @staticmethod
    def backward(ctx, gy, gpartial=None):
        x, wp, y = ctx.saved_tensors
        g = ctx.g
        gy = gy.contiguous()
        gb_fused = None
        if g.act != _lib.ACT_NONE and ctx.has_bias and ctx.needs_input_grad[2] and _ACT_BIAS_FUSED and gy.shape[-1] <= 2048:
            gpre, gb_fused = act_bwd_bias_raw(gy, y, g.act, g.slope)   # the bias gradient rides the activation's backward pass
        else:
            gpre = act_bwd_raw(gy, y, g.act, g.slope) if g.act != _lib.ACT_NONE else gy
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = conv_dgrad_raw(gpre, wp, tuple(x.shape), g)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            gwp, gbp = conv_wgrad_raw(x, gpre, g, ctx.has_bias and gb_fused is None)
            if gb_fused is not None:
                gbp = gb_fused
            gw = unpack_weight_grad(gwp, ctx.w_shape)
            if ctx.has_bias:
                gb = gbp if (gbp.numel() == ctx.nb and "viewgrad" not in _DBG) else gbp[: ctx.nb].clone()
        gres = gy if ctx.has_res else None
        return gx, gw, gb, gres, None, None


class _Conv2dSkip(Function):
    """A bias-free 1x1 stride-1 conv whose INPUT also feeds a skip connection (the ResNet bottleneck's conv1 and its identity
    branch, resnetmulti_v2.py:40-56): ``y, partial, x_skip = f(x, w)`` with ``x_skip`` aliasing ``x``.  Autograd would sum the
    two gradients of x with a separate pass over the 1024-channel tensor (66 bf16 adds per step, 4.8 ms); here the skip gradient
    rides the dgrad launch as its epilogue residual: gx = conv(gy, wt) + g_skip, rounded once."""

    @staticmethod
    def forward(ctx, x, w, want_stats):
        g = ConvGeom(1, 1, 1, 1, 0, _lib.PAD_ZERO, _lib.ACT_NONE, 0.0, _lib.ENGINE_AUTO)
        wp = pack_weight_cached(w, x.dtype, cis=x.shape[-1], with_dgrad=True)
        if want_stats:
            y, partial = conv_fwd_raw(x, wp, None, None, g, True)
        else:
            y, partial = conv_fwd_raw(x, wp, None, None, g), None
        if partial is None:
            partial = torch.empty(0, dtype=torch.float32, device=x.device)
        ctx.g, ctx.w_shape = g, tuple(w.shape)
        ctx.save_for_backward(x, wp)
        ctx.mark_non_differentiable(partial)
        return y, partial, x.view_as(x)

    @staticmethod
    def backward(ctx, gy, gpartial, gskip):
        x, wp = ctx.saved_tensors
        g = ctx.g
        gy = gy.contiguous()
        gx = gw = None
        if ctx.needs_input_grad[0]:
            wt = getattr(wp, "_cgb_wt", None)
            n, hi, wi, ci = x.shape
            d = g.desc(n, hi, wi, ci, wp.shape[0], gy.dtype)
            if gskip is not None and wt is not None and wt.shape == (ci, 1, wp.shape[0]) and _L().cgb_conv2d_uses_tcgen05(C.byref(d), 1):
                gx = conv_fwd_raw(gy, wt, None, gskip.contiguous(), g)      # the 1x1 dgrad IS a 1x1 conv with the transposed packing
            else:
                gx = conv_dgrad_raw(gy, wp, tuple(x.shape), g)
                if gskip is not None:
                    gx = gx + gskip
        if ctx.needs_input_grad[1]:
            gwp, _ = conv_wgrad_raw(x, gy, g, False)
            gw = unpack_weight_grad(gwp, ctx.w_shape)
        return gx, gw, None


# Off by default: measured neutral on the full step (200.0 vs 200.8 ms, scripts/r02/gpu37.sh) — the residual operand of the TMA-path
# epilogue is read row-per-thread (32 sectors in 32 lines per load), which costs what the separate bf16 add cost.  It becomes
# a win once the epilogue's second operand arrives by TMA.
_CONV_SKIP = os.environ.get("CGB_CONV_SKIP", "0") != "0"


def conv2d_skip(x, w, want_stats=False):
    """(y, partial or None, x_skip) — see :class:`_Conv2dSkip`; w: [co, ci, 1, 1], no bias, stride 1."""
    assert w.shape[2] == 1 and w.shape[3] == 1
    y, partial, xs = _Conv2dSkip.apply(x, w, want_stats)
    return y, (partial if partial.numel() else None), xs


def conv2d(x, w, bias=None, residual=None, *, stride=1, dil=1, pad=0, pad_mode=_lib.PAD_ZERO,
           act=_lib.ACT_NONE, slope=0.2, engine=_lib.ENGINE_AUTO, want_stats=False):
    """want_stats: returns (y, partial) — see :func:`conv_fwd_raw`; partial is None when the conv did not produce statistics."""
    g = ConvGeom(w.shape[2], w.shape[3], stride, dil, pad, pad_mode, act, slope, engine)
    if want_stats:
        y, partial = _Conv2d.apply(x, w, bias, residual, g, True)
        return y, (partial if partial.numel() else None)
    return _Conv2d.apply(x, w, bias, residual, g)


def conv_bn_act(x, w, bn, bias=None, residual=None, *, stride=1, dil=1, pad=0, pad_mode=_lib.PAD_ZERO, act=_lib.ACT_NONE,
                slope=0.2, dual=False):
    """conv -> BatchNorm (+ residual) -> activation for an ``nn.BatchNorm2d`` container ``bn`` (resnetmulti_v2.py:40-56,
    blocks.py:138-144): in train mode the batch statistics are accumulated by the conv's own epilogue, so the chain is the conv
    launch, one tiny finalize and ONE pass over the conv output (normalise + affine + residual + activation)."""
    batch_stats = bool(bn.training or bn.running_mean is None)
    if batch_stats:
        y, partial = conv2d(x, w, bias, None, stride=stride, dil=dil, pad=pad, pad_mode=pad_mode, want_stats=True)
    else:
        y, partial = conv2d(x, w, bias, None, stride=stride, dil=dil, pad=pad, pad_mode=pad_mode), None
    return batchnorm_act(y, bn, residual, act, slope, partial=partial, dual=dual)


# two aliases of a bottleneck's output (one per consumer) so that BatchNorm's backward sums their gradients in its own first pass
_BN_DUAL = os.environ.get("CGB_BN_DUAL", "1") != "0"


_SPADE_BIAS_FUSED = os.environ.get("CGB_SPADE_BIAS_FUSED", "1") != "0"


class _Spade(Function):
    """SPADE.forward (climategan/norms.py:174-186) + the leaky-relu that follows it in
    SPADEResnetBlock (blocks.py:372-373), as one autograd node.

    out = act( IN(x) * (1 + conv_g(a)) + conv_b(a) ),  a = relu(conv_sh(seg))
    mean/rstd are the instance-norm statistics of x (shared by norm_0 / norm_s of a block); the
    backward differentiates through them.
    """

    @staticmethod
    def forward(ctx, x, mean, rstd, seg, w_sh, b_sh, w_g, b_g, w_b, b_b, act, slope, engine, seg_is_col, batch_stats=False):
        dt = x.dtype
        n, h, w_, cs = x.shape
        # batch_stats: mean / rstd are [1, cs] statistics over (N, H, W) (the masker's BatchNorm flavour, norms.py:154-155) —
        # the modulation kernels then see the batch as ONE sample of N*H*W pixels (same memory, NHWC)
        ns_, hw_ = (1, n * h * w_) if batch_stats else (n, h * w_)
        c = w_g.shape[0]
        k = w_sh.shape[2]
        pad = k // 2
        g_gb = ConvGeom(k, k, 1, 1, pad, _lib.PAD_ZERO, _lib.ACT_NONE, 0.0, engine)
        if seg_is_col:
            # seg holds im2col patches [.., tap*cin + ch]: mlp_shared becomes a 1x1 conv with K = k*k*cin
            g_sh = ConvGeom(1, 1, 1, 1, 0, _lib.PAD_ZERO, _lib.ACT_RELU, 0.0, engine)
            o_sh, i_sh = w_sh.shape[0], w_sh.shape[1]

            def build_sh():
                wp = torch.zeros(round8(o_sh), 1, seg.shape[-1], dtype=dt, device=x.device)
                wp[:o_sh, 0, : k * k * i_sh] = w_sh.detach().permute(0, 2, 3, 1).reshape(o_sh, k * k * i_sh)
                return wp

            wp_sh = cached_pack([w_sh], ("spade_sh_col", dt, seg.shape[-1]), build_sh)
        else:
            g_sh = ConvGeom(k, k, 1, 1, pad, _lib.PAD_ZERO, _lib.ACT_RELU, 0.0, engine)
            wp_sh = pack_weight_cached(w_sh, dt, cis=seg.shape[-1])
        bp_sh = pad_bias(b_sh, wp_sh.shape[0])
        actv = conv_fwd_raw(seg, wp_sh, bp_sh, None, g_sh)
        nh = actv.shape[-1]

        # gamma || beta as ONE conv with co = 2*cs (norms.py:180-182 share the input actv)
        def build_gb():
            if _PACK_KERNEL and w_g.dtype == torch.float32:   # two launches: each half is a [cs][taps][nh] packing of its own
                wp = torch.empty(2 * cs, k * k, nh, dtype=dt, device=x.device)
                for half, wsrc in ((wp[:cs], w_g), (wp[cs:], w_b)):
                    check(_L().cgb_pack_weight(_p(wsrc.detach().contiguous()), _p(half), _DT[dt], c, nh, k * k, cs, nh, _st()),
                          "pack_weight")
            else:
                wp = torch.zeros(2 * cs, k * k, nh, dtype=dt, device=x.device)
                wp[:c] = w_g.detach().permute(0, 2, 3, 1).reshape(c, k * k, nh)
                wp[cs:cs + c] = w_b.detach().permute(0, 2, 3, 1).reshape(c, k * k, nh)
            bp = torch.zeros(2 * cs, dtype=torch.float32, device=x.device)
            bp[:c] = b_g.detach()
            bp[cs:cs + c] = b_b.detach()
            return wp, bp

        wp_gb, bp_gb = cached_pack([w_g, w_b, b_g, b_b], ("spade_gb", dt, cs, nh), build_gb)
        gb = conv_fwd_raw(actv, wp_gb, bp_gb, None, g_gb)
        out = torch.empty_like(x)
        check(_L().cgb_spade_modulate_fwd(_p(x), _p(mean), _p(rstd), _p(gb), _p(out), _DT[dt], ns_, hw_, cs,
                                          act, slope, _st()), "spade_modulate_fwd")
        ctx.save_for_backward(x, mean, rstd, seg, actv, gb, wp_gb, wp_sh)
        ctx.meta = (act, slope, g_sh, g_gb, tuple(w_sh.shape), tuple(w_g.shape), c, cs, seg_is_col, ns_, hw_)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, mean, rstd, seg, actv, gb, wp_gb, wp_sh = ctx.saved_tensors
        act, slope, g_sh, g_gb, sh_shape, g_shape, c, cs, seg_is_col, ns_, hw_ = ctx.meta
        dt = x.dtype
        gout = gout.contiguous()
        ggb = torch.empty_like(gb)
        gx = torch.empty_like(x)
        want_w = any(ctx.needs_input_grad[4:10])   # False while the painter is frozen (painter loss for the masker)
        # one zero-filled allocation: the instance-norm sums [ns, cs, 2] and, behind them, the bias gradients of mlp_gamma / mlp_beta
        # [2 cs] — the modulation pass returns the column sums of ggb, so the gamma||beta wgrad launch needs no column-sum pass
        buf = torch.zeros((ns_ * cs * 2 + (2 * cs if want_w and _SPADE_BIAS_FUSED else 0),), dtype=torch.float64, device=x.device)
        sums = buf[: ns_ * cs * 2].view(ns_, cs, 2)
        bsum = buf[ns_ * cs * 2:] if want_w and _SPADE_BIAS_FUSED else None
        if bsum is not None:
            check(_L().cgb_spade_modulate_bwd_bias(_p(x), _p(mean), _p(rstd), _p(gb), _p(gout), _p(ggb), _p(gx), _p(sums), _p(bsum),
                                                   _DT[dt], ns_, hw_, cs, act, slope, _st()), "spade_modulate_bwd_bias")
        else:
            check(_L().cgb_spade_modulate_bwd(_p(x), _p(mean), _p(rstd), _p(gb), _p(gout), _p(ggb), _p(gx), _p(sums),
                                              _DT[dt], ns_, hw_, cs, act, slope, _st()), "spade_modulate_bwd")
        check(_L().cgb_instnorm_bwd(_p(x), _p(mean), _p(rstd), _p(sums), _p(gx), _DT[dt], ns_, hw_, cs, _st()),
              "instnorm_bwd")
        # gamma||beta conv: weight grads + data grad (ReLU of mlp_shared fused as a mask)
        want_seg = ctx.needs_input_grad[3]
        gw_sh = gb_sh = gw_g = gb_g = gw_b = gb_b = gactv = None
        if want_w or want_seg:
            gactv = conv_dgrad_raw(ggb, wp_gb, tuple(actv.shape), g_gb, _lib.ACT_RELU, actv)
        if want_w:
            gwp_gb, gbp_gb = conv_wgrad_raw(actv, ggb, g_gb, bsum is None)
            if bsum is not None:
                gbp_gb = bsum.float()
            bias_tap = seg_is_col and has_bias_tap(sh_shape[1], sh_shape[2])   # the patches' constant-one channel (im2col)
            gwp_sh, gbp_sh = conv_wgrad_raw(seg, gactv, g_sh, not bias_tap)
            gw_g = unpack_weight_grad(gwp_gb[:cs], g_shape)
            gw_b = unpack_weight_grad(gwp_gb[cs:], g_shape)
            gb_g = gbp_gb[:c].clone()
            gb_b = gbp_gb[cs:cs + c].clone()
            if seg_is_col:
                o_sh, i_sh, kk, _ = sh_shape
                gw_sh = gwp_sh[:o_sh, 0, : kk * kk * i_sh].reshape(o_sh, kk, kk, i_sh).permute(0, 3, 1, 2).contiguous()
            else:
                gw_sh = unpack_weight_grad(gwp_sh, sh_shape)
            if bias_tap:   # column k*k*cin of the patch GEMM's weight gradient = sum over pixels of gactv = the bias gradient
                gb_sh = gwp_sh[: sh_shape[0], 0, sh_shape[2] * sh_shape[2] * sh_shape[1]].clone()
            else:
                gb_sh = gbp_sh[: sh_shape[0]].clone()
        if not ctx.needs_input_grad[0]:
            gx = None
        gseg = None
        if want_seg:
            # differentiable conditioning (the masker's make_m_cond(d, s, x) with gen.m.spade.detach = false): dgrad of mlp_shared
            if seg_is_col:
                raise NotImplementedError("SPADE: gradient w.r.t. an im2col'd conditioning tensor is not built")
            gseg = conv_dgrad_raw(gactv, wp_sh, tuple(seg.shape), g_sh)
        return gx, None, None, gseg, gw_sh, gb_sh, gw_g, gb_g, gw_b, gb_b, None, None, None, None, None


def spade(x, mean, rstd, seg, w_sh, b_sh, w_g, b_g, w_b, b_b, act=_lib.ACT_NONE, slope=0.2,
          engine=_lib.ENGINE_AUTO, seg_is_col=False, batch_stats=False):
    """seg: conditioning storage tensor at x's resolution, or (seg_is_col) its im2col patches from
    :func:`im2col` — then SPADE.mlp_shared runs as one K=round8(9*cond_nc) GEMM.  batch_stats: mean / rstd are [1, Cs]
    statistics over the whole batch (train-mode BatchNorm flavour); the backward differentiates through them too."""
    return _Spade.apply(x, mean, rstd, seg, w_sh, b_sh, w_g, b_g, w_b, b_b, act, slope, engine, seg_is_col, batch_stats)


def batchnorm_stats_update(x, bn):
    """Train-mode statistics of a parameter-free ``nn.BatchNorm2d`` (SPADE's batch flavour): per-channel mean and
    1/sqrt(biased var + eps) over (N, H, W) as [1, Cs] tensors, with the running-statistics update F.batch_norm does
    (momentum, unbiased variance, num_batches_tracked)."""
    _chk_storage(x)
    n, h, w, cs = x.shape
    mean, rstd = instnorm_stats(x.view(1, n * h, w, cs), bn.eps)
    if bn.track_running_stats and bn.running_mean is not None:
        c = bn.running_mean.numel()
        momentum = 0.1 if bn.momentum is None else bn.momentum
        _BEPOCH[0] += 1   # folded eval-mode packings built from running statistics are stale
        check(_L().cgb_bn_update_running(_p(mean), _p(rstd), _p(bn.running_mean), _p(bn.running_var), c, n * h * w,
                                         float(momentum), float(bn.eps), _st()), "bn_update_running")
        bn.num_batches_tracked.add_(1)
    return mean, rstd


class _ResizeNearest(Function):
    @staticmethod
    def forward(ctx, x, ho, wo):
        _chk_storage(x)
        n, hi, wi, c = x.shape
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_resize_nearest_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, ho, wo, c, _st()), "resize_nearest")
        ctx.shape = (n, hi, wi, c)
        ctx.f = ho // hi if (hi and ho % hi == 0 and wo % wi == 0 and ho // hi == wo // wi) else 0
        return y

    @staticmethod
    def backward(ctx, gy):
        n, hi, wi, c = ctx.shape
        gy = gy.contiguous()
        gx = torch.empty(ctx.shape, dtype=gy.dtype, device=gy.device)
        if ctx.f >= 1:
            check(_L().cgb_upsample_nearest_bwd(_p(gy), _p(gx), _DT[gy.dtype], n, hi, wi, ctx.f, c, _st()), "upsample_bwd")
        else:   # any other ratio (e.g. the masker's conditioning tensor down-sized to the latent's resolution)
            check(_L().cgb_resize_nearest_bwd(_p(gy), _p(gx), _DT[gy.dtype], n, hi, wi, gy.shape[1], gy.shape[2], c, _st()),
                  "resize_nearest_bwd")
        return gx, None, None


def resize_nearest(x, ho, wo):
    """F.interpolate(x, size=(ho, wo), mode='nearest') on a storage tensor."""
    return _ResizeNearest.apply(x, ho, wo)


def upsample2x(x):
    """InterpolateNearest2d(scale_factor=2) (climategan/blocks.py:11-43)."""
    return _ResizeNearest.apply(x, x.shape[1] * 2, x.shape[2] * 2)


class _Act(Function):
    @staticmethod
    def forward(ctx, x, act, slope):
        y = torch.empty_like(x)
        check(_L().cgb_act_fwd(_p(x), _p(y), _DT[x.dtype], x.numel(), act, slope, _st()), "act_fwd")
        ctx.save_for_backward(y)
        ctx.meta = (act, slope)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        return act_bwd_raw(gy.contiguous(), y, *ctx.meta), None, None


def activation(x, act, slope=0.2):
    return _Act.apply(x, act, slope)


class _ToStorage(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        x = x.contiguous().float()
        n, c, h, w = x.shape
        cs = round8(c)
        y = torch.empty((n, h, w, cs), dtype=dtype, device=x.device)
        check(_L().cgb_nchw_to_nhwc(_p(x), _p(y), _DT[dtype], n, c, h * w, cs, _st()), "nchw_to_nhwc")
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        n, h, w, cs = gy.shape
        gx = torch.empty((n, ctx.c, h, w), dtype=torch.float32, device=gy.device)
        check(_L().cgb_nhwc_to_nchw(_p(gy), _p(gx), _DT[gy.dtype], n, ctx.c, h * w, cs, _st()), "nhwc_to_nchw")
        return gx, None


class _FromStorage(Function):
    @staticmethod
    def forward(ctx, x, c):
        _chk_storage(x)
        n, h, w, cs = x.shape
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        check(_L().cgb_nhwc_to_nchw(_p(x), _p(y), _DT[x.dtype], n, c, h * w, cs, _st()), "nhwc_to_nchw")
        ctx.meta = (cs, x.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        cs, dtype = ctx.meta
        gy = gy.contiguous().float()
        n, c, h, w = gy.shape
        gx = torch.empty((n, h, w, cs), dtype=dtype, device=gy.device)
        check(_L().cgb_nchw_to_nhwc(_p(gy), _p(gx), _DT[dtype], n, c, h * w, cs, _st()), "nchw_to_nhwc")
        return gx, None


def to_storage(x_nchw: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """NCHW fp32 (reference layout) -> NHWC storage tensor with zero channel padding."""
    _lib.require_device()
    if not _on_device(x_nchw):
        raise _lib.CgbError("climategan_b200 tensors must live on a CUDA device (no CPU path)")
    return _ToStorage.apply(x_nchw, dtype)


def from_storage(x: torch.Tensor, c: int) -> torch.Tensor:
    """NHWC storage tensor -> NCHW fp32 with the logical channel count ``c``."""
    return _FromStorage.apply(x, c)


# ------------------------------------------------------------------------------------------------
# compositing / losses
# ------------------------------------------------------------------------------------------------
class _MaskCond(Function):
    @staticmethod
    def forward(ctx, x, m, dtype):
        n, c, h, w = x.shape
        cond = torch.empty((n, h, w, 8), dtype=dtype, device=x.device)
        check(_L().cgb_mask_cond(_p(x), _p(m), _p(cond), _DT[dtype], n, h * w, 8, _st()), "mask_cond")
        ctx.save_for_backward(x)
        return cond

    @staticmethod
    def backward(ctx, gcond):
        (x,) = ctx.saved_tensors
        n, _, h, w = x.shape
        gcond = gcond.contiguous()
        gm = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
        check(_L().cgb_mask_cond_bwd(_p(x), _p(gcond), _p(gm), _DT[gcond.dtype], n, h * w, gcond.shape[-1], _st()), "mask_cond_bwd")
        return None, gm, None


def mask_cond(x: torch.Tensor, m: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Painter conditioning ``x * (1 - m)`` (generator.py:294) written straight into storage layout.  Differentiable w.r.t. the
    mask only (the painter loss for the masker feeds the masker's prediction in, trainer.py:1618-1651); x is data."""
    _lib.require_device()
    x = x.detach().contiguous().float()
    m = m.contiguous().float()
    n, c, h, w = x.shape
    assert c == 3 and m.shape == (n, 1, h, w), (x.shape, m.shape)
    return _MaskCond.apply(x, m, dtype)


class _Paste(Function):
    @staticmethod
    def forward(ctx, x, m, fake):
        x = x.contiguous().float()
        m = m.contiguous().float()
        fake = fake.contiguous()
        n, _, h, w = x.shape
        out = torch.empty_like(x)
        check(_L().cgb_paste_fwd(_p(x), _p(m), _p(fake), _p(out), n, h * w, _st()), "paste_fwd")
        if ctx.needs_input_grad[1]:
            ctx.save_for_backward(m, x, fake)
        else:
            ctx.save_for_backward(m)
        return out

    @staticmethod
    def backward(ctx, gout):
        m = ctx.saved_tensors[0]
        gout = gout.contiguous()
        n, _, h, w = gout.shape
        gf = gm = None
        if ctx.needs_input_grad[2]:
            gf = torch.empty_like(gout)
            check(_L().cgb_paste_bwd(_p(gout), _p(m), _p(gf), n, h * w, _st()), "paste_bwd")
        if ctx.needs_input_grad[1]:
            _, x, fake = ctx.saved_tensors
            gm = torch.empty_like(m)
            check(_L().cgb_paste_bwd_mask(_p(gout), _p(x), _p(fake), _p(gm), n, h * w, _st()), "paste_bwd_mask")
        return None, gm, gf


def paste(x, m, fake):
    """``x * (1 - m) + fake * m`` (generator.py:295-296); gradient flows to ``fake`` and, when it asks for one, to the mask."""
    return _Paste.apply(x.detach(), m, fake)


class _L1(Function):
    @staticmethod
    def forward(ctx, a, b):
        a = a.contiguous()
        b = b.contiguous()
        loss = torch.zeros((), dtype=torch.float32, device=a.device)
        ga = torch.empty_like(a)
        check(_L().cgb_l1_loss(_p(a), _p(b), _p(loss), _p(ga), a.numel(), 1.0 / a.numel(), _st()), "l1_loss")
        ctx.save_for_backward(ga)
        return loss

    @staticmethod
    def backward(ctx, g):
        (ga,) = ctx.saved_tensors
        return ga * g, None


def l1_loss(a, b):
    """nn.L1Loss()(a, b) (mean reduction; climategan/losses.py:290-301) — fp32 tensors."""
    return _L1.apply(a, b)


# ------------------------------------------------------------------------------------------------
# spectral norm
# ------------------------------------------------------------------------------------------------
class _SpectralWeight(Function):
    """SpectralNorm._update_u_v (climategan/norms.py:100-112): one power iteration that mutates
    u and v in place (on every call, train or eval), then w = w_bar / sigma with sigma = u.(W v)
    differentiated w.r.t. w_bar only (u, v are .data in the reference)."""

    @staticmethod
    def forward(ctx, w_bar, u, v):
        rows = w_bar.shape[0]
        cols = w_bar.numel() // rows
        wb = w_bar.detach().contiguous()
        sigma = torch.empty((1,), dtype=torch.float32, device=w_bar.device)
        check(_L().cgb_spectral_power_iter(_p(wb), _p(u), _p(v), _p(sigma), rows, cols, _st()), "spectral_power_iter")
        w = wb / sigma
        ctx.save_for_backward(w, sigma)
        # u, v are kept BY REFERENCE, not snapshotted: the reference's autograd graph holds the u/v Parameters themselves and
        # _update_u_v swaps their .data (norms.py:106-108), so when a layer runs twice before one backward (mask decoder and
        # AdvEnt discriminators: r batch then s batch) both backward passes see the u, v of the last forward.  Kept as is.
        ctx.u, ctx.v = u, v
        # d(sigma)/du = W v is the one factor autograd evaluated at FORWARD time (a saved intermediate, not a leaf):
        # W v = sigma * u_forward since u = normalize(W v) and sigma = u.(W v)
        ctx.u_fwd = u.detach().clone() if u.requires_grad else None
        return w

    @staticmethod
    def backward(ctx, gw):
        w, sigma = ctx.saved_tensors
        u, v = ctx.u, ctx.v
        rows = w.shape[0]
        # d(w_bar/sigma) = g/sigma - <g, w_bar>/sigma^2 * u v^T = (g - <g, w> u v^T) / sigma
        dot = (gw * w).sum()
        gwb = (gw - dot * torch.outer(u, v).view_as(w)) / sigma
        gu = gv = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            # Trainer.run_epoch un-freezes the discriminator with `param.requires_grad = True` on EVERY parameter
            # (trainer.py:971-973), u and v included, so from then on sigma = u.(W v) is differentiated w.r.t. them too and
            # ExtraAdam moves them: dL/dsigma = -<g, w_bar>/sigma^2 = -dot/sigma ; dsigma/du = W v ; dsigma/dv = W^T u.
            w2 = (w * sigma).view(rows, -1)
            coef = -dot / sigma
            if ctx.needs_input_grad[1]:
                gu = coef * sigma * (ctx.u_fwd if ctx.u_fwd is not None else u)
            if ctx.needs_input_grad[2]:
                gv = coef * torch.mv(w2.t(), u)
        return gwb.view_as(w), gu, gv


def spectral_weight(w_bar, u, v):
    return _SpectralWeight.apply(w_bar, u, v)


# ------------------------------------------------------------------------------------------------
# discriminator path: instance norm + activation, avg-pool between scales, GAN / feature-matching losses
# ------------------------------------------------------------------------------------------------
class _InstNormAct(Function):
    """act(InstanceNorm2d(affine=False)(x)) — the conv -> IN -> LeakyReLU(0.2) block of NLayerDiscriminator
    (climategan/discriminator.py:120-133, 146-148)."""

    @staticmethod
    def forward(ctx, x, act, slope, eps):
        mean, rstd = instnorm_stats(x, eps)
        n, h, w, c = x.shape
        y = torch.empty_like(x)
        check(_L().cgb_instnorm_apply_fwd(_p(x), _p(mean), _p(rstd), _p(y), _DT[x.dtype], n, h * w, c, act, slope, _st()),
              "instnorm_apply_fwd")
        ctx.save_for_backward(x, mean, rstd)
        ctx.meta = (act, slope)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd = ctx.saved_tensors
        act, slope = ctx.meta
        n, h, w, c = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        sums = torch.zeros((n, c, 2), dtype=torch.float64, device=x.device)
        check(_L().cgb_instnorm_apply_bwd(_p(x), _p(mean), _p(rstd), _p(gy), _p(gx), _p(sums), _DT[x.dtype], n, h * w, c,
                                          act, slope, _st()), "instnorm_apply_bwd")
        check(_L().cgb_instnorm_bwd(_p(x), _p(mean), _p(rstd), _p(sums), _p(gx), _DT[x.dtype], n, h * w, c, _st()),
              "instnorm_bwd")
        return gx, None, None, None


def instnorm_act(x, act=_lib.ACT_NONE, slope=0.2, eps=1e-5):
    return _InstNormAct.apply(x, act, slope, eps)


class _AvgPool3s2(Function):
    """nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False) (discriminator.py:223-225)."""

    @staticmethod
    def forward(ctx, x):
        _chk_storage(x)
        n, hi, wi, c = x.shape
        ho, wo = (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_avgpool3s2_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, c, _st()), "avgpool3s2_fwd")
        ctx.shape = (n, hi, wi, c)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, hi, wi, c = ctx.shape
        gy = gy.contiguous()
        gx = torch.empty(ctx.shape, dtype=gy.dtype, device=gy.device)
        check(_L().cgb_avgpool3s2_bwd(_p(gy), _p(gx), _DT[gy.dtype], n, hi, wi, c, _st()), "avgpool3s2_bwd")
        return gx


def avgpool3s2(x):
    return _AvgPool3s2.apply(x)


LOSS_BCE_LOGITS, LOSS_MSE, LOSS_HINGE_D_REAL, LOSS_HINGE_D_FAKE, LOSS_NEG_MEAN = 0, 1, 2, 3, 4


class _ConstTargetLoss(Function):
    @staticmethod
    def forward(ctx, x, kind, target):
        x = x.contiguous().float()
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        gx = torch.empty_like(x)
        if isinstance(target, torch.Tensor):   # a device-resident target (graphs.StepTape slot)
            check(_L().cgb_const_target_loss_dev(_p(x), _p(loss), _p(gx), x.numel(), kind, _p(target), 1.0 / x.numel(), _st()),
                  "const_target_loss_dev")
        else:
            check(_L().cgb_const_target_loss(_p(x), _p(loss), _p(gx), x.numel(), kind, float(target), 1.0 / x.numel(), _st()),
                  "const_target_loss")
        ctx.save_for_backward(gx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gx,) = ctx.saved_tensors
        return gx * g, None, None


def const_target_loss(x, kind, target=0.0):
    """mean-reduced BCE-with-logits / MSE / hinge against a constant target (fp32 tensor of any shape); ``target`` is a float
    or a 1-element fp32 device tensor read at execution time."""
    return _ConstTargetLoss.apply(x, kind, target)


class _L1Storage(Function):
    @staticmethod
    def forward(ctx, a, b, logical_count):
        _chk_storage(a)
        assert a.shape == b.shape and a.dtype == b.dtype
        loss = torch.zeros((), dtype=torch.float32, device=a.device)
        ga = torch.empty_like(a)
        check(_L().cgb_l1_loss_storage(_p(a), _p(b.contiguous()), _p(loss), _p(ga), _DT[a.dtype], a.numel(),
                                       1.0 / logical_count, _st()), "l1_loss_storage")
        ctx.save_for_backward(ga)
        return loss

    @staticmethod
    def backward(ctx, g):
        (ga,) = ctx.saved_tensors
        return ga * g.to(ga.dtype), None, None


def l1_loss_storage(a, b, c_logical):
    """nn.L1Loss()(a, b.detach()) for two storage tensors with ``c_logical`` real channels (pad channels are equal
    zeros on both sides and contribute nothing; the mean divides by the logical element count)."""
    n, h, w, _ = a.shape
    return _L1Storage.apply(a, b.detach(), n * h * w * c_logical)


# ------------------------------------------------------------------------------------------------
# VGG perceptual loss edge
# ------------------------------------------------------------------------------------------------
class _MaxPool2(Function):
    @staticmethod
    def forward(ctx, x):
        _chk_storage(x)
        n, hi, wi, c = x.shape
        y = torch.empty((n, hi // 2, wi // 2, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_maxpool2_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, c, _st()), "maxpool2_fwd")
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y = ctx.saved_tensors
        n, hi, wi, c = x.shape
        gx = torch.zeros_like(x) if (hi % 2 or wi % 2) else torch.empty_like(x)
        check(_L().cgb_maxpool2_bwd(_p(x), _p(y), _p(gy.contiguous()), _p(gx), _DT[x.dtype], n, hi, wi, c, _st()), "maxpool2_bwd")
        return gx


def maxpool2(x):
    """nn.MaxPool2d(kernel_size=2, stride=2) on a storage tensor."""
    return _MaxPool2.apply(x)


class _VggPreprocess(Function):
    @staticmethod
    def forward(ctx, img, m, dtype):
        img = img.contiguous().float()
        n, c, h, w = img.shape
        assert c == 3
        mm = None if m is None else m.detach().contiguous().float()
        y = torch.empty((n, h, w, 8), dtype=dtype, device=img.device)
        check(_L().cgb_vgg_preprocess_fwd(_p(img), _p(mm), _p(y), _DT[dtype], n, h * w, _st()), "vgg_preprocess_fwd")
        ctx.mm = mm
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        n, h, w, _ = gy.shape
        gx = torch.empty((n, 3, h, w), dtype=torch.float32, device=gy.device)
        check(_L().cgb_vgg_preprocess_bwd(_p(gy), _p(ctx.mm), _p(gx), _DT[gy.dtype], n, h * w, _st()), "vgg_preprocess_bwd")
        return gx, None, None


def vgg_preprocess(img, m, dtype):
    """vgg_preprocess(img * m) (tutils.py:416-427; trainer.py:1281-1283) -> storage tensor [N,H,W,8]."""
    return _VggPreprocess.apply(img, m, dtype)


def cat_mask_image(m, img):
    """torch.cat([m, img], dim=1) (trainer.py:1363) with gradient flowing to ``img`` only — a pure copy; kept as a
    torch op (memory plumbing, no arithmetic)."""
    return torch.cat([m.detach(), img], dim=1)


# ------------------------------------------------------------------------------------------------
# masker inference helpers (forward only: the masker is built for inference so far)
# ------------------------------------------------------------------------------------------------
def conv2d_infer(x, wp, bias, residual=None, *, k, stride=1, dil=1, pad=0, pad_mode=_lib.PAD_ZERO, act=_lib.ACT_NONE,
                 slope=0.2, res_before_act=0):
    """Forward-only conv on pre-packed weights (eval-mode BatchNorm already folded into wp / bias)."""
    g = ConvGeom(k, k, stride, dil, pad, pad_mode, act, slope, _lib.ENGINE_AUTO, res_before_act)
    return conv_fwd_raw(x, wp, bias, residual, g)


def _pool_out(sz):
    o = -(-(sz - 3) // 2) + 1
    if (o - 1) * 2 >= sz:
        o -= 1
    return o


class _MaxPool3s2Ceil(Function):
    @staticmethod
    def forward(ctx, x):
        _chk_storage(x)
        n, hi, wi, c = x.shape
        ho, wo = _pool_out(hi), _pool_out(wi)
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_maxpool3s2_ceil_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, ho, wo, c, _st()), "maxpool3s2_ceil")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        n, hi, wi, c = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        check(_L().cgb_maxpool3s2_ceil_bwd(_p(x), _p(gy), _p(gx), _DT[x.dtype], n, hi, wi, gy.shape[1], gy.shape[2], c, _st()),
              "maxpool3s2_ceil_bwd")
        return gx


class _MaxPool3s2Pad1(Function):
    @staticmethod
    def forward(ctx, x):
        _chk_storage(x)
        n, hi, wi, c = x.shape
        ho, wo = (hi + 2 - 3) // 2 + 1, (wi + 2 - 3) // 2 + 1
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_maxpool3s2_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, ho, wo, c, 1, _st()), "maxpool3s2")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        n, hi, wi, c = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        check(_L().cgb_maxpool3s2_bwd(_p(x), _p(gy), _p(gx), _DT[x.dtype], n, hi, wi, gy.shape[1], gy.shape[2], c, 1, _st()),
              "maxpool3s2_bwd")
        return gx


def maxpool3s2_pad1(x):
    """nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (deeplab/resnet101_v3.py:75)."""
    return _MaxPool3s2Pad1.apply(x)


def maxpool3s2_ceil(x):
    """nn.MaxPool2d(3, stride=2, padding=0, ceil_mode=True) (deeplab/resnetmulti_v2.py:76-78)."""
    return _MaxPool3s2Ceil.apply(x)


class _ResizeBilinear(Function):
    @staticmethod
    def forward(ctx, x, ho, wo, ac):
        _chk_storage(x)
        n, hi, wi, c = x.shape
        y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_resize_bilinear_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, ho, wo, c, ac, _st()), "resize_bilinear")
        ctx.meta = (n, hi, wi, c, ho, wo, ac)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, hi, wi, c, ho, wo, ac = ctx.meta
        gy = gy.contiguous()
        gx = torch.empty((n, hi, wi, c), dtype=gy.dtype, device=gy.device)
        check(_L().cgb_resize_bilinear_bwd(_p(gy), _p(gx), _DT[gy.dtype], n, hi, wi, ho, wo, c, ac, _st()), "resize_bilinear_bwd")
        return gx, None, None, None


def resize_bilinear(x, ho, wo, align_corners=False):
    return _ResizeBilinear.apply(x, ho, wo, 1 if align_corners else 0)


def resize_bicubic(x, ho, wo):
    """Forward only (depth.py:144-149 runs it only when the prediction is not at the target size, i.e. at inference)."""
    _chk_storage(x)
    if x.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("bicubic resize has no backward (not on the training path at the reference's sizes)")
    n, hi, wi, c = x.shape
    y = torch.empty((n, ho, wo, c), dtype=x.dtype, device=x.device)
    check(_L().cgb_resize_bicubic_fwd(_p(x), _p(y), _DT[x.dtype], n, hi, wi, ho, wo, c, _st()), "resize_bicubic")
    return y


class _ChannelMean(Function):
    @staticmethod
    def forward(ctx, x, c_logical):
        _chk_storage(x)
        n, h, w, cs = x.shape
        y = torch.empty((n, h, w, 8), dtype=x.dtype, device=x.device)
        check(_L().cgb_channel_mean(_p(x), _p(y), _DT[x.dtype], n * h * w, cs, c_logical, _st()), "channel_mean")
        ctx.meta = (n, h, w, cs, c_logical)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, h, w, cs, c_logical = ctx.meta
        gy = gy.contiguous()
        gx = torch.empty((n, h, w, cs), dtype=gy.dtype, device=gy.device)
        check(_L().cgb_channel_mean_bwd(_p(gy), _p(gx), _DT[gy.dtype], n * h * w, cs, c_logical, _st()), "channel_mean_bwd")
        return gx, None


def channel_mean(x, c_logical):
    """torch.mean(x, dim=1, keepdim=True) (depth.py:142) -> storage [N,H,W,8] with 1 real channel."""
    return _ChannelMean.apply(x, c_logical)


def _mul_raw(a, b):
    y = torch.empty_like(a)
    check(_L().cgb_mul(_p(a), _p(b), _p(y), _DT[a.dtype], a.numel(), _st()), "mul")
    return y


class _Mul(Function):
    @staticmethod
    def forward(ctx, a, b):
        _chk_storage(a)
        assert a.shape == b.shape and a.dtype == b.dtype
        b = b.contiguous()
        ctx.save_for_backward(a, b)
        return _mul_raw(a, b)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous()
        ga = _mul_raw(g, b) if ctx.needs_input_grad[0] else None
        gb = _mul_raw(g, a) if ctx.needs_input_grad[1] else None
        return ga, gb


def mul(a, b):
    """Elementwise product of two storage tensors (z * z_depth, deeplab_v2.py:193, blocks.py:306)."""
    return _Mul.apply(a, b)


def _broadcast_hw_raw(src, h, w, scale):
    n, c = src.shape[0], src.shape[-1]
    dst = torch.empty((n, h, w, c), dtype=src.dtype, device=src.device)
    check(_L().cgb_broadcast_hw(_p(src.contiguous()), _p(dst), _DT[src.dtype], n, h * w, c, float(scale), _st()), "broadcast_hw")
    return dst


def _spatial_mean(x):
    mean, _ = instnorm_stats(x)
    return mean


class _GlobalMean(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return _spatial_mean(x).to(x.dtype).view(x.shape[0], 1, 1, x.shape[-1]).contiguous()

    @staticmethod
    def backward(ctx, g):
        n, h, w, c = ctx.shape
        return _broadcast_hw_raw(g.contiguous(), h, w, 1.0 / (h * w))


def global_mean(x):
    """AdaptiveAvgPool2d(1) as a storage tensor [N,1,1,Cs] (ASPP global branch, deeplab_v2.py:97-102)."""
    return _GlobalMean.apply(x)


class _BroadcastHW(Function):
    @staticmethod
    def forward(ctx, x, h, w):
        ctx.hw = (h, w)
        return _broadcast_hw_raw(x, h, w, 1.0)

    @staticmethod
    def backward(ctx, g):
        h, w = ctx.hw
        g = g.contiguous()
        gm = (_spatial_mean(g) * float(h * w)).to(g.dtype)
        return gm.view(g.shape[0], 1, 1, g.shape[-1]).contiguous(), None, None


def broadcast_hw(x, h, w):
    """F.interpolate of a 1x1 map to (h, w) (bilinear, align_corners=True: a pure broadcast; deeplab_v2.py:116)."""
    assert x.shape[1] == 1 and x.shape[2] == 1
    return _BroadcastHW.apply(x, h, w)


class _ReflectPad(Function):
    @staticmethod
    def forward(ctx, x, pad):
        _chk_storage(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_reflect_pad_fwd(_p(x), _p(y), _DT[x.dtype], n, h, w, c, pad, _st()), "reflect_pad_fwd")
        ctx.meta = (n, h, w, c, pad)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, h, w, c, pad = ctx.meta
        gy = gy.contiguous()
        gx = torch.empty((n, h, w, c), dtype=gy.dtype, device=gy.device)
        check(_L().cgb_reflect_pad_bwd(_p(gy), _p(gx), _DT[gy.dtype], n, h, w, c, pad, _st()), "reflect_pad_bwd")
        return gx, None


def reflect_pad(x, pad):
    """nn.ReflectionPad2d(pad) (blocks.py:66-67) as an explicit padded copy: the conv that follows then runs with pad 0
    on the tcgen05 engine, and the adjoint folds the border gradients back."""
    return x if pad == 0 else _ReflectPad.apply(x, pad)


def _dropout_raw(x, p, seed):
    y = torch.empty_like(x)
    if isinstance(seed, torch.Tensor):   # device-resident seed (graphs.StepTape slot)
        check(_L().cgb_dropout_dev(_p(x), _p(y), _DT[x.dtype], x.numel(), p, _p(seed), _st()), "dropout_dev")
    else:
        check(_L().cgb_dropout(_p(x), _p(y), _DT[x.dtype], x.numel(), p, seed, _st()), "dropout")
    return y


class _ReplicatePad(Function):
    @staticmethod
    def forward(ctx, x, pad):
        _chk_storage(x)
        n, h, w, c = x.shape
        y = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=x.dtype, device=x.device)
        check(_L().cgb_replicate_pad_fwd(_p(x), _p(y), _DT[x.dtype], n, h, w, c, pad, _st()), "replicate_pad_fwd")
        ctx.meta = (n, h, w, c, pad)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, h, w, c, pad = ctx.meta
        gy = gy.contiguous()
        gx = torch.empty((n, h, w, c), dtype=gy.dtype, device=gy.device)
        check(_L().cgb_replicate_pad_bwd(_p(gy), _p(gx), _DT[gy.dtype], n, h, w, c, pad, _st()), "replicate_pad_bwd")
        return gx, None


def replicate_pad(x, pad):
    """nn.ReplicationPad2d(pad) (blocks.py:68-69) as an explicit padded copy; the conv that follows runs with pad 0."""
    return x if pad == 0 else _ReplicatePad.apply(x, pad)


class _AffineNC(Function):
    """y = act(x * scale[n, c] + shift[n, c]) on a storage tensor; differentiable w.r.t. x, scale and shift."""

    @staticmethod
    def forward(ctx, x, scale, shift, act, slope):
        _chk_storage(x)
        n, h, w, c = x.shape
        scale, shift = scale.contiguous().float(), shift.contiguous().float()
        assert scale.shape == (n, c) and shift.shape == (n, c), (scale.shape, shift.shape, x.shape)
        y = torch.empty_like(x)
        check(_L().cgb_affine_nc_fwd(_p(x), _p(scale), _p(shift), _p(y), _DT[x.dtype], n, h * w, c, act, slope, _st()), "affine_nc_fwd")
        ctx.save_for_backward(x, scale, y if act != _lib.ACT_NONE else None)
        ctx.meta = (act, slope)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, scale, y = ctx.saved_tensors
        act, slope = ctx.meta
        n, h, w, c = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x)
        sums = torch.zeros((n, c, 2), dtype=torch.float64, device=x.device)
        check(_L().cgb_affine_nc_bwd(_p(x), _p(y), _p(gy), _p(scale), _p(gx), _p(sums), _DT[x.dtype], n, h * w, c, act, slope, _st()),
              "affine_nc_bwd")
        sf = sums.float()
        return gx, sf[..., 1].contiguous(), sf[..., 0].contiguous(), None, None


def affine_nc(x, scale, shift, act=_lib.ACT_NONE, slope=0.2):
    """Per-(sample, channel) affine map + activation (cgb_affine_nc_fwd / _bwd); scale, shift: fp32 [N, Cs]."""
    return _AffineNC.apply(x, scale, shift, act, slope)


class _Moments(Function):
    """Per-(sample, channel) first and second moments E[x], E[x^2] over the pixels of a storage tensor (fp32 [N, Cs] each).
    Backward: gx = (g1 + 2 x g2) / HW — one cgb_affine_nc_fwd pass."""

    @staticmethod
    def forward(ctx, x):
        _chk_storage(x)
        mean, rstd = instnorm_stats(x, 0.0)
        var = torch.where(torch.isfinite(rstd), 1.0 / (rstd * rstd), torch.zeros_like(rstd))   # (a constant channel: var = 0)
        ctx.save_for_backward(x)
        return mean, var + mean * mean

    @staticmethod
    def backward(ctx, g1, g2):
        (x,) = ctx.saved_tensors
        n, h, w, c = x.shape
        inv = 1.0 / (h * w)
        g1 = torch.zeros((n, c), device=x.device) if g1 is None else g1
        g2 = torch.zeros((n, c), device=x.device) if g2 is None else g2
        gx = torch.empty_like(x)
        scale, shift = (2.0 * inv * g2).float().contiguous(), (inv * g1).float().contiguous()
        check(_L().cgb_affine_nc_fwd(_p(x), _p(scale), _p(shift), _p(gx), _DT[x.dtype], n, h * w, c, _lib.ACT_NONE, 0.0, _st()),
              "affine_nc_fwd")
        return gx


def moments(x):
    return _Moments.apply(x)


class _Dropout(Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        _chk_storage(x)
        ctx.meta = (p, seed)
        return _dropout_raw(x, p, seed)

    @staticmethod
    def backward(ctx, g):
        p, seed = ctx.meta
        return _dropout_raw(g.contiguous(), p, seed), None, None


_dropout_calls = [0]


def _draw_dropout_seed():
    return [int(torch.randint(0, 2 ** 62, (1,)).item())]


def dropout(x, p, training):
    """nn.Dropout(p) (deeplab_v2.py:106,148,152).  The keep-mask comes from a counter-based hash of (seed, element index);
    the seed is drawn from torch's CPU generator so torch.manual_seed makes runs reproducible.  While a CUDA graph is being
    captured the seed lives in device memory and is re-drawn, from the same generator, before every replay."""
    if not training or p == 0.0:
        return x
    from . import graphs

    _dropout_calls[0] += 1
    tape = graphs.current_tape()
    seed = _draw_dropout_seed()[0] if tape is None else tape.ints(_draw_dropout_seed, 1)[0]
    return _Dropout.apply(x, float(p), seed)


class _BatchNormAct(Function):
    """act(BatchNorm2d(x) (+ residual)) with batch statistics (train) or running statistics (eval) — one statistics pass
    and one apply pass over x; the backward is two passes (see include/cgb200.h)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, nbt, training, momentum, eps, act, slope, partial=None,
                dual=False):
        _chk_storage(x)
        n, h, w, cs = x.shape
        npix = n * h * w
        c = weight.numel() if weight is not None else (running_mean.numel() if running_mean is not None else cs)
        dev = x.device
        batch_stats = bool(training or running_mean is None)
        wp = bp = None
        if weight is not None:
            if c == cs and weight.dtype == torch.float32 and "aliasparam" not in _DBG:
                wp, bp = weight.detach(), bias.detach()
            else:
                wp = torch.zeros(cs, dtype=torch.float32, device=dev)
                wp[:c] = weight.detach()
                bp = torch.zeros(cs, dtype=torch.float32, device=dev)
                bp[:c] = bias.detach()
        y = torch.empty_like(x)
        res = residual.contiguous() if residual is not None else None
        stats = torch.empty((2, cs), dtype=torch.float32, device=dev)
        mean, rstd = stats[0], stats[1]
        if batch_stats and partial is not None:
            # the statistics came out of the producing conv's epilogue: finalize + running update + ONE apply pass
            assert partial.shape[-1] == cs and partial.dtype == torch.float32, (partial.shape, cs)
            upd = running_mean is not None and training
            if upd:
                _BEPOCH[0] += 1
            check(_L().cgb_bn_train_fwd_partials(_p(x), _p(partial), partial.shape[0], _p(wp), _p(bp), _p(res), _p(y), _p(mean),
                                                 _p(rstd), _p(running_mean) if upd else None, _p(running_var) if upd else None,
                                                 _p(nbt) if upd else None, _DT[x.dtype], npix, cs, c, float(momentum), float(eps),
                                                 act, slope, _st()), "bn_train_fwd_partials")
        elif batch_stats:
            ws = torch.empty((_stats_ws(1, npix, cs),), dtype=torch.float64, device=dev)
            upd = running_mean is not None and training
            if upd:
                _BEPOCH[0] += 1   # folded eval-mode packings (fold_bn) built from the running statistics are now stale
            check(_L().cgb_bn_train_fwd(_p(x), _p(wp), _p(bp), _p(res), _p(y), _p(mean), _p(rstd), _p(ws),
                                        _p(running_mean) if upd else None, _p(running_var) if upd else None,
                                        _p(nbt) if upd else None, _DT[x.dtype], npix, cs, c, float(momentum), float(eps), act,
                                        slope, _st()), "bn_train_fwd")
        else:
            mean.zero_()
            rstd.fill_(1.0)
            mean[:c] = running_mean
            rstd[:c] = torch.rsqrt(running_var + eps)
            check(_L().cgb_bn_apply_fwd(_p(x), _p(mean), _p(rstd), _p(wp), _p(bp), _p(res), _p(y), _DT[x.dtype], npix, cs, act,
                                        slope, _st()), "bn_apply_fwd")
        ctx.save_for_backward(x, mean, rstd, wp, y if act != _lib.ACT_NONE else None)
        ctx.meta = (act, slope, c, residual is not None, batch_stats)
        ctx.dual = bool(dual)
        if dual:
            # the output feeds TWO consumers (a bottleneck's conv1 and identity branch): hand out two aliases so that the
            # backward receives their gradients separately and sums them inside its first pass (cgb_bn_train_bwd2) —
            # autograd would sum them with a separate pass over the tensor.  An unused alias arrives as None.
            ctx.set_materialize_grads(False)
            return y, y.view_as(y)
        return y

    @staticmethod
    def backward(ctx, gy, gy2=None):
        x, mean, rstd, wp, y = ctx.saved_tensors
        act, slope, c, has_res, batch_stats = ctx.meta
        n, h, w, cs = x.shape
        npix = n * h * w
        if gy is None:
            gy, gy2 = gy2, None
        if gy is None:
            return (None,) * 14
        gy = gy.contiguous()
        if gy2 is not None:
            gy2 = gy2.contiguous()
            if not batch_stats:
                gy, gy2 = gy + gy2, None
        gpre = torch.empty_like(x)
        k = ("bnbwd", npix, cs)
        nd = _WS_CACHE.get(k)
        if nd is None:
            nd = _WS_CACHE[k] = int(_L().cgb_bn_bwd_ws_doubles(npix, cs))
        sums_buf = torch.empty((nd,), dtype=torch.float64, device=x.device)   # sums[cs][2] + per-chunk partials
        sums = sums_buf[: 2 * cs].view(cs, 2)
        gw = gb = gx = None
        if batch_stats:
            gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            if gy2 is not None:
                check(_L().cgb_bn_train_bwd2(_p(x), _p(mean), _p(rstd), _p(wp), _p(y), _p(gy), _p(gy2), _p(gpre), _p(gx), _p(sums),
                                             _DT[x.dtype], npix, cs, act, slope, _st()), "bn_train_bwd2")
            else:
                check(_L().cgb_bn_train_bwd(_p(x), _p(mean), _p(rstd), _p(wp), _p(y), _p(gy), _p(gpre), _p(gx), _p(sums),
                                            _DT[x.dtype], npix, cs, act, slope, _st()), "bn_train_bwd")
        else:
            check(_L().cgb_bn_apply_bwd(_p(x), _p(mean), _p(rstd), _p(y), _p(gy), _p(gpre), _p(sums), _DT[x.dtype], npix, cs,
                                        act, slope, _st()), "bn_apply_bwd")
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                check(_L().cgb_bn_bwd_finalize(_p(x), _p(mean), _p(rstd), _p(wp), _p(torch.zeros_like(sums)), _p(gpre), _p(gx),
                                               _DT[x.dtype], npix, cs, _st()), "bn_bwd_finalize")
        if wp is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            sf = sums[:c].float()
            gw = sf[:, 1] if ctx.needs_input_grad[1] else None
            gb = sf[:, 0] if ctx.needs_input_grad[2] else None
        return gx, gw, gb, (gpre if has_res else None), None, None, None, None, None, None, None, None, None, None


def batchnorm_act(x, bn, residual=None, act=_lib.ACT_NONE, slope=0.2, partial=None, dual=False):
    """``act(bn(x) (+ residual))`` for an ``nn.BatchNorm2d`` parameter container ``bn`` (train: batch statistics + running
    update, exactly F.batch_norm's semantics; eval: running statistics)."""
    momentum = 0.1 if bn.momentum is None else bn.momentum
    nbt = bn.num_batches_tracked if (bn.training and bn.track_running_stats) else None   # incremented inside the kernel
    return _BatchNormAct.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, nbt, bn.training, momentum,
                               bn.eps, act, slope, partial, dual)


class _MakeMCond(Function):
    """generator.py:196-230 on storage tensors; differentiable w.r.t. d and s (gen.m.spade.detach is false by default,
    defaults.yaml:182), not w.r.t. the resized image."""

    @staticmethod
    def forward(ctx, d, s, xr, ns):
        _chk_storage(d)
        _chk_storage(s)
        n, h, w, ss = s.shape
        c_out = 1 + ns + (3 if xr is not None else 0)
        cs_out = round8(c_out)
        mm = torch.empty((n, 2), dtype=torch.float32, device=d.device)
        out = torch.empty((n, h, w, cs_out), dtype=d.dtype, device=d.device)
        check(_L().cgb_make_m_cond(_p(d), _p(s), _p(xr), _p(mm), _p(out), _DT[d.dtype], n, h * w, ss, ns, cs_out, _st()),
              "make_m_cond")
        ctx.save_for_backward(d, out, mm)
        ctx.meta = (ss, ns, cs_out)
        return out

    @staticmethod
    def backward(ctx, gout):
        d, out, mm = ctx.saved_tensors
        ss, ns, cs_out = ctx.meta
        n, h, w, _ = out.shape
        gout = gout.contiguous()
        gd = torch.empty_like(d)
        gs = torch.empty((n, h, w, ss), dtype=out.dtype, device=out.device)
        check(_L().cgb_make_m_cond_bwd(_p(d), _p(out), _p(mm), _p(gout), _p(gd), _p(gs), _DT[out.dtype], n, h * w, ss, ns, cs_out,
                                       _st()), "make_m_cond_bwd")
        return gd, gs, None, None


def make_m_cond(d, s, xr, ns):
    """generator.py:196-230: cat[normalize(d), softmax(s, dim=1), x bilinear-resized] -> [N,H,W,round8(1+ns+3)]."""
    return _MakeMCond.apply(d, s, xr.detach() if xr is not None else None, ns)


# ------------------------------------------------------------------------------------------------
# masker losses (NCHW fp32 tensors, as Trainer.masker_{d,s,m}_loss receive them; trainer.py:1389-1616)
# ------------------------------------------------------------------------------------------------
def _f32c(t):
    return t.contiguous().float()


class _SoftmaxNCHW(Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        check(_L().cgb_softmax_nchw_fwd(_p(x), _p(y), n, c, h * w, _st()), "softmax_nchw_fwd")
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        n, c, h, w = y.shape
        gx = torch.empty_like(y)
        check(_L().cgb_softmax_nchw_bwd(_p(y), _p(_f32c(gy)), _p(gx), n, c, h * w, _st()), "softmax_nchw_bwd")
        return gx


def softmax_nchw(x):
    """torch.softmax(x, dim=1) (trainer.py:1449,1475)."""
    return _SoftmaxNCHW.apply(x)


class _FusedLoss(Function):
    """Loss kernels that produce the scalar and the gradient w.r.t. their first argument in one pass."""

    @staticmethod
    def forward(ctx, x, run):
        x = _f32c(x)
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        run(x, loss, gx)
        ctx.save_for_backward(gx)
        return loss

    @staticmethod
    def backward(ctx, g):
        (gx,) = ctx.saved_tensors
        return (gx * g if gx is not None else None), None


def cross_entropy_nchw(logits, target):
    """nn.CrossEntropyLoss()(logits [N,C,H,W], target [N,H,W] int64) (losses.py:106-112)."""
    n, c, h, w = logits.shape
    tgt = target.contiguous().long()
    assert tgt.shape == (n, h, w), (tgt.shape, logits.shape)

    def run(x, loss, gx):
        check(_L().cgb_cross_entropy_nchw(_p(x), _p(tgt), _p(loss), _p(gx), n, c, h * w, _st()), "cross_entropy_nchw")

    return _FusedLoss.apply(logits, run)


def minent_loss(prob, version=1, lambda_var=0.1):
    """MinentLoss (losses.py:177-196) on a probability map [N,C,H,W]."""
    n, c, h, w = prob.shape

    def run(x, loss, gx):
        acc = torch.empty((1,), dtype=torch.float64, device=x.device)
        check(_L().cgb_minent_loss(_p(x), _p(loss), _p(gx), _p(acc), n, c, h * w, version, float(lambda_var), _st()), "minent_loss")

    return _FusedLoss.apply(prob, run)


def tv_loss(x):
    """TVLoss(tvloss_weight=1) (losses.py:140-171)."""
    n, c, h, w = x.shape

    def run(xx, loss, gx):
        check(_L().cgb_tv_loss(_p(xx), _p(loss), _p(gx), n, c, h, w, _st()), "tv_loss")

    return _FusedLoss.apply(x, run)


def bce_logits_loss(x, target):
    """nn.BCEWithLogitsLoss()(x, target) with a tensor target (losses.py:419; trainer.py:1550)."""
    tgt = _f32c(target.detach())
    assert tgt.shape == x.shape, (tgt.shape, x.shape)

    def run(xx, loss, gx):
        check(_L().cgb_bce_logits_loss(_p(xx), _p(tgt), _p(loss), _p(gx), xx.numel(), _st()), "bce_logits_loss")

    return _FusedLoss.apply(x, run)


def ground_intersection_loss(pred, ground):
    """GroundIntersectionLoss (losses.py:449-455): mean(1.0 * ((ground - pred) > 0.5)); piecewise constant (no gradient)."""
    p, g = _f32c(pred.detach()), _f32c(ground.detach())
    loss = torch.zeros((), dtype=torch.float32, device=p.device)
    check(_L().cgb_ground_intersection_loss(_p(p), _p(g), _p(loss), p.numel(), _st()), "ground_intersection_loss")
    return loss


def sigm_loss(pred, target, gmweight=0.5, scales=4):
    """SIGMLoss(gmweight, scale=4) (losses.py:232-278) on depth maps [N,1,H,W]."""
    n, c, h, w = pred.shape
    assert c == 1 and tuple(target.shape) == tuple(pred.shape), (pred.shape, target.shape)
    tgt = _f32c(target.detach())

    def run(x, loss, gx):
        ws = torch.empty((16 + 2 * x.numel(),), dtype=torch.float32, device=x.device)
        if gx is None:
            gx = torch.empty_like(x)
        check(_L().cgb_sigm_loss(_p(x), _p(tgt), _p(loss), _p(gx), _p(ws), n, h, w, float(gmweight), scales, _st()), "sigm_loss")

    return _FusedLoss.apply(pred, run)


def dada_depth_loss(pred, target):
    """DADADepthLoss (losses.py:596-620): reverse Huber with the batch-wide threshold 0.2 * max|pred - target|."""
    assert tuple(pred.shape) == tuple(target.shape) and pred.shape[1] == 1, (pred.shape, target.shape)
    tgt = _f32c(target.detach())

    def run(x, loss, gx):
        check(_L().cgb_dada_depth_loss(_p(x), _p(tgt), _p(loss), _p(gx), x.numel(), _st()), "dada_depth_loss")

    return _FusedLoss.apply(pred, run)


class _DiffAug(Function):
    @staticmethod
    def forward(ctx, x, params, cut_h, cut_w):
        x = _f32c(x)
        n, c, h, w = x.shape
        sums = torch.zeros(n, dtype=torch.float64, device=x.device)
        y = torch.empty_like(x)
        L = _L()
        check(L.cgb_diff_aug_sum(_p(x), _p(params), _p(sums), n, c, h, w, cut_h, cut_w, 0, _st()), "diff_aug_sum")
        check(L.cgb_diff_aug_fwd(_p(x), _p(params), _p(sums), _p(y), n, c, h, w, cut_h, cut_w, _st()), "diff_aug_fwd")
        ctx.save_for_backward(params)
        ctx.cut = (cut_h, cut_w)
        return y

    @staticmethod
    def backward(ctx, gy):
        (params,) = ctx.saved_tensors
        cut_h, cut_w = ctx.cut
        gy = _f32c(gy)
        n, c, h, w = gy.shape
        gsums = torch.zeros(n, dtype=torch.float64, device=gy.device)
        gx = torch.empty_like(gy)
        L = _L()
        check(L.cgb_diff_aug_sum(_p(gy), _p(params), _p(gsums), n, c, h, w, cut_h, cut_w, 1, _st()), "diff_aug_sum")
        check(L.cgb_diff_aug_bwd(_p(gy), _p(params), _p(gsums), _p(gx), n, c, h, w, cut_h, cut_w, _st()), "diff_aug_bwd")
        return gx, None, None, None


def diff_aug(x, params, cut_h=0, cut_w=0):
    """One differentiable augmentation of an NCHW fp32 image batch (DiffTransforms, transforms.py:493-626) from per-sample draws
    ``params`` [N, 8] on the device — see include/cgb200.h for the row layout.  Three launches forward, two backward."""
    _lib.require_device()
    if not _on_device(x):
        raise _lib.CgbError("climategan_b200 tensors must live on a CUDA device (no CPU path)")
    if x.dim() != 4 or x.shape[1] > 8 or tuple(params.shape) != (x.shape[0], 8) or params.dtype != torch.float32:
        raise ValueError(f"diff_aug: x [N,C<=8,H,W] and params [N,8] fp32 expected, got {tuple(x.shape)} / {tuple(params.shape)}")
    return _DiffAug.apply(x, params.contiguous(), int(cut_h), int(cut_w))


def argmax_confusion(pred, label):
    """Confusion matrix of argmax(pred, dim=1) against an integer label, on the device, in one launch — what accuracy and mIOU
    (eval_metrics.py:68-130) are functions of.  pred [N,C,H,W] float, label [N,H,W] / [N,1,H,W] integer-valued.
    Returns (conf int64 [C, C+1] — row = predicted class, column = label, last column = labels outside [0, C) —, label.max())."""
    _lib.require_device()
    if not _on_device(pred):
        raise _lib.CgbError("climategan_b200 tensors must live on a CUDA device (no CPU path)")
    n, c = pred.shape[0], pred.shape[1]
    hw = pred[0, 0].numel()
    lab = label.to(pred.device)
    if lab.numel() != n * hw:
        raise ValueError(f"argmax_confusion: label {tuple(label.shape)} does not match prediction {tuple(pred.shape)}")
    lab = lab.reshape(n, hw).long().contiguous()
    x = _f32c(pred.detach())
    out = torch.zeros(c * (c + 1) + 1, dtype=torch.int64, device=pred.device)
    out[-1] = torch.iinfo(torch.int64).min
    check(_L().cgb_argmax_confusion(_p(x), _p(lab), _p(out), _p(out[-1:]), n, c, hw, _st()), "argmax_confusion")
    return out[:-1].view(c, c + 1), out[-1]


class _EntropyNCHW(Function):
    @staticmethod
    def forward(ctx, prob, depth):
        prob = _f32c(prob)
        n, c, h, w = prob.shape
        d = None if depth is None else _f32c(depth.detach())
        if d is not None:
            assert d.shape == (n, 1, h, w), (d.shape, prob.shape)
        out = torch.empty_like(prob)
        check(_L().cgb_entropy_nchw(_p(prob), _p(d), None, _p(out), n, c, h * w, 0, _st()), "entropy_nchw")
        ctx.save_for_backward(prob, d)
        return out

    @staticmethod
    def backward(ctx, ge):
        prob, d = ctx.saved_tensors
        n, c, h, w = prob.shape
        gp = torch.empty_like(prob)
        check(_L().cgb_entropy_nchw(_p(prob), _p(d), _p(_f32c(ge)), _p(gp), n, c, h * w, 1, _st()), "entropy_nchw_bwd")
        return gp, None


def prob_2_entropy(prob, depth=None):
    """prob_2_entropy(prob) [* depth] (losses.py:466-471, 541-543): the AdvEnt discriminator's input."""
    return _EntropyNCHW.apply(prob, depth)


class _SigmoidPair(Function):
    @staticmethod
    def forward(ctx, logits):
        logits = _f32c(logits)
        n, c, h, w = logits.shape
        assert c == 1
        out = torch.empty((n, 2, h, w), dtype=torch.float32, device=logits.device)
        check(_L().cgb_sigmoid_pair(_p(logits), None, _p(out), n, h * w, 0, _st()), "sigmoid_pair")
        ctx.save_for_backward(logits)
        return out

    @staticmethod
    def backward(ctx, g):
        (logits,) = ctx.saved_tensors
        n, _, h, w = logits.shape
        gl = torch.empty_like(logits)
        check(_L().cgb_sigmoid_pair(_p(logits), _p(_f32c(g)), _p(gl), n, h * w, 1, _st()), "sigmoid_pair_bwd")
        return gl


def sigmoid_pair(logits):
    """cat([sigmoid(l), 1 - sigmoid(l)], dim=1) (trainer.py:1532-1534) for mask logits [N,1,H,W]."""
    return _SigmoidPair.apply(logits)
