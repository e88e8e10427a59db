"""CPU emulation of a SUBSET of the libcgb200 C ABI, for tests only: the entry points of the painter's forward + backward
(convolutions, instance norm, SPADE modulation, nearest resizes, layout edges, paste, L1, spectral norm) written with plain
PyTorch fp32 ops on the raw pointers the product passes, plus the discriminator / VGG / GAN-loss / ExtraAdam entry points of
the painter train step.  With it the product's Python layer — weight packing, the fused
gamma||beta packing, every autograd Function of that path — is checked NUMERICALLY against the reference goldens on CPU
(tests/test_emulated.py), where the dry-run harness (tests/dryrun.py) only checks shapes and argument types.

It follows the contracts written in include/cgb200.h, one function per entry point, and is deliberately naive.  The product
never imports it; an entry point that is not emulated raises."""
from __future__ import annotations

import contextlib
import ctypes as C
import types

import torch
import torch.nn.functional as F

from climategan_b200 import _lib, ops
from tests.dryrun import NoopLib

_DT = {0: torch.float32, 1: torch.bfloat16}
_CT = {torch.float32: C.c_float, torch.bfloat16: C.c_uint16, torch.float64: C.c_double, torch.int64: C.c_int64}


def _addr(p):
    if p is None:
        return None
    return p.value if hasattr(p, "value") else int(p)


def _t(p, shape, dtype):
    """A writable torch view of the memory at pointer ``p``."""
    a = _addr(p)
    if a is None:
        return None
    n = 1
    for s in shape:
        n *= int(s)
    buf = (_CT[dtype] * n).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=n).view(*[int(s) for s in shape])


def _act(x, act, slope):
    if act == 0:
        return x
    if act == 1:
        return torch.relu(x)
    if act == 2:
        return F.leaky_relu(x, slope)
    if act == 3:
        return torch.tanh(x)
    return torch.sigmoid(x)


def _dact_from_out(y, act, slope):
    if act == 0:
        return torch.ones_like(y)
    if act == 1:
        return (y > 0).to(y.dtype)
    if act == 2:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, slope))
    if act == 3:
        return 1 - y * y
    return y * (1 - y)


class EmuLib(NoopLib):
    def __getattr__(self, name):
        fn = self.__class__.__dict__.get("e_" + name[4:]) if name.startswith("cgb_") else None
        if fn is None:
            base = NoopLib.__getattr__(self, name)
            if name in ("cgb_instnorm_ws_doubles", "cgb_bn_bwd_ws_doubles", "cgb_version", "cgb_last_error", "cgb_device_ok",
                        "cgb_conv2d_uses_tcgen05", "cgb_launch_count"):
                return base
            raise NotImplementedError(f"{name} is not emulated (tests/emulib.py covers the painter path)")
        check = NoopLib.__getattr__(self, name)   # the type-checking stand-in: validates the arguments, counts the call

        def call(*args):
            check(*args)
            fn(self, *args)
            return 0

        return call

    # ---- convolution ----------------------------------------------------------------------------------------------
    @staticmethod
    def _desc(dref):
        return dref._obj if hasattr(dref, "_obj") else dref

    def _conv_operands(self, d, x, w):
        dt = _DT[d.dtype]
        X = _t(x, (d.n, d.hi, d.wi, d.ci), dt).float().permute(0, 3, 1, 2)
        W = _t(w, (d.co, d.kh * d.kw, d.ci), dt).float().view(d.co, d.kh, d.kw, d.ci).permute(0, 3, 1, 2)
        pad = d.pad
        if d.pad_mode == 1 and d.pad > 0:
            X = F.pad(X, (d.pad,) * 4, mode="reflect")
            pad = 0
        return dt, X, W, pad

    def e_conv2d_fwd(self, dref, x, w, bias, residual, y, stream):
        d = self._desc(dref)
        dt, X, W, pad = self._conv_operands(d, x, w)
        b = _t(bias, (d.co,), torch.float32)
        Y = F.conv2d(X, W, b, stride=d.stride, padding=pad, dilation=d.dil).permute(0, 2, 3, 1)
        R = _t(residual, (d.n, d.ho, d.wo, d.co), dt)
        if R is not None and d.res_before_act:
            Y = Y + R.float()
        Y = _act(Y, d.act, d.slope)
        if R is not None and not d.res_before_act:
            Y = Y + R.float()
        _t(y, (d.n, d.ho, d.wo, d.co), dt).copy_(Y)

    def e_conv2d_dgrad(self, dref, gy, w, wt, dact, mask_src, gx, stream):
        d = self._desc(dref)
        dt = _DT[d.dtype]
        assert not (d.pad_mode == 1 and d.pad > 0)
        GY = _t(gy, (d.n, d.ho, d.wo, d.co), dt).float().permute(0, 3, 1, 2)
        W = _t(w, (d.co, d.kh * d.kw, d.ci), dt).float().view(d.co, d.kh, d.kw, d.ci).permute(0, 3, 1, 2)
        GX = torch.nn.grad.conv2d_input((d.n, d.ci, d.hi, d.wi), W, GY, stride=d.stride, padding=d.pad, dilation=d.dil)
        GX = GX.permute(0, 2, 3, 1)
        M = _t(mask_src, (d.n, d.hi, d.wi, d.ci), dt)
        if M is not None and dact != 0:
            GX = GX * _dact_from_out(M.float(), dact, d.slope)
        _t(gx, (d.n, d.hi, d.wi, d.ci), dt).copy_(GX)

    def e_conv2d_wgrad(self, dref, x, gy, gw, gbias, accumulate, stream):
        d = self._desc(dref)
        dt = _DT[d.dtype]
        X = _t(x, (d.n, d.hi, d.wi, d.ci), dt).float().permute(0, 3, 1, 2)
        pad = d.pad
        if d.pad_mode == 1 and d.pad > 0:
            X = F.pad(X, (d.pad,) * 4, mode="reflect")
            pad = 0
        GY = _t(gy, (d.n, d.ho, d.wo, d.co), dt).float().permute(0, 3, 1, 2)
        GW = torch.nn.grad.conv2d_weight(X, (d.co, d.ci, d.kh, d.kw), GY, stride=d.stride, padding=pad, dilation=d.dil)
        GW = GW.permute(0, 2, 3, 1).reshape(d.co, d.kh * d.kw, d.ci)
        out = _t(gw, (d.co, d.kh * d.kw, d.ci), torch.float32)
        out.copy_(out + GW if accumulate else GW)
        gb = _t(gbias, (d.co,), torch.float32)
        if gb is not None:
            s = GY.sum((0, 2, 3))
            gb.copy_(gb + s if accumulate else s)

    def e_im2col(self, x, y, dtype, n, h, w, cs_in, c, k, pad, dil, cs_out, stream):
        dt = _DT[dtype]
        X = _t(x, (n, h, w, cs_in), dt).float()[..., :c].permute(0, 3, 1, 2)
        cols = F.unfold(X, k, dilation=dil, padding=pad)                    # [n, c*k*k, h*w], channel-major then tap
        cols = cols.view(n, c, k * k, h * w).permute(0, 3, 2, 1).reshape(n, h, w, k * k * c)   # tap-major, channel minor
        Y = _t(y, (n, h, w, cs_out), dt)
        Y.zero_()
        Y[..., : k * k * c] = cols

    # ---- norms ----------------------------------------------------------------------------------------------------
    def e_instnorm_stats(self, x, dtype, n, hw, c, eps, ws, mean, rstd, stream):
        X = _t(x, (n, hw, c), _DT[dtype]).double()
        m = X.mean(1)
        v = X.var(1, unbiased=False)
        _t(mean, (n, c), torch.float32).copy_(m)
        _t(rstd, (n, c), torch.float32).copy_(1.0 / torch.sqrt(v + eps))

    def e_spade_modulate_fwd(self, x, mean, rstd, gb, out, dtype, n, hw, c, act, slope, stream):
        dt = _DT[dtype]
        X = _t(x, (n, hw, c), dt).float()
        xhat = (X - _t(mean, (n, 1, c), torch.float32)) * _t(rstd, (n, 1, c), torch.float32)
        GB = _t(gb, (n, hw, 2 * c), dt).float()
        _t(out, (n, hw, c), dt).copy_(_act(xhat * (1 + GB[..., :c]) + GB[..., c:], act, slope))

    def e_spade_modulate_bwd(self, x, mean, rstd, gb, gout, ggb, gxhat, sums, dtype, n, hw, c, act, slope, stream):
        dt = _DT[dtype]
        X = _t(x, (n, hw, c), dt).float()
        xhat = (X - _t(mean, (n, 1, c), torch.float32)) * _t(rstd, (n, 1, c), torch.float32)
        GB = _t(gb, (n, hw, 2 * c), dt).float()
        pre = xhat * (1 + GB[..., :c]) + GB[..., c:]
        if act == 0:
            der = torch.ones_like(pre)
        elif act == 1:
            der = (pre > 0).float()
        else:
            der = torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, slope))
        gs = _t(gout, (n, hw, c), dt).float() * der
        _t(ggb, (n, hw, 2 * c), dt).copy_(torch.cat([gs * xhat, gs], -1))
        gxh = gs * (1 + GB[..., :c])
        _t(gxhat, (n, hw, c), dt).copy_(gxh)
        S = _t(sums, (n, c, 2), torch.float64)
        S[..., 0] += gxh.double().sum(1)
        S[..., 1] += (gxh * xhat).double().sum(1)

    def e_instnorm_bwd(self, x, mean, rstd, sums, g, dtype, n, hw, c, stream):
        dt = _DT[dtype]
        X = _t(x, (n, hw, c), dt).float()
        r = _t(rstd, (n, 1, c), torch.float32)
        xhat = (X - _t(mean, (n, 1, c), torch.float32)) * r
        S = _t(sums, (n, c, 2), torch.float64).float()
        G = _t(g, (n, hw, c), dt)
        G.copy_(r * (G.float() - S[:, None, :, 0] / hw - xhat * S[:, None, :, 1] / hw))

    # ---- resampling / elementwise / layout ---------------------------------------------------------------------------
    def e_resize_nearest_fwd(self, x, y, dtype, n, hi, wi, ho, wo, c, stream):
        dt = _DT[dtype]
        X = _t(x, (n, hi, wi, c), dt).float().permute(0, 3, 1, 2)
        _t(y, (n, ho, wo, c), dt).copy_(F.interpolate(X, size=(ho, wo), mode="nearest").permute(0, 2, 3, 1))

    def e_upsample_nearest_bwd(self, gy, gx, dtype, n, hi, wi, f, c, stream):
        dt = _DT[dtype]
        G = _t(gy, (n, hi, f, wi, f, c), dt).float()
        _t(gx, (n, hi, wi, c), dt).copy_(G.sum((2, 4)))

    def e_act_fwd(self, x, y, dtype, count, act, slope, stream):
        dt = _DT[dtype]
        _t(y, (count,), dt).copy_(_act(_t(x, (count,), dt).float(), act, slope))

    def e_act_bwd(self, gy, y, gx, dtype, count, act, slope, stream):
        dt = _DT[dtype]
        _t(gx, (count,), dt).copy_(_t(gy, (count,), dt).float() * _dact_from_out(_t(y, (count,), dt).float(), act, slope))

    def e_nchw_to_nhwc(self, x, y, dtype, n, c, hw, cs, stream):
        Y = _t(y, (n, hw, cs), _DT[dtype])
        Y.zero_()
        Y[..., :c] = _t(x, (n, c, hw), torch.float32).permute(0, 2, 1)

    def e_nhwc_to_nchw(self, x, y, dtype, n, c, hw, cs, stream):
        _t(y, (n, c, hw), torch.float32).copy_(_t(x, (n, hw, cs), _DT[dtype]).float()[..., :c].permute(0, 2, 1))

    def e_mask_cond(self, x, m, cond, dtype, n, hw, cs, stream):
        Cd = _t(cond, (n, hw, cs), _DT[dtype])
        Cd.zero_()
        Cd[..., :3] = (_t(x, (n, 3, hw), torch.float32) * (1 - _t(m, (n, 1, hw), torch.float32))).permute(0, 2, 1)

    def e_paste_fwd(self, x, m, fake, out, n, hw, stream):
        M = _t(m, (n, 1, hw), torch.float32)
        _t(out, (n, 3, hw), torch.float32).copy_(_t(x, (n, 3, hw), torch.float32) * (1 - M) + _t(fake, (n, 3, hw), torch.float32) * M)

    def e_paste_bwd(self, gout, m, gfake, n, hw, stream):
        _t(gfake, (n, 3, hw), torch.float32).copy_(_t(gout, (n, 3, hw), torch.float32) * _t(m, (n, 1, hw), torch.float32))

    def e_l1_loss(self, a, b, loss, ga, count, scale, stream):
        d = _t(a, (count,), torch.float32) - _t(b, (count,), torch.float32)
        _t(loss, (1,), torch.float32).add_(scale * d.abs().sum())
        G = _t(ga, (count,), torch.float32)
        if G is not None:
            G.copy_(scale * torch.sign(d))

    def e_spectral_power_iter(self, w, u, v, sigma, rows, cols, stream):
        W = _t(w, (rows, cols), torch.float32)
        U, V = _t(u, (rows,), torch.float32), _t(v, (cols,), torch.float32)
        nv = torch.mv(W.t(), U)
        V.copy_(nv / (nv.norm() + 1e-12))
        wv = torch.mv(W, V)
        U.copy_(wv / (wv.norm() + 1e-12))
        _t(sigma, (1,), torch.float32).copy_(U.dot(wv).reshape(1))


    # ---- discriminator / VGG / losses / optimiser (the painter train step) --------------------------------------------------
    def e_instnorm_apply_fwd(self, x, mean, rstd, y, dtype, n, hw, c, act, slope, stream):
        dt = _DT[dtype]
        xhat = (_t(x, (n, hw, c), dt).float() - _t(mean, (n, 1, c), torch.float32)) * _t(rstd, (n, 1, c), torch.float32)
        _t(y, (n, hw, c), dt).copy_(_act(xhat, act, slope))

    def e_instnorm_apply_bwd(self, x, mean, rstd, gy, gxhat, sums, dtype, n, hw, c, act, slope, stream):
        dt = _DT[dtype]
        xhat = (_t(x, (n, hw, c), dt).float() - _t(mean, (n, 1, c), torch.float32)) * _t(rstd, (n, 1, c), torch.float32)
        if act == 0:
            der = torch.ones_like(xhat)
        elif act == 1:
            der = (xhat > 0).float()
        else:
            der = torch.where(xhat > 0, torch.ones_like(xhat), torch.full_like(xhat, slope))
        g = _t(gy, (n, hw, c), dt).float() * der
        _t(gxhat, (n, hw, c), dt).copy_(g)
        S = _t(sums, (n, c, 2), torch.float64)
        S[..., 0] += g.double().sum(1)
        S[..., 1] += (g * xhat).double().sum(1)

    def e_avgpool3s2_fwd(self, x, y, dtype, n, hi, wi, c, stream):
        dt = _DT[dtype]
        ho, wo = (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
        X = _t(x, (n, hi, wi, c), dt).float().permute(0, 3, 1, 2)
        _t(y, (n, ho, wo, c), dt).copy_(F.avg_pool2d(X, 3, 2, 1, count_include_pad=False).permute(0, 2, 3, 1))

    def e_avgpool3s2_bwd(self, gy, gx, dtype, n, hi, wi, c, stream):
        dt = _DT[dtype]
        ho, wo = (hi - 1) // 2 + 1, (wi - 1) // 2 + 1
        with torch.enable_grad():   # (called from inside an autograd backward, where grad mode is off)
            X = torch.zeros(n, c, hi, wi, requires_grad=True)
            Y = F.avg_pool2d(X, 3, 2, 1, count_include_pad=False)
            (gX,) = torch.autograd.grad(Y, X, _t(gy, (n, ho, wo, c), dt).float().permute(0, 3, 1, 2))
        _t(gx, (n, hi, wi, c), dt).copy_(gX.permute(0, 2, 3, 1))

    def e_const_target_loss(self, x, loss, gx, count, kind, target, scale, stream):
        v = _t(x, (count,), torch.float32)
        if kind == 0:
            l = torch.clamp(v, min=0) - v * target + torch.log1p(torch.exp(-v.abs()))
            g = torch.sigmoid(v) - target
        elif kind == 1:
            l, g = (v - target) ** 2, 2 * (v - target)
        elif kind == 2:
            l, g = -torch.clamp(v - 1, max=0), torch.where(v < 1, -torch.ones_like(v), torch.zeros_like(v))
        elif kind == 3:
            l, g = -torch.clamp(-v - 1, max=0), torch.where(v > -1, torch.ones_like(v), torch.zeros_like(v))
        else:
            l, g = -v, -torch.ones_like(v)
        _t(loss, (1,), torch.float32).add_(scale * l.sum())
        G = _t(gx, (count,), torch.float32)
        if G is not None:
            G.copy_(g * scale)

    def e_l1_loss_storage(self, a, b, loss, ga, dtype, count, scale, stream):
        dt = _DT[dtype]
        d = _t(a, (count,), dt).float() - _t(b, (count,), dt).float()
        _t(loss, (1,), torch.float32).add_(scale * d.abs().sum())
        G = _t(ga, (count,), dt)
        if G is not None:
            G.copy_(scale * torch.sign(d))

    def e_vgg_preprocess_fwd(self, x, m, y, dtype, n, hw, stream):
        X = _t(x, (n, 3, hw), torch.float32)
        M = _t(m, (n, 1, hw), torch.float32)
        X = X * M if M is not None else X
        mean = torch.tensor([103.939, 116.779, 123.680]).view(1, 3, 1)
        Y = _t(y, (n, hw, 8), _DT[dtype])
        Y.zero_()
        Y[..., :3] = ((X.flip(1) + 1) * 127.5 - mean).permute(0, 2, 1)

    def e_vgg_preprocess_bwd(self, gy, m, gx, dtype, n, hw, stream):
        G = _t(gy, (n, hw, 8), _DT[dtype]).float()[..., :3].permute(0, 2, 1).flip(1) * 127.5
        M = _t(m, (n, 1, hw), torch.float32)
        _t(gx, (n, 3, hw), torch.float32).copy_(G * M if M is not None else G)

    def e_maxpool2_fwd(self, x, y, dtype, n, hi, wi, c, stream):
        dt = _DT[dtype]
        X = _t(x, (n, hi, wi, c), dt).float().permute(0, 3, 1, 2)
        _t(y, (n, hi // 2, wi // 2, c), dt).copy_(F.max_pool2d(X, 2, 2).permute(0, 2, 3, 1))

    def e_maxpool2_bwd(self, x, y, gy, gx, dtype, n, hi, wi, c, stream):
        dt = _DT[dtype]
        with torch.enable_grad():
            X = _t(x, (n, hi, wi, c), dt).float().permute(0, 3, 1, 2).clone().requires_grad_(True)
            (gX,) = torch.autograd.grad(F.max_pool2d(X, 2, 2), X, _t(gy, (n, hi // 2, wi // 2, c), dt).float().permute(0, 3, 1, 2))
        _t(gx, (n, hi, wi, c), dt).copy_(gX.permute(0, 2, 3, 1))

    def e_extra_adam(self, p, g, m, v, c, count, lr, b1, b2, eps, wd, step, mode, save_copy, stream):
        P, G = _t(p, (count,), torch.float32), _t(g, (count,), torch.float32)
        M, V, Cc = _t(m, (count,), torch.float32), _t(v, (count,), torch.float32), _t(c, (count,), torch.float32)
        gi = G + wd * P if wd != 0 else G
        M.copy_(b1 * M + (1 - b1) * gi)
        V.copy_(b2 * V + (1 - b2) * gi * gi)
        step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
        u = -step_size * M / (V.sqrt() + eps)
        if mode == 0:
            if save_copy:
                Cc.copy_(P)
            P.add_(u)
        else:
            P.copy_(Cc + u)


@contextlib.contextmanager
def emulated_library():
    real = _lib.lib()
    fake = EmuLib(real)
    saved = (_lib._lib, ops._on_device, torch.cuda.current_stream)
    _lib._lib = fake
    ops._on_device = lambda x: True
    torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0)
    try:
        yield fake
    finally:
        _lib._lib, ops._on_device, torch.cuda.current_stream = saved
        ops.invalidate_weight_cache()
