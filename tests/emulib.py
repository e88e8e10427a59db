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

_DT = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}
_CT = {torch.float32: C.c_float, torch.bfloat16: C.c_uint16, torch.float16: C.c_uint16, torch.float64: C.c_double, torch.int64: C.c_int64}


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
    if act == 5:
        return F.selu(x)
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
    if act == 5:
        sc, al = 1.0507009873554804934193349852946, 1.6732632423543772848170429916717
        return torch.where(y > 0, torch.full_like(y, sc), y + sc * al)
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

    def e_pack_weight(self, w, wp, dtype, o, i, taps, cos, cis, stream):
        W = _t(w, (o, i, taps), torch.float32)
        P = _t(wp, (cos, taps, cis), _DT[dtype])
        P.zero_()
        P[:o, :, :i] = W.permute(0, 2, 1)

    def e_pack_weight_dual(self, w, wp, wt, dtype, o, i, taps, cos, cis, stream):
        self.e_pack_weight(w, wp, dtype, o, i, taps, cos, cis, stream)
        W = _t(w, (o, i, taps), torch.float32)
        T = _t(wt, (cis, taps, cos), _DT[dtype])
        T.zero_()
        T[:i, :, :o] = W.flip(2).permute(1, 2, 0)      # wt[ci][taps-1-t][co] = w[co][ci][t]

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

    def e_spade_modulate_bwd_bias(self, x, mean, rstd, gb, gout, ggb, gxhat, sums, bsum, dtype, n, hw, c, act, slope, stream):
        self.e_spade_modulate_bwd(x, mean, rstd, gb, gout, ggb, gxhat, sums, dtype, n, hw, c, act, slope, stream)
        # (the kernel sums the fp32 values before they are rounded to the storage type; the emulation sums what was stored —
        #  identical in fp32 storage, within bf16 rounding otherwise)
        B = _t(bsum, (2 * c,), torch.float64)
        B += _t(ggb, (n, hw, 2 * c), _DT[dtype]).double().sum((0, 1))

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

    def e_act_bwd_bias(self, gy, y, gx, gbias, dtype, npix, c, act, slope, stream):
        dt = _DT[dtype]
        o = _t(gy, (npix, c), dt).float() * _dact_from_out(_t(y, (npix, c), dt).float(), act, slope)
        _t(gx, (npix, c), dt).copy_(o)
        _t(gbias, (c,), torch.float32).copy_(o.sum(0))

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

    def e_const_target_loss_dev(self, x, loss, gx, count, kind, target_dev, scale, stream):
        self.e_const_target_loss(x, loss, gx, count, kind, float(_t(target_dev, (1,), torch.float32)[0]), scale, stream)

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


    # ---- masker training path -------------------------------------------------------------------------------------------
    def e_bn_train_fwd(self, x, weight, bias, residual, y, mean, rstd, ws, running_mean, running_var, nbt, dtype, npix, c,
                       c_logical, momentum, eps, act, slope, stream):
        dt = _DT[dtype]
        X = _t(x, (npix, c), dt).double()
        m, v = X.mean(0), X.var(0, unbiased=False)
        Mn, Rs = _t(mean, (c,), torch.float32), _t(rstd, (c,), torch.float32)
        Mn.copy_(m)
        Rs.copy_(1.0 / torch.sqrt(v + eps))
        rm, rv = _t(running_mean, (c_logical,), torch.float32), _t(running_var, (c_logical,), torch.float32)
        if rm is not None and rv is not None:
            unb = v[:c_logical] * npix / (npix - 1) if npix > 1 else v[:c_logical]
            rm.copy_((1 - momentum) * rm + momentum * m[:c_logical].float())
            rv.copy_((1 - momentum) * rv + momentum * unb.float())
            nb = _t(nbt, (1,), torch.int64)
            if nb is not None:
                nb.add_(1)
        self.e_bn_apply_fwd(x, mean, rstd, weight, bias, residual, y, dtype, npix, c, act, slope, stream)

    def e_bn_apply_fwd(self, x, mean, rstd, weight, bias, residual, y, dtype, npix, c, act, slope, stream):
        dt = _DT[dtype]
        xhat = (_t(x, (npix, c), dt).float() - _t(mean, (c,), torch.float32)) * _t(rstd, (c,), torch.float32)
        W, B = _t(weight, (c,), torch.float32), _t(bias, (c,), torch.float32)
        pre = xhat * W + B if W is not None else xhat
        R = _t(residual, (npix, c), dt)
        if R is not None:
            pre = pre + R.float()
        _t(y, (npix, c), dt).copy_(_act(pre, act, slope))

    def e_bn_train_bwd2(self, x, mean, rstd, weight, y, gy, gy2, gpre, gx, sums, dtype, npix, c, act, slope, stream):
        if _addr(gy2) is None:
            return self.e_bn_train_bwd(x, mean, rstd, weight, y, gy, gpre, gx, sums, dtype, npix, c, act, slope, stream)
        dt = _DT[dtype]
        tot = (_t(gy, (npix, c), dt).float() + _t(gy2, (npix, c), dt).float()).to(dt).contiguous()
        self._keep = tot                                   # keep the temporary alive for the call
        return self.e_bn_train_bwd(x, mean, rstd, weight, y, tot.data_ptr(), gpre, gx, sums, dtype, npix, c, act, slope, stream)

    def e_bn_train_bwd(self, x, mean, rstd, weight, y, gy, gpre, gx, sums, dtype, npix, c, act, slope, stream):
        dt = _DT[dtype]
        r = _t(rstd, (c,), torch.float32)
        xhat = (_t(x, (npix, c), dt).float() - _t(mean, (c,), torch.float32)) * r
        gp = _t(gy, (npix, c), dt).float()
        if act != 0:
            gp = gp * _dact_from_out(_t(y, (npix, c), dt).float(), act, slope)
        _t(gpre, (npix, c), dt).copy_(gp)
        S = _t(sums, (c, 2), torch.float64)
        S[:, 0] = gp.double().sum(0)
        S[:, 1] = (gp * xhat).double().sum(0)
        GX = _t(gx, (npix, c), dt)
        if GX is not None:
            W = _t(weight, (c,), torch.float32)
            k = r * W if W is not None else r
            GX.copy_(k * (gp - S[:, 0].float() / npix - xhat * S[:, 1].float() / npix))

    def e_bn_update_running(self, mean, rstd, running_mean, running_var, c, count, momentum, eps, stream):
        m, r = _t(mean, (c,), torch.float32), _t(rstd, (c,), torch.float32)
        var = torch.clamp(1.0 / (r.double() ** 2) - eps, min=0)
        unb = var * count / (count - 1) if count > 1 else var
        rm, rv = _t(running_mean, (c,), torch.float32), _t(running_var, (c,), torch.float32)
        rm.copy_((1 - momentum) * rm + momentum * m)
        rv.copy_((1 - momentum) * rv + momentum * unb.float())

    def _nchw(self, p, n, h, w, c, dt):
        return _t(p, (n, h, w, c), dt).float().permute(0, 3, 1, 2)

    def _vjp(self, fn, X, gy):
        with torch.enable_grad():
            X = X.clone().requires_grad_(True)
            (g,) = torch.autograd.grad(fn(X), X, gy)
        return g

    def e_reflect_pad_fwd(self, x, y, dtype, n, h, w, c, pad, stream):
        dt = _DT[dtype]
        _t(y, (n, h + 2 * pad, w + 2 * pad, c), dt).copy_(F.pad(self._nchw(x, n, h, w, c, dt), (pad,) * 4, mode="reflect").permute(0, 2, 3, 1))

    def e_reflect_pad_bwd(self, gy, gx, dtype, n, h, w, c, pad, stream):
        dt = _DT[dtype]
        g = self._vjp(lambda X: F.pad(X, (pad,) * 4, mode="reflect"), torch.zeros(n, c, h, w), self._nchw(gy, n, h + 2 * pad, w + 2 * pad, c, dt))
        _t(gx, (n, h, w, c), dt).copy_(g.permute(0, 2, 3, 1))

    def e_replicate_pad_fwd(self, x, y, dtype, n, h, w, c, pad, stream):
        dt = _DT[dtype]
        _t(y, (n, h + 2 * pad, w + 2 * pad, c), dt).copy_(F.pad(self._nchw(x, n, h, w, c, dt), (pad,) * 4, mode="replicate").permute(0, 2, 3, 1))

    def e_replicate_pad_bwd(self, gy, gx, dtype, n, h, w, c, pad, stream):
        dt = _DT[dtype]
        g = self._vjp(lambda X: F.pad(X, (pad,) * 4, mode="replicate"), torch.zeros(n, c, h, w), self._nchw(gy, n, h + 2 * pad, w + 2 * pad, c, dt))
        _t(gx, (n, h, w, c), dt).copy_(g.permute(0, 2, 3, 1))

    def e_affine_nc_fwd(self, x, scale, shift, y, dtype, n, hw, c, act, slope, stream):
        dt = _DT[dtype]
        X = _t(x, (n, hw, c), dt).float()
        out = _act(X * _t(scale, (n, 1, c), torch.float32) + _t(shift, (n, 1, c), torch.float32), act, slope)
        _t(y, (n, hw, c), dt).copy_(out)

    def e_affine_nc_bwd(self, x, y, gy, scale, gx, sums, dtype, n, hw, c, act, slope, stream):
        dt = _DT[dtype]
        X, G = _t(x, (n, hw, c), dt).float(), _t(gy, (n, hw, c), dt).float()
        if act != 0:
            G = G * _dact_from_out(_t(y, (n, hw, c), dt).float(), act, slope)
        _t(gx, (n, hw, c), dt).copy_(G * _t(scale, (n, 1, c), torch.float32))
        S = _t(sums, (n, c, 2), torch.float64)
        S[..., 0] += G.double().sum(1)
        S[..., 1] += (G.double() * X.double()).sum(1)

    @staticmethod
    def _pool(X, pad, ho):
        hi = X.shape[-2]
        ceil = ho != (hi + 2 * pad - 3) // 2 + 1
        return F.max_pool2d(X, 3, 2, pad, ceil_mode=ceil)

    def e_maxpool3s2_ceil_fwd(self, x, y, dtype, n, hi, wi, ho, wo, c, stream):
        self.e_maxpool3s2_fwd(x, y, dtype, n, hi, wi, ho, wo, c, 0, stream)

    def e_maxpool3s2_ceil_bwd(self, x, gy, gx, dtype, n, hi, wi, ho, wo, c, stream):
        self.e_maxpool3s2_bwd(x, gy, gx, dtype, n, hi, wi, ho, wo, c, 0, stream)

    def e_maxpool3s2_fwd(self, x, y, dtype, n, hi, wi, ho, wo, c, pad, stream):
        dt = _DT[dtype]
        Y = self._pool(self._nchw(x, n, hi, wi, c, dt), pad, ho)
        assert tuple(Y.shape[-2:]) == (ho, wo), (Y.shape, ho, wo)
        _t(y, (n, ho, wo, c), dt).copy_(Y.permute(0, 2, 3, 1))

    def e_maxpool3s2_bwd(self, x, gy, gx, dtype, n, hi, wi, ho, wo, c, pad, stream):
        dt = _DT[dtype]
        g = self._vjp(lambda X: self._pool(X, pad, ho), self._nchw(x, n, hi, wi, c, dt), self._nchw(gy, n, ho, wo, c, dt))
        _t(gx, (n, hi, wi, c), dt).copy_(g.permute(0, 2, 3, 1))

    def e_resize_bilinear_fwd(self, x, y, dtype, n, hi, wi, ho, wo, c, align_corners, stream):
        dt = _DT[dtype]
        Y = F.interpolate(self._nchw(x, n, hi, wi, c, dt), size=(ho, wo), mode="bilinear", align_corners=bool(align_corners))
        _t(y, (n, ho, wo, c), dt).copy_(Y.permute(0, 2, 3, 1))

    def e_resize_bilinear_bwd(self, gy, gx, dtype, n, hi, wi, ho, wo, c, align_corners, stream):
        dt = _DT[dtype]
        g = self._vjp(lambda X: F.interpolate(X, size=(ho, wo), mode="bilinear", align_corners=bool(align_corners)),
                      torch.zeros(n, c, hi, wi), self._nchw(gy, n, ho, wo, c, dt))
        _t(gx, (n, hi, wi, c), dt).copy_(g.permute(0, 2, 3, 1))

    def e_resize_bicubic_fwd(self, x, y, dtype, n, hi, wi, ho, wo, c, stream):
        dt = _DT[dtype]
        Y = F.interpolate(self._nchw(x, n, hi, wi, c, dt), size=(ho, wo), mode="bicubic", align_corners=False)
        _t(y, (n, ho, wo, c), dt).copy_(Y.permute(0, 2, 3, 1))

    def e_resize_nearest_bwd(self, gy, gx, dtype, n, hi, wi, ho, wo, c, stream):
        dt = _DT[dtype]
        g = self._vjp(lambda X: F.interpolate(X, size=(ho, wo), mode="nearest"), torch.zeros(n, c, hi, wi), self._nchw(gy, n, ho, wo, c, dt))
        _t(gx, (n, hi, wi, c), dt).copy_(g.permute(0, 2, 3, 1))

    def e_channel_mean(self, x, y, dtype, pixels, cs, c_logical, stream):
        dt = _DT[dtype]
        Y = _t(y, (pixels, 8), dt)
        Y.zero_()
        Y[:, 0] = _t(x, (pixels, cs), dt).float()[:, :c_logical].mean(1)

    def e_channel_mean_bwd(self, gy, gx, dtype, pixels, cs, c_logical, stream):
        dt = _DT[dtype]
        G = _t(gx, (pixels, cs), dt)
        G.zero_()
        G[:, :c_logical] = (_t(gy, (pixels, 8), dt).float()[:, :1] / c_logical).expand(pixels, c_logical)

    def e_mul(self, a, b, y, dtype, count, stream):
        dt = _DT[dtype]
        _t(y, (count,), dt).copy_(_t(a, (count,), dt).float() * _t(b, (count,), dt).float())

    def e_broadcast_hw(self, src, dst, dtype, n, hw, c, scale, stream):
        dt = _DT[dtype]
        _t(dst, (n, hw, c), dt).copy_((_t(src, (n, 1, c), dt).float() * scale).expand(n, hw, c))

    def e_dropout(self, x, y, dtype, count, p, seed, stream):
        X = _t(x, (count,), _DT[dtype])
        if p == 0:
            _t(y, (count,), _DT[dtype]).copy_(X)
            return
        # a mask that is a function of (seed, element index), like the kernel's counter-based one (not the same stream: the
        # parity fixtures run with p = 0); the same call on the gradient is the backward
        keep = torch.rand(count, generator=torch.Generator().manual_seed(int(seed) % (2 ** 63))) >= p
        _t(y, (count,), _DT[dtype]).copy_(X.float() * keep / (1 - p))

    def e_dropout_dev(self, x, y, dtype, count, p, seed_dev, stream):
        self.e_dropout(x, y, dtype, count, p, int(_t(seed_dev, (1,), torch.int64)[0]), stream)

    def e_im2col_strided(self, x, y, dtype, n, h, w, cs_in, c, k, pad, dil, stride, cs_out, stream):
        dt = _DT[dtype]
        ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
        X = _t(x, (n, h, w, cs_in), dt).float()[..., :c].permute(0, 3, 1, 2)
        cols = F.unfold(X, k, dilation=dil, padding=pad, stride=stride)
        cols = cols.view(n, c, k * k, ho * wo).permute(0, 3, 2, 1).reshape(n, ho, wo, k * k * c)
        Y = _t(y, (n, ho, wo, cs_out), dt)
        Y.zero_()
        Y[..., : k * k * c] = cols

    def e_col2im_strided(self, g, gx, dtype, n, h, w, cs_in, c, k, pad, dil, stride, cs_col, stream):
        dt = _DT[dtype]
        ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
        G = _t(g, (n, ho, wo, cs_col), dt).float()[..., : k * k * c]
        cols = G.reshape(n, ho * wo, k * k, c).permute(0, 3, 2, 1).reshape(n, c * k * k, ho * wo)   # unfold's channel-major order
        X = F.fold(cols, (h, w), k, dilation=dil, padding=pad, stride=stride)                        # [n, c, h, w]
        GX = _t(gx, (n, h, w, cs_in), dt)
        GX.zero_()
        GX[..., :c] = X.permute(0, 2, 3, 1)

    @staticmethod
    def _m_cond(D, S, XR):
        n = D.shape[0]
        mn = D.reshape(n, -1).min(1)[0].reshape(n, 1, 1)
        t = D - mn
        t = t / t.reshape(n, -1).max(1)[0].reshape(n, 1, 1)
        parts = [t, torch.softmax(S, -1)]
        if XR is not None:
            parts.append(XR)
        return torch.cat(parts, -1)

    def e_make_m_cond(self, d, s, xr, mm, out, dtype, n, hw, ss, ns, cs_out, stream):
        dt = _DT[dtype]
        D = _t(d, (n, hw, 8), dt).float()[..., :1]
        S = _t(s, (n, hw, ss), dt).float()[..., :ns]
        XR = _t(xr, (n, hw, 8), dt)
        XR = XR.float()[..., :3] if XR is not None else None
        MM = _t(mm, (n, 2), torch.float32)
        MM[:, 0], MM[:, 1] = D.reshape(n, -1).min(1)[0], D.reshape(n, -1).max(1)[0]
        res = self._m_cond(D, S, XR)
        O = _t(out, (n, hw, cs_out), dt)
        O.zero_()
        O[..., : res.shape[-1]] = res

    def e_make_m_cond_bwd(self, d, out, mm, gout, gd, gs, dtype, n, hw, ss, ns, cs_out, stream):
        dt = _DT[dtype]
        D = _t(d, (n, hw, 8), dt).float()[..., :1]
        # the softmax is re-derived from the forward's output p: gs = p (g - <g, p>)
        P = _t(out, (n, hw, cs_out), dt).float()[..., 1:1 + ns]
        G = _t(gout, (n, hw, cs_out), dt).float()
        g0, gp = G[..., :1], G[..., 1:1 + ns]
        with torch.enable_grad():
            Dg = D.clone().requires_grad_(True)
            mn = Dg.reshape(n, -1).min(1)[0].reshape(n, 1, 1)
            t = Dg - mn
            t = t / t.reshape(n, -1).max(1)[0].reshape(n, 1, 1)
            (gD,) = torch.autograd.grad(t, Dg, g0)
        GD = _t(gd, (n, hw, 8), dt)
        GD.zero_()
        GD[..., :1] = gD
        GS = _t(gs, (n, hw, ss), dt)
        GS.zero_()
        GS[..., :ns] = P * (gp - (gp * P).sum(-1, keepdim=True))

    def e_mask_cond_bwd(self, x, gcond, gm, dtype, n, hw, cs, stream):
        G = _t(gcond, (n, hw, cs), _DT[dtype]).float()[..., :3].permute(0, 2, 1)
        _t(gm, (n, 1, hw), torch.float32).copy_(-(_t(x, (n, 3, hw), torch.float32) * G).sum(1, keepdim=True))

    def e_paste_bwd_mask(self, gout, x, fake, gm, n, hw, stream):
        d = _t(fake, (n, 3, hw), torch.float32) - _t(x, (n, 3, hw), torch.float32)
        _t(gm, (n, 1, hw), torch.float32).copy_((_t(gout, (n, 3, hw), torch.float32) * d).sum(1, keepdim=True))

    # ---- masker losses (NCHW fp32): value added to `loss`, gradient of that value written (include/cgb200.h) ----------------
    def _loss(self, fn, X, loss, gx):
        with torch.enable_grad():
            Xg = X.clone().requires_grad_(True)
            val = fn(Xg)
            g = torch.autograd.grad(val, Xg, allow_unused=True)[0] if gx is not None else None
        _t(loss, (1,), torch.float32).add_(val.detach().float())
        if gx is not None:
            gx.copy_(g if g is not None else torch.zeros_like(X))

    def e_softmax_nchw_fwd(self, x, y, n, c, hw, stream):
        _t(y, (n, c, hw), torch.float32).copy_(torch.softmax(_t(x, (n, c, hw), torch.float32), 1))

    def e_softmax_nchw_bwd(self, y, gy, gx, n, c, hw, stream):
        Y, G = _t(y, (n, c, hw), torch.float32), _t(gy, (n, c, hw), torch.float32)
        _t(gx, (n, c, hw), torch.float32).copy_(Y * (G - (G * Y).sum(1, keepdim=True)))

    def e_cross_entropy_nchw(self, logits, target, loss, glogits, n, c, hw, stream):
        T = _t(target, (n, hw), torch.int64)
        self._loss(lambda X: F.cross_entropy(X, T), _t(logits, (n, c, hw), torch.float32), loss, _t(glogits, (n, c, hw), torch.float32))

    @staticmethod
    def _entropy(P, c):
        import math

        return -P * torch.log2(P + 1e-30) / math.log2(c)

    def e_entropy_nchw(self, p, depth, ge, out, n, c, hw, backward, stream):
        P = _t(p, (n, c, hw), torch.float32)
        D = _t(depth, (n, 1, hw), torch.float32)
        f = (lambda X: self._entropy(X, c) * D) if D is not None else (lambda X: self._entropy(X, c))
        O = _t(out, (n, c, hw), torch.float32)
        if not backward:
            O.copy_(f(P))
        else:
            O.copy_(self._vjp(f, P, _t(ge, (n, c, hw), torch.float32)))

    def e_minent_loss(self, p, loss, gp, acc, n, c, hw, version, lambda_var, stream):
        def f(X):
            e = self._entropy(X, c)
            if version == 1:
                return e.sum() / (n * hw)
            dm = e - e.sum() / (n * hw)
            return (e + lambda_var * dm * dm).sum() / (n * hw)

        self._loss(f, _t(p, (n, c, hw), torch.float32), loss, _t(gp, (n, c, hw), torch.float32))

    def e_sigmoid_pair(self, logits, gprob, out, n, hw, backward, stream):
        L = _t(logits, (n, 1, hw), torch.float32)
        sg = torch.sigmoid(L)
        if not backward:
            _t(out, (n, 2, hw), torch.float32).copy_(torch.cat([sg, 1 - sg], 1))
        else:
            G = _t(gprob, (n, 2, hw), torch.float32)
            _t(out, (n, 1, hw), torch.float32).copy_((G[:, :1] - G[:, 1:]) * sg * (1 - sg))

    def e_tv_loss(self, x, loss, gx, n, c, h, w, stream):
        def f(X):
            h_tv = torch.pow(X[:, :, 1:, :] - X[:, :, : h - 1, :], 2).sum()
            w_tv = torch.pow(X[:, :, :, 1:] - X[:, :, :, : w - 1], 2).sum()
            return 2 * (h_tv / (c * (h - 1) * w) + w_tv / (c * h * (w - 1))) / n

        self._loss(f, _t(x, (n, c, h, w), torch.float32), loss, _t(gx, (n, c, h, w), torch.float32))

    def e_bce_logits_loss(self, x, target, loss, gx, count, stream):
        T = _t(target, (count,), torch.float32)
        self._loss(lambda X: F.binary_cross_entropy_with_logits(X, T), _t(x, (count,), torch.float32), loss, _t(gx, (count,), torch.float32))

    def e_ground_intersection_loss(self, pred, ground, loss, count, stream):
        v = (1.0 * ((_t(ground, (count,), torch.float32) - _t(pred, (count,), torch.float32)) > 0.5)).mean()
        _t(loss, (1,), torch.float32).add_(v)

    def e_dada_depth_loss(self, pred, label, loss, gpred, count, stream):
        L = _t(label, (count,), torch.float32)

        def f(P):   # DADADepthLoss.loss_calc_depth, losses.py:603-617 (the threshold is taken with .item(): a constant)
            adiff = torch.abs(P - L)
            c = 0.2 * float(adiff.max())
            t1 = adiff * (adiff <= c).float()
            t2 = (adiff * adiff + c * c) / (2 * c) * (adiff > c).float()
            return (t1.sum() + t2.sum()) / count

        self._loss(f, _t(pred, (count,), torch.float32), loss, _t(gpred, (count,), torch.float32))

    def e_sigm_loss(self, pred, target, loss, gpred, ws, n, h, w, gmweight, scales, stream):
        T = _t(target, (n, 1, h, w), torch.float32)

        def f(P):   # SIGMLoss.forward, losses.py:250-278 (Sobel kernels expanded to [B,1,3,3] as there)
            t_p, t_t = torch.median(P), torch.median(T)
            s_p, s_t = torch.mean(torch.abs(P - t_p)), torch.mean(torch.abs(T - t_t))
            R = (P - t_p) / s_p - (T - t_t) / s_t
            sx = torch.tensor([[1.0, 0, -1], [2, 0, -2], [1, 0, -1]]).expand(n, 1, 3, 3)
            sy = torch.tensor([[1.0, 2, 1], [0, 0, 0], [-1, -2, -1]]).expand(n, 1, 3, 3)
            gm = 0
            for k in range(scales):
                R_ = F.interpolate(R, scale_factor=1 / 2 ** k)
                gm = gm + torch.sum(torch.abs(F.conv2d(R_, sx)) + torch.abs(F.conv2d(R_, sy)))
            return 0.5 / (h * w) * torch.sum(torch.abs(R)) + gmweight / (h * w) * gm

        self._loss(f, _t(pred, (n, 1, h, w), torch.float32), loss, _t(gpred, (n, 1, h, w), torch.float32))


    # ---- inference events (Trainer.infer_all): restated from the contracts in include/cgb200.h ------------------------------
    @staticmethod
    def _norm(t, mm, lo=0.0, hi=1.0):   # tutils.normalize :566-577 with the per-sample (min, max) the caller computed
        n = t.shape[0]
        mn = mm[:, 0].view(n, *([1] * (t.dim() - 1)))
        mx = mm[:, 1].view(n, *([1] * (t.dim() - 1)))
        return lo + (hi - lo) * (t - mn) / (mx - mn)

    # ---- DiffTransforms (transforms.py:493-626), restated from the header contract ------------------------------------------
    @staticmethod
    def _diffaug_live(P, n, h, w, cut_h, cut_w):
        """[n, h, w] bool over OUTPUT pixels: source inside the image and outside the cutout; plus the source indices."""
        ii = torch.arange(h).view(1, h, 1)
        jj = torch.arange(w).view(1, 1, w)
        tx, ty = P[:, 3].long().view(n, 1, 1), P[:, 4].long().view(n, 1, 1)
        si, sj = ii + tx, jj + ty
        live = (si >= 0) & (si < h) & (sj >= 0) & (sj < w)
        if cut_h > 0:
            a = P[:, 5].long().view(n, 1, 1) - cut_h // 2
            b = P[:, 6].long().view(n, 1, 1) - cut_w // 2
            cut = (ii >= a.clamp(min=0)) & (ii <= (a + cut_h - 1).clamp(max=h - 1)) & (jj >= b.clamp(min=0)) & (jj <= (b + cut_w - 1).clamp(max=w - 1))
            live = live & ~cut
        return live, si.clamp(0, h - 1).expand(n, h, w), sj.clamp(0, w - 1).expand(n, h, w)

    def e_diff_aug_sum(self, t, params, sums, n, c, h, w, cut_h, cut_w, mode, stream):
        T = _t(t, (n, c, h, w), torch.float32).double()
        S = _t(sums, (n,), torch.float64)
        if mode == 1:
            live, _, _ = self._diffaug_live(_t(params, (n, 8), torch.float32), n, h, w, cut_h, cut_w)
            T = T * live.unsqueeze(1)
        S += T.sum((1, 2, 3))

    def e_diff_aug_fwd(self, x, params, sums, y, n, c, h, w, cut_h, cut_w, stream):
        X = _t(x, (n, c, h, w), torch.float32)
        P = _t(params, (n, 8), torch.float32)
        b, cf, sf = (P[:, k].view(n, 1, 1, 1) for k in range(3))
        mean_b = (_t(sums, (n,), torch.float64) / (c * h * w)).float().view(n, 1, 1, 1) + b
        v2 = (X + b - mean_b) * cf + mean_b
        mch = v2.mean(1, keepdim=True)
        v3 = (v2 - mch) * sf + mch
        live, si, sj = self._diffaug_live(P, n, h, w, cut_h, cut_w)
        nn_ = torch.arange(n).view(n, 1, 1).expand(n, h, w)
        moved = v3.permute(0, 2, 3, 1)[nn_, si, sj].permute(0, 3, 1, 2)
        _t(y, (n, c, h, w), torch.float32).copy_(moved * live.unsqueeze(1))

    def e_diff_aug_bwd(self, gy, params, gsums, gx, n, c, h, w, cut_h, cut_w, stream):
        G = _t(gy, (n, c, h, w), torch.float32)
        P = _t(params, (n, 8), torch.float32)
        cf, sf = P[:, 1].view(n, 1, 1, 1), P[:, 2].view(n, 1, 1, 1)
        live, si, sj = self._diffaug_live(P, n, h, w, cut_h, cut_w)
        g3 = torch.zeros(n, h, w, c)
        nn_ = torch.arange(n).view(n, 1, 1).expand(n, h, w)
        src = (G * live.unsqueeze(1)).permute(0, 2, 3, 1)
        g3.index_put_((nn_[live], si[live], sj[live]), src[live])          # a translation is one-to-one on the live pixels
        g3 = g3.permute(0, 3, 1, 2)
        g2 = sf * g3 + (1 - sf) * g3.mean(1, keepdim=True)
        through = (1 - cf) * (_t(gsums, (n,), torch.float64) / (c * h * w)).float().view(n, 1, 1, 1)
        _t(gx, (n, c, h, w), torch.float32).copy_(cf * g2 + through)

    def e_argmax_confusion(self, logits, label, conf, label_max, n, c, hw, stream):   # eval_metrics.py:68-124 (header contract)
        P = torch.argmax(_t(logits, (n, c, hw), torch.float32), 1).reshape(-1)
        L = _t(label, (n * hw,), torch.int64)
        col = torch.where((L >= 0) & (L < c), L, torch.full_like(L, c))
        Cm = _t(conf, (c * (c + 1),), torch.int64)
        Cm += torch.bincount(P * (c + 1) + col, minlength=c * (c + 1))
        M = _t(label_max, (1,), torch.int64)
        M[0] = max(int(M[0]), int(L.max()))

    def e_minmax_per_sample(self, x, mm, n, count, stream):
        X = _t(x, (n, count), torch.float32)
        M = _t(mm, (n, 2), torch.float32)
        M[:, 0], M[:, 1] = X.min(1)[0], X.max(1)[0]

    def e_fire_tone(self, x, mm, out, gray_sum, n, hw, contrast, brightness, stream):
        from torchvision.transforms.functional import adjust_brightness, adjust_contrast

        t = self._norm(_t(x, (n, 3, hw), torch.float32), _t(mm, (n, 2), torch.float32), 0, 255).clone()   # fire.py:80-86
        t[:, 2] -= 20
        t[:, 1] -= 10
        t[:, 0] += 40
        t = t.clamp_(0, 255).to(torch.uint8).view(n, 3, hw, 1)
        t = adjust_brightness(adjust_contrast(t, contrast_factor=contrast), brightness_factor=brightness)            # :90-91
        _t(out, (n, 3, hw), torch.float32).copy_(t.view(n, 3, hw).float())

    def e_sky_mask(self, seg, out, n, c, hs, ws, sky_idx, crop_bottom, stream):
        m = (torch.argmax(_t(seg, (n, c, hs, ws), torch.float32), 1) == sky_idx).float()                              # tutils :579-596
        if crop_bottom:
            m[:, 2 * hs // 3:, :] = 0                                                                                    # fire.py:95-97
        _t(out, (n, hs, ws), torch.float32).copy_(m)

    def e_plane_resize_nearest(self, x, y, n, hi, wi, ho, wo, stream):
        _t(y, (n, ho, wo), torch.float32).copy_(F.interpolate(_t(x, (n, 1, hi, wi), torch.float32), (ho, wo))[:, 0])

    def e_box_dilate(self, x, tmp, y, n, h, w, radius_w, radius_h, stream):
        X = _t(x, (n, 1, h, w), torch.float32)                                    # increase_sky_mask on a binary mask, fire.py:15-47
        Y = F.max_pool2d(F.pad(X, (radius_w, radius_w, radius_h, radius_h)), (2 * radius_h + 1, 2 * radius_w + 1), 1)
        _t(y, (n, h, w), torch.float32).copy_((Y[:, 0] >= 1).float() if False else Y[:, 0].clamp(max=1))

    def e_gauss_blur(self, x, tmp, y, n, h, w, ksize, sigma, stream):
        ax = torch.arange(ksize, dtype=torch.float64) - ksize // 2              # kornia get_gaussian_kernel2d + filter2d, restated
        g = torch.exp(-ax * ax / (2 * sigma * sigma))                            # (separable, in fp64: the 2-D kernel is an outer product)
        g = g / g.sum()
        X = F.pad(_t(x, (n, 1, h, w), torch.float32).double(), (ksize // 2,) * 4, mode="reflect")
        X = F.conv2d(X, g.view(1, 1, 1, ksize))
        X = F.conv2d(X, g.view(1, 1, ksize, 1))
        _t(y, (n, h, w), torch.float32).copy_(X[:, 0])

    def e_fire_paste(self, img, sky, out, n, h, w, fr, fg, fb, transparency, brightness, stream):
        from torchvision.transforms.functional import adjust_brightness

        I = _t(img, (n, 3, h, w), torch.float32)
        m = transparency / 255.0 * _t(sky, (n, 1, h, w), torch.float32)                                               # fire.py:130-133
        filt = torch.tensor([fr, fg, fb]).view(1, 3, 1, 1)
        t = m * filt + (1.0 - m) * I
        t = adjust_brightness(t.to(torch.uint8), brightness).float()                                                   # :120-121
        t[:, :, 0, 0] = 255.0                                                                                           # :124-125
        t[:, :, -1, -1] = 0.0
        _t(out, (n, 3, h, w), torch.float32).copy_(t)

    def e_smog(self, x, mmx, d, mmd, out, n, h, w, hd, wd, airlight, beta, alpha, yr, yg, yb, stream):
        X = self._norm(_t(x, (n, 3, h, w), torch.float32), _t(mmx, (n, 2), torch.float32))                             # srgb2lrgb, tutils :534-538
        irr = torch.where(X <= 0.04045, X / 12.92, ((X + 0.055) / 1.055) ** 2.4)
        D = self._norm(_t(d, (n, 1, hd, wd), torch.float32), _t(mmd, (n, 2), torch.float32), 0.3, 1.0)               # trainer.py:1911-1913
        D = 1.0 / D
        mm2 = torch.stack([D.reshape(n, -1).min(1)[0], D.reshape(n, -1).max(1)[0]], 1)
        D = self._norm(D, mm2, 0.1, 1.0)
        D = F.interpolate(D, size=(h, w), mode="bilinear", align_corners=True)
        tr = torch.exp(-beta * D)
        sm = tr * irr + (1 - tr) * airlight
        sm = torch.where(sm <= 0.0031308, 12.92 * sm, 1.055 * torch.pow(sm, 1 / 2.4) - 0.055)                        # lrgb2srgb :541-563
        yel = torch.tensor([yr, yg, yb]).view(1, 3, 1, 1)
        _t(out, (n, 3, h, w), torch.float32).copy_(sm * (1 - alpha) + yel * alpha)

    def e_perlin_noise(self, angles, out, h, w, res0, res1, stream):
        import math

        A = _t(angles, (res0 + 1, res1 + 1), torch.float32)                                                             # rand_perlin_2d, tutils :648-686
        delta = (res0 / h, res1 / w)
        d = (h // res0, w // res1)
        grid = torch.stack(torch.meshgrid(torch.arange(0, res0, delta[0]), torch.arange(0, res1, delta[1]), indexing="ij"), -1) % 1
        grads = torch.stack((torch.cos(A), torch.sin(A)), -1)

        def tile(s1, s2):
            return grads[s1[0]:s1[1], s2[0]:s2[1]].repeat_interleave(d[0], 0).repeat_interleave(d[1], 1)

        def dot(grad, shift):
            return (torch.stack((grid[:h, :w, 0] + shift[0], grid[:h, :w, 1] + shift[1]), -1) * grad[:h, :w]).sum(-1)

        n00, n10 = dot(tile([0, -1], [0, -1]), [0, 0]), dot(tile([1, None], [0, -1]), [-1, 0])
        n01, n11 = dot(tile([0, -1], [1, None]), [0, -1]), dot(tile([1, None], [1, None]), [-1, -1])
        t = grid[:h, :w]
        t = 6 * t ** 5 - 15 * t ** 4 + 10 * t ** 3
        res = math.sqrt(2) * torch.lerp(torch.lerp(n00, n10, t[..., 0]), torch.lerp(n01, n11, t[..., 0]), t[..., 1])
        _t(out, (1, h, w), torch.float32).copy_(res.view(1, h, w))

    def e_cloudy_mix(self, x, seg, noise, mm_noise, out, n, h, w, c, hs, ws, sky_idx, weight, stream):
        X = _t(x, (n, 3, h, w), torch.float32)
        S = F.interpolate(_t(seg, (n, c, hs, ws), torch.float32), (h, w), mode="bilinear")                            # generator.py:318-323
        mask = (torch.argmax(S, 1, keepdim=True) == sky_idx).float()
        nz = _t(noise, (1, 1, h, w), torch.float32) - _t(mm_noise, (1, 2), torch.float32)[0, 0]                      # mix_noise, tutils :689-694
        _t(out, (n, 3, h, w), torch.float32).copy_(mask * (weight * nz + (1 - weight) * X) + (1 - mask) * X)

    def e_to_uint8_nhwc(self, x, mm, out, n, hw, stream):
        t = self._norm(_t(x, (n, 3, hw), torch.float32), _t(mm, (n, 2), torch.float32))                                # trainer.py:312-327
        buf = (C.c_uint8 * (n * hw * 3)).from_address(_addr(out))
        torch.frombuffer(buf, dtype=torch.uint8).view(n, hw, 3).copy_((t.permute(0, 2, 1) * 255).to(torch.uint8))

    def e_mask_to_uint8(self, m, out, bin_value, count, stream):
        buf = (C.c_uint8 * count).from_address(_addr(out))
        torch.frombuffer(buf, dtype=torch.uint8).copy_(((_t(m, (count,), torch.float32) > bin_value) * 255).to(torch.uint8))


@contextlib.contextmanager
def emulated_library():
    real = _lib.lib()
    fake = EmuLib(real)
    from climategan_b200.trainer import Trainer

    saved = (_lib._lib, ops._on_device, torch.cuda.current_stream, Trainer._pinned)
    _lib._lib = fake
    ops._on_device = lambda x: True
    torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)
    Trainer._pinned = lambda self, key, like: torch.empty(like.shape, dtype=like.dtype)   # (pinned host memory needs a CUDA driver)
    try:
        yield fake
    finally:
        _lib._lib, ops._on_device, torch.cuda.current_stream, Trainer._pinned = saved
        ops.invalidate_weight_cache()
