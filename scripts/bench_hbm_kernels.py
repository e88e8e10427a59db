"""Achieved HBM bandwidth of the HBM-bound kernels of the path (north_star: "upsample / mask-paste / fire / smog compositing paths
are coalesced vectorised HBM kernels ... evidenced by achieved HBM GB/s"), one launch class per line, at the shapes the full train
step / infer_all run them on (8 images, 640x640): CUDA events on the launching stream, 20 launches after 3 warm-ups, the L2
flushed (a 256 MB write) before every timed launch unless --warm.  Bytes = the algorithmic bytes of the launch (every operand
read once, every result written once).  Usage:  python scripts/bench_hbm_kernels.py [--warm] [--only substr] [--json out]
Under ncu:  ncu --set full -k regex:<kernel> -c 3 python scripts/bench_hbm_kernels.py --only <name> --reps 1"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from climategan_b200 import _lib, events, ops  # noqa: E402
from climategan_b200.utils import Dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--warm", action="store_true", help="no L2 flush between launches (operands may be L2-resident, as right after the producer)")
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--json", default="")
args = ap.parse_args()

dev = torch.device("cuda:0")
_lib.require_device()
peak = 6554.9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bf = torch.bfloat16


def st(n, h, w, c, dt=bf):
    return torch.randn(n, h, w, c, device=dev).to(dt)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(args.reps):
        if not args.warm:
            flush_buf.zero_()
        torch.cuda._sleep(400_000)   # ~0.2 ms of GPU spin: the host enqueues e0 / the launch / e1 while it runs (no host gap in the timing)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / args.reps


CASES = []


def case(name, nbytes, fn):
    if args.only and args.only not in name:
        return
    CASES.append((name, nbytes, fn))


def bn_cases(tag, n, h, w, c):
    x, r, gy = st(n, h, w, c), st(n, h, w, c), st(n, h, w, c)
    el = x.numel()
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    mean, rstd = ops.instnorm_stats(x.view(1, n * h, w, c))
    y = torch.empty_like(x)
    L = _lib.lib()
    p, s_ = ops._p, ops._st

    def fwd(res):
        return lambda: L.cgb_bn_apply_fwd(p(x), p(mean), p(rstd), p(bn.weight.detach()), p(bn.bias.detach()), p(res), p(y), 1, n * h * w, c, 1, 0.0, s_())

    case(f"bn_apply_fwd {tag}", el * 2 * 2, fwd(None))
    case(f"bn_apply_fwd+residual {tag}", el * 2 * 3, fwd(r))
    nd = int(L.cgb_bn_bwd_ws_doubles(n * h * w, c))
    sums = torch.empty(nd, dtype=torch.float64, device=dev)
    gpre, gx = torch.empty_like(x), torch.empty_like(x)
    case(f"bn_apply_bwd {tag}", el * 2 * 4,
         lambda: L.cgb_bn_apply_bwd(p(x), p(mean), p(rstd), p(y), p(gy), p(gpre), p(sums), 1, n * h * w, c, 1, 0.0, s_()))
    case(f"bn_bwd_finalize {tag}", el * 2 * 3,
         lambda: L.cgb_bn_bwd_finalize(p(x), p(mean), p(rstd), p(bn.weight.detach()), p(sums), p(gpre), p(gx), 1, n * h * w, c, s_()))
    case(f"in_stats(batch) {tag}", el * 2, lambda: ops.instnorm_stats(x.view(1, n * h, w, c)))


bn_cases("8x80x80x256", 8, 80, 80, 256)
bn_cases("8x80x80x1024", 8, 80, 80, 1024)
bn_cases("8x160x160x256", 8, 160, 160, 256)

# instance-norm / SPADE elementwise at the painter's last level (C = 20 -> 24 storage channels) and one level up
for tag, (n, h, w, c) in {"8x640x640x24": (8, 640, 640, 24), "8x320x320x40": (8, 320, 320, 40)}.items():
    x, gb, gout = st(n, h, w, c), st(n, h, w, 2 * c), st(n, h, w, c)
    el = x.numel()
    case(f"in_stats {tag}", el * 2, lambda x=x: ops.instnorm_stats(x))
    mean, rstd = ops.instnorm_stats(x)
    out, ggb, gxh = torch.empty_like(x), torch.empty_like(gb), torch.empty_like(x)
    sums = torch.zeros(n, c, 2, dtype=torch.float64, device=dev)
    L, p, s_ = _lib.lib(), ops._p, ops._st
    case(f"spade_mod_fwd {tag}", el * 2 * 4,
         lambda x=x, mean=mean, rstd=rstd, gb=gb, out=out, n=n, h=h, w=w, c=c: L.cgb_spade_modulate_fwd(p(x), p(mean), p(rstd), p(gb), p(out), 1, n, h * w, c, 2, 0.2, s_()))
    case(f"spade_mod_bwd {tag}", el * 2 * 7,
         lambda x=x, mean=mean, rstd=rstd, gb=gb, gout=gout, ggb=ggb, gxh=gxh, sums=sums, n=n, h=h, w=w, c=c:
         L.cgb_spade_modulate_bwd(p(x), p(mean), p(rstd), p(gb), p(gout), p(ggb), p(gxh), p(sums), 1, n, h * w, c, 2, 0.2, s_()))
    case(f"in_bwd {tag}", el * 2 * 2,
         lambda x=x, mean=mean, rstd=rstd, gxh=gxh, sums=sums, n=n, h=h, w=w, c=c: L.cgb_instnorm_bwd(p(x), p(mean), p(rstd), p(sums), p(gxh), 1, n, h * w, c, s_()))
    case(f"upsample2x {tag}", el * 2 * 5, lambda x=x: ops.upsample2x(x))

# layout edges and compositing (NCHW fp32 images, 8 x 3 x 640 x 640)
N, S = 8, 640
img = torch.rand(N, 3, S, S, device=dev) * 2 - 1
msk = (torch.rand(N, 1, S, S, device=dev) > 0.5).float()
fake = torch.rand(N, 3, S, S, device=dev) * 2 - 1
dpt = torch.rand(N, 1, S // 4, S // 4, device=dev)
case("nchw_to_nhwc 8x3x640x640 -> bf16 [..,8]", img.numel() * 4 + N * S * S * 8 * 2, lambda: ops.to_storage(img, bf))
x8 = ops.to_storage(img, bf)
case("nhwc_to_nchw bf16 [..,8] -> 8x3x640x640", img.numel() * 4 + N * S * S * 8 * 2, lambda: ops.from_storage(x8, 3))
case("mask_cond x(1-m) -> storage", img.numel() * 4 + msk.numel() * 4 + N * S * S * 8 * 2, lambda: ops.mask_cond(img, msk, bf))
case("paste x(1-m)+fake*m", (3 * img.numel() + msk.numel()) * 4, lambda: ops.paste(img, msk, fake))
smog_opts = Dict(airlight=0.76, beta=2, vr=1, yellow_color=[224, 192, 29], alpha=20)
case("smog (minmax + fused transmission/sRGB/yellow)", (2 * img.numel()) * 4 + img.numel() * 4 + dpt.numel() * 8, lambda: events.add_smog(img, dpt, smog_opts))
case("to_uint8_nhwc (minmax + normalise)", img.numel() * 4 * 2 + img.numel(), lambda: events.to_uint8_nhwc(img))
plane = torch.rand(N, S, S, device=dev)
tmp, outp = torch.empty_like(plane), torch.empty_like(plane)
case("gauss_blur 281 taps separable (2 passes)", plane.numel() * 4 * 4,
     lambda: _lib.lib().cgb_gauss_blur(ops._p(plane), ops._p(tmp), ops._p(outp), N, S, S, 281, 140.5, ops._st()))

# optimiser: the generator's flat buffer (105.7 M parameters)
npar = 105_700_000 // 64 * 64
bufs = [torch.zeros(npar, device=dev) for _ in range(5)]
bufs[1].normal_()
case("extra_adam 105.7M params", npar * 4 * 7,
     lambda: _lib.lib().cgb_extra_adam(*[ops._p(b) for b in bufs], npar, 1e-4, 0.5, 0.999, 1e-8, 0.0, 3, 0, 1, ops._st()))
gy = st(8, 80, 80, 1024)
gbias = torch.zeros(1024, device=dev)

rows = []
print(f"{'kernel':58s} {'ms':>8s} {'GB/s':>8s} {'of %d' % peak:>8s}   ({'L2 warm' if args.warm else 'L2 flushed'})")
for name, nbytes, fn in CASES:
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    rows.append({"kernel": name, "ms": ms, "bytes": nbytes, "gbs": gbs, "frac": gbs / peak})
    print(f"{name:58s} {ms:8.4f} {gbs:8.0f} {gbs / peak:8.2f}")
if args.json:
    json.dump({"peak_gbs": peak, "l2": "warm" if args.warm else "flushed", "rows": rows}, open(args.json, "w"), indent=1)
