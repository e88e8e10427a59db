"""cta_group::2 streaming kernel (CGB_TC2=1) against the one-CTA streaming kernel: same operands, bitwise-equal outputs expected
(identical MMA sequence per accumulator), and timing.  Run each mode in its own process:
    CGB_TC2=0 python scripts/exp/tc2_check.py save ; CGB_TC2=1 python scripts/exp/tc2_check.py check"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from climategan_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
mode = sys.argv[1]
CASES = {
    # which, n, ci, co, h, w, k, pad, dil, stride
    "r1": ("fwd", 8, 256, 1024, 80, 80, 1, 0, 1, 1),
    "r1b": ("fwd", 8, 1024, 256, 80, 80, 1, 0, 1, 1),
    "r3": ("fwd", 8, 256, 256, 80, 80, 3, 2, 2, 1),
    "r3d": ("dgrad", 8, 256, 256, 80, 80, 3, 2, 2, 1),
    "odd": ("fwd", 3, 64, 72, 37, 29, 3, 1, 1, 1),          # odd pixel-tile count, ragged edges, N = 72 (not a multiple of 16 -> one-CTA path)
    "odd80": ("fwd", 5, 64, 80, 37, 29, 3, 1, 1, 1),        # odd pixel-tile count on the pair path
    "s2": ("fwd", 4, 64, 128, 64, 64, 1, 0, 1, 2),
    "aspp": ("fwd", 8, 2048, 256, 80, 80, 3, 12, 12, 1),
    "l4": ("fwd", 8, 512, 2048, 80, 80, 1, 0, 1, 1),
    "stats": ("fwd_stats", 8, 256, 256, 80, 80, 3, 2, 2, 1),
    "vgg3": ("fwd", 8, 256, 256, 160, 160, 3, 1, 1, 1),
    "vgg3d": ("dgrad", 8, 256, 256, 160, 160, 3, 1, 1, 1),
    "vgg4": ("fwd", 8, 512, 512, 80, 80, 3, 1, 1, 1),
    "r4": ("fwd", 8, 512, 512, 80, 80, 3, 4, 4, 1),
    "d3": ("fwd", 16, 512, 512, 39, 39, 4, 1, 1, 1),
    "gb160": ("fwd", 8, 128, 160, 160, 160, 3, 1, 1, 1),
    "l4b": ("fwd", 8, 2048, 512, 80, 80, 1, 0, 1, 1),
    "r1d": ("dgrad", 8, 1024, 256, 80, 80, 1, 0, 1, 1),
    "w_r3": ("wgrad", 8, 256, 256, 80, 80, 3, 2, 2, 1),
    "w_r1": ("wgrad", 8, 256, 1024, 80, 80, 1, 0, 1, 1),
    "w_r1b": ("wgrad", 8, 1024, 256, 80, 80, 1, 0, 1, 1),
    "w_r4": ("wgrad", 8, 512, 512, 80, 80, 3, 4, 4, 1),
    "w_l4": ("wgrad", 8, 2048, 512, 80, 80, 1, 0, 1, 1),
    "w_aspp": ("wgrad", 8, 2048, 256, 80, 80, 3, 12, 12, 1),
    "w_s2": ("wgrad", 4, 256, 512, 40, 40, 1, 0, 1, 2),
    "w_128": ("wgrad", 8, 128, 128, 80, 80, 3, 1, 1, 1),
}
out = {}
_only = [t for t in os.environ.get("ONLY", "").split(",") if t]
if _only:
    CASES = {k_: v for k_, v in CASES.items() if k_ in _only}
for name, (which, n, ci, co, h, w, k, pad, dil, stride) in CASES.items():
    torch.manual_seed(len(name) + ci)
    ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
    x = torch.randn(n, h, w, ci, device=dev).bfloat16()
    wp = (torch.randn(co, k * k, ci, device=dev) * 0.05).bfloat16()
    bias = torch.randn(co, device=dev)
    gy = torch.randn(n, ho, wo, co, device=dev).bfloat16()
    g = ops.ConvGeom(k, k, stride, dil, pad, _lib.PAD_ZERO, _lib.ACT_RELU if which == "fwd" else _lib.ACT_NONE, 0.2, _lib.ENGINE_TCGEN05)

    def run():
        if which == "fwd":
            return ops.conv_fwd_raw(x, wp, bias, None, g)
        if which == "fwd_stats":
            y, part = ops.conv_fwd_raw(x, wp, None, None, g, True)
            return torch.cat([y.float().flatten(), part.sum(0).flatten()])
        if which == "wgrad":
            return ops.conv_wgrad_raw(x, gy, g, False)[0]
        return ops.conv_dgrad_raw(gy, wp, (n, h, w, ci), g)

    y = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out[name] = y.float().cpu()
    print(f"{name:6s} {which:9s} {ci}->{co} k{k} d{dil} s{stride} @{h}x{w} n={n}: {ms * 1e3:8.1f} us", flush=True)
path = "/tmp/tc2_ref.pt"
if mode == "save":
    torch.save(out, path)
else:
    ref = torch.load(path)
    bad = 0
    for k_, v in out.items():
        if k_ == "stats" or k_.startswith("w_"):     # (atomically-ordered fp32 reductions: not bit-reproducible run to run)
            d = float((v - ref[k_]).abs().max() / ref[k_].abs().max())
            ok = d < 1e-5
        else:
            d = float((v - ref[k_]).abs().max())
            ok = d == 0.0
        print(("OK  " if ok else "FAIL"), k_, "max |diff| =", d)
        bad += 0 if ok else 1
    sys.exit(1 if bad else 0)
