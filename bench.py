#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the ClimateGAN hot path on B200.

Workload (BASELINE.json configs[1], SURVEY.md §8d "C1"): Painter-only SPADE generator forward +
backward — OmniGenerator.paint(m, x) -> L1 to a target -> backward — batch 16 per GPU, 640x640,
bf16 storage / fp32 accumulate, synthetic data, random-init weights.  One "step" = one such pass.
N > 1: one process per GPU (torchrun), each rank its own batch slice (weak scaling), painter
gradients all-reduced (mean) over NCCL as one flat bucket after backward.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm's CPU path (the
oracle port — the reference itself is Python and /root/reference does not travel to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# Algorithmic conv FLOPs per image at 640x640 ([probe] hooks on F.conv2d in the reference):
PAINTER_FWD_GFLOP = 522.86
PAINTER_STEP_GFLOP = 1551.6  # fwd + dgrad + wgrad, minus dgrad into the 3-channel conditioning


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample-batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_rate(batch: int, size: int, steps: int, warmup: int):
    """img/s of the reference algorithm (oracle port, PyTorch fp32) on the host cores: paint + L1 + backward."""
    from climategan_b200.painter import PainterSpadeDecoder
    from climategan_b200.utils import default_painter_opts
    from oracle import painter_oracle as po  # CPU baseline leg only

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    holder = PainterSpadeDecoder(default_painter_opts())  # parameter container only (random init)
    sd = {k: v.detach().clone() for k, v in holder.state_dict().items()}
    for k, v in sd.items():
        v.requires_grad_(not k.endswith(("_u", "_v")))
    x = torch.rand(batch, 3, size, size) * 2 - 1
    m = (torch.rand(batch, 1, size, size) > 0.5).float()
    t = torch.rand(batch, 3, size, size) * 2 - 1
    z = size // 2 ** 7
    times = []
    for i in range(warmup + steps):
        for v in sd.values():
            v.grad = None
        t0 = time.perf_counter()
        out = po.paint(sd, m, x, z, z, po.n_up_spades_of(sd))
        loss = torch.nn.functional.l1_loss(out, t)
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean_t = sum(times) / len(times)
    return batch / mean_t, mean_t, cores


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b = args.cpu_sample_batch
    steps = max(1, min(args.steps, 3))
    warmup = max(1, min(args.warmup, 1))
    rate, t, cores = cpu_reference_rate(b, args.size, steps, warmup)
    sample = f"painter paint+L1+backward, batch {b} of the batch-{args.batch} workload, {args.size}x{args.size}, fp32, {steps} timed steps"
    line = {
        "impl": "reference", "metric": "painter_fwd_bwd_images_per_sec", "value": rate, "unit": "img/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": rate, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {
        "workload": f"C1 painter-only SPADE generator fwd+bwd (OmniGenerator.paint + L1 + backward), "
                    f"batch {args.batch}/GPU, {args.size}x{args.size}",
        "batch_per_gpu": args.batch, "size": args.size,
        "parallelism": f"dp{args.gpus} (per-image batch split, NCCL all-reduce of the painter grad bucket)",
        "l2_policy": "working set per step (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    from climategan_b200 import _lib, ops
    from climategan_b200.generator import OmniGenerator
    from climategan_b200.utils import default_painter_opts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    lib = _lib.lib()
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B, S = args.batch, args.size

    torch.manual_seed(0)  # identical initial weights on every rank
    G = OmniGenerator(default_painter_opts(), latent_shape=S, storage_dtype=dtype).to(dev).train()
    params = [p for p in G.painter.parameters() if p.requires_grad]

    gen = torch.Generator().manual_seed(1234 + rank)  # each rank its own slice of the global batch
    hx = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).pin_memory()
    hm = (torch.rand(B, 1, S, S, generator=gen) > 0.5).float().pin_memory()
    ht = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).pin_memory()
    dx, dm, dt_ = hx.to(dev), hm.to(dev), ht.to(dev)

    from climategan_b200.parallel import GradBucket

    bucket = GradBucket(params) if world > 1 else None

    def allreduce_grads():
        if bucket is not None:
            bucket.allreduce()  # ONE all-reduce (mean) of the flat painter gradient bucket per step

    def step(x, m, t):
        for p in params:
            p.grad = None
        out = G.paint(m, x)
        loss = ops.l1_loss(out, t)
        loss.backward()
        allreduce_grads()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up
    for _ in range(args.warmup):
        step(dx, dm, dt_)
    barrier()

    # ---- device-resident timed region (value), with clocks + per-launch conv timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.cgb_launch_count_reset()
    lib.cgb_prof_enable(1)
    total_ms = timed(lambda: step(dx, dm, dt_), args.steps)
    lib.cgb_prof_enable(0)
    launches = int(lib.cgb_launch_count())
    clocks = sampler.stop() if rank == 0 else None
    import ctypes as C

    buf = C.create_string_buffer(1 << 20)
    lib.cgb_prof_dump(buf, len(buf))
    prof = [ln.split() for ln in buf.value.decode().strip().splitlines() if ln.strip()]

    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end-to-end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            x = hx.to(dev, non_blocking=True)
            m = hm.to(dev, non_blocking=True)
            t = ht.to(dev, non_blocking=True)
            loss = step(x, m, t)
            return float(loss.item())  # D2H read of the step result

        e2e_step()
        e2e_ms = timed(e2e_step, args.steps) / args.steps
        e2e = {"value": world * B / (e2e_ms / 1e3), "unit": "img/s",
               "h2d_bytes_per_step": int((hx.numel() + hm.numel() + ht.numel()) * 4), "d2h_bytes_per_step": 4,
               "ms_per_step": e2e_ms}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of conv time in the timed region)
    peaks, peak_src = load_peaks()
    names = {0: "fwd", 1: "dgrad", 2: "wgrad"}
    rows = []
    for f in prof:
        which, tc, n, hi, wi, ci, ho, wo, co, kh, kw, stride, dil, count = map(int, f[:14])
        tot = float(f[14])
        # algorithmic flops with STORAGE channels (the padded channels are real MMA work but not algorithmic
        # work; report logical = storage here only when they coincide, else scale by the logical fraction below)
        flops = 2.0 * n * ho * wo * co * ci * kh * kw
        rows.append(dict(op=names[which], engine="tcgen05" if tc else "simt", n=n, hi=hi, wi=wi, ho=ho, wo=wo, ci=ci, co=co, k=kh,
                         stride=stride, dil=dil, count=count, total_ms=tot, flops_storage=flops))
    conv_ms = sum(r["total_ms"] for r in rows) or 1e-9
    rows.sort(key=lambda r: -r["total_ms"])
    roofline = None
    roofline_tensor = None
    tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    hpeak = peaks["hbm_gbs"]

    def describe(r):
        """Roofline entry of one (op, shape) class.  The bound is decided by arithmetic intensity: algorithmic FLOPs
        over the bytes the launch must move (operands read once + result written once, storage dtype) against the
        machine ridge (measured bf16 peak / measured HBM bandwidth)."""
        avg_ms = r["total_ms"] / r["count"]
        flops = _logical_flops(r)
        esz = 2 if args.dtype == "bf16" else 4
        px_in, px_out = r["n"] * r["hi"] * r["wi"], r["n"] * r["ho"] * r["wo"]
        wbytes = r["ci"] * r["co"] * r["k"] * r["k"] * esz
        if r["op"] == "wgrad":
            byts = esz * (px_in * r["ci"] + px_out * r["co"]) + 4 * r["ci"] * r["co"] * r["k"] * r["k"]
        else:
            byts = esz * (px_in * r["ci"] + px_out * r["co"]) + wbytes
        ridge = tpeak * 1e12 / (hpeak * 1e9)
        name = f"conv {r['op']} [{r['engine']}] {r['ci']}->{r['co']} {r['k']}x{r['k']} @{r['hi']}x{r['wi']} n={r['n']}"
        common = {"kernel": name, "avg_launch_ms": avg_ms, "share_of_conv_time": r["total_ms"] / conv_ms,
                  "conv_time_share_of_step": conv_ms / total_ms, "algorithmic_flops_per_launch": flops,
                  "algorithmic_bytes_per_launch": byts, "traffic": None}
        if flops / byts >= ridge:
            a = flops / (avg_ms * 1e-3) / 1e12
            return dict(bound="tensor", achieved=a, peak=tpeak, unit="TFLOP/s", frac=a / tpeak,
                        peak_source=f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)", **common)
        a = byts / (avg_ms * 1e-3) / 1e9
        return dict(bound="hbm", achieved=a, peak=hpeak, unit="GB/s", frac=a / hpeak,
                    peak_source=f"{peak_src} hbm_gbs", **common)

    if rows:
        roofline = describe(rows[0])  # the (op, shape) class with the largest share of the timed region
        for r in rows:                # and the largest tensor-bound class, for the tensor-pipe figure
            d = describe(r)
            if d["bound"] == "tensor":
                roofline_tensor = d
                break
    step_tflops = world * B * PAINTER_STEP_GFLOP * 1e9 / (ms_per_step * 1e-3) / 1e12
    peak = tpeak

    cpu = None
    if not args.no_cpu_baseline:
        rate, t, cores = cpu_reference_rate(args.cpu_sample_batch, S, 2, 1)
        cpu = {"value": rate, "unit": "img/s", "cores": cores, "kind": "port",
               "sample": f"oracle (PyTorch fp32 restatement of the reference) paint+L1+backward, batch "
                         f"{args.cpu_sample_batch} at {S}x{S}, 2 timed steps ({t:.1f} s/step)"}

    line = {
        "metric": "painter_fwd_bwd_images_per_sec", "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "roofline_top_tensor_kernel": roofline_tensor, "cpu_baseline": cpu,
        "step_tflops_algorithmic": step_tflops,
        "step_frac_of_bf16_peak": step_tflops / (world * peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])),
        "top_kernels": [
            {"kernel": f"{r['op']}[{r['engine']}] {r['ci']}->{r['co']} k{r['k']} @{r['hi']}x{r['wi']}",
             "count": r["count"], "total_ms": round(r["total_ms"], 3)} for r in rows[:int(os.environ.get("CGB_TOPK", "8"))]],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _logical_flops(r):
    """Algorithmic FLOPs of one launch: 2*N*Ho*Wo*Cout*Cin*k*k with LOGICAL channel counts.
    Storage channels are logical channels rounded up to 8 (20->24; the fused gamma||beta conv has
    2*round8(C)); the painter's logical counts are 3,20,40,80,...,640 and 128."""
    def logical(cs):
        table = {8: 3, 24: 20, 48: 40}  # 8: the 3-channel conditioning; 24: 20 ch; 48: gamma||beta of 20 ch
        return table.get(cs, cs)
    return 2.0 * r["n"] * r["ho"] * r["wo"] * logical(r["co"]) * logical(r["ci"]) * r["k"] * r["k"]


if __name__ == "__main__":
    main()
