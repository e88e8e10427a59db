#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the ClimateGAN hot path on B200.

Workloads:
  full     (default) BASELINE.json's metric: the full Masker+Painter G+D train step — Trainer.update_G + Trainer.update_D on
           tasks [d, s, m, p] (DeepLab-v2 ResNet-101 masker in train mode, SPADE painter, OmniDiscriminator, VGG loss,
           every masker loss, ExtraAdam), 8 images per domain (r, s, rf) per GPU, 640x640, bf16 storage / fp32 accumulate.
           "images/sec" = per-domain images per second (the reference's batch_size convention, data.py:512).
  painter  C1 / BASELINE.json configs[1]: painter-only OmniGenerator.paint + L1 + backward, 16 images per GPU.
One "step" = one such pass over one batch of synthetic input.  N > 1: one process per GPU (torchrun), each rank its own
batch slice (weak scaling), G and D gradients all-reduced (mean) over NCCL as flat buckets after each backward.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload full|painter]

Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm's CPU path (the oracle port — the reference
itself is Python and /root/reference does not travel to the GPU box).
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
PAINTER_STEP_GFLOP = 1551.6   # C1: painter fwd + dgrad + wgrad, minus dgrad into the 3-channel conditioning
FULL_STEP_GFLOP = 10113.0     # C3: one (r, s, rf) image triple through update_G + update_D (encoder 4x fwd + 2x bwd, ...)
INFER_GFLOP = 1339.8          # C4: masker forward 816.9 + painter forward 522.9 per image


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=["full", "painter", "infer"],
                    help="full = Masker+Painter G+D train step (BASELINE.json metric; SURVEY.md §8d C3, 8 images/domain/GPU); "
                         "painter = C1 painter-only fwd+bwd (configs[1], 16 images/GPU); "
                         "infer = C4 Trainer.infer_all (masker + painter + flood/wildfire/smog compositing, 16 images/GPU)")
    ap.add_argument("--batch", type=int, default=0, help="images per domain per GPU per step (default: 8 full, 16 painter)")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="CPU baseline sample batch (default: 2 full, 1 painter)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.batch <= 0:
        a.batch = 8 if a.workload == "full" else 16
    if a.workload == "infer":
        a.no_cpu_baseline = True   # the CPU arm of this workload is the reference's own infer_all, which cannot travel
    if a.cpu_sample_batch <= 0:
        a.cpu_sample_batch = 2 if a.workload == "full" else 1
    return a


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def load_traffic():
    """dram bytes per launch of the dominant kernels from the committed `ncu --set full` captures (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


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


def metric_name(args):
    return {"full": "full_train_step_images_per_sec", "painter": "painter_fwd_bwd_images_per_sec",
            "infer": "infer_all_images_per_sec"}[args.workload]


def workload_config(args):
    if args.workload == "full":
        return {
            "workload": f"C3 full Masker+Painter G+D train step (Trainer.update_G + update_D, tasks d,s,m,p; deeplabv2 ResNet-101 "
                        f"encoder + DADA depth + DeepLab-v2 seg + base mask decoders, SPADE painter, OmniDiscriminator, VGG loss), "
                        f"{args.batch} images per domain (r,s,rf) per GPU, {args.size}x{args.size}",
            "batch_per_domain_per_gpu": args.batch, "domains": ["r", "s", "rf"], "size": args.size,
            "images_per_sec_convention": "per-domain images/s (reference batch_size convention); x3 for domain-images/s",
            "parallelism": f"dp{args.gpus} (per-image batch split; NCCL all-reduce of the flat G and D gradient buckets)",
            "l2_policy": "working set per step (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
        }
    if args.workload == "infer":
        return {
            "workload": f"C4 Trainer.infer_all (deeplabv2 masker + SPADE painter inference, flood + wildfire + smog compositing, "
                        f"uint8 NHWC outputs), batch {args.batch}/GPU, {args.size}x{args.size}",
            "batch_per_gpu": args.batch, "size": args.size, "parallelism": f"dp{args.gpus} (independent replicas, no collective)",
            "l2_policy": "activations per batch exceed the 126 MB L2; no flush needed",
        }
    return {
        "workload": f"C1 painter-only SPADE generator fwd+bwd (OmniGenerator.paint + L1 + backward), "
                    f"batch {args.batch}/GPU, {args.size}x{args.size}",
        "batch_per_gpu": args.batch, "size": args.size,
        "parallelism": f"dp{args.gpus} (per-image batch split, NCCL all-reduce of the painter grad bucket)",
        "l2_policy": "working set per step (tens of GB of activations) exceeds the 126 MB L2; no flush needed",
    }


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_painter_rate(batch: int, size: int, steps: int, warmup: int):
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
    sample = f"oracle (PyTorch fp32 restatement of the reference) paint+L1+backward, batch {batch} at {size}x{size}, {steps} timed steps"
    return batch / mean_t, mean_t, cores, sample


def cpu_full_rate(batch: int, size: int, steps: int, warmup: int):
    """images/s (per domain) of the reference algorithm's full G+D step on the host cores: oracle/full_step_oracle.py
    (get_G_loss + backward, get_D_loss + backward; the optimiser's elementwise update is not timed — < 1 % on CPU)."""
    from climategan_b200.discriminator import OmniDiscriminator
    from climategan_b200.generator import OmniGenerator
    from climategan_b200.losses import Vgg19
    from climategan_b200.utils import full_opts, synth_batch
    from oracle import full_step_oracle as fo  # CPU baseline leg only

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    opts = full_opts(nblocks=(3, 4, 23, 3), size=size, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3)
    G, D, V = OmniGenerator(opts, latent_shape=(size, size)), OmniDiscriminator(opts), Vgg19()  # parameter containers only
    gsd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    dsd = {k: v.detach().clone() for k, v in D.state_dict().items()}
    vsd = {k: v.detach().clone() for k, v in V.state_dict().items()}
    for k, v in gsd.items():
        if v.dtype.is_floating_point and not k.endswith(("_u", "_v", "running_mean", "running_var")):
            v.requires_grad_(True)
    for k, v in dsd.items():
        if v.dtype.is_floating_point and not k.endswith(("_u", "_v")):
            v.requires_grad_(True)
    mdb = synth_batch(opts, batch, size, 1)
    z = size // 2 ** 7
    times = []
    for i in range(warmup + steps):
        for v in list(gsd.values()) + list(dsd.values()):
            v.grad = None
        t0 = time.perf_counter()
        loss, _ = fo.full_g_loss(gsd, dsd, vsd, mdb, z)
        loss.backward()
        ld, _ = fo.full_d_loss(gsd, dsd, mdb, z)
        ld.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean_t = sum(times) / len(times)
    sample = (f"oracle (PyTorch fp32 restatement of the reference's Trainer.update_G/update_D losses + backward), {batch} images "
              f"per domain (r,s,rf) at {size}x{size} (2 is the minimum: train-mode BatchNorm), {steps} timed step(s), {warmup} warm-up")
    return batch / mean_t, mean_t, cores, sample


def cpu_rate(args, steps, warmup):
    if args.workload == "full":
        return cpu_full_rate(args.cpu_sample_batch, args.size, steps, warmup)
    return cpu_painter_rate(args.cpu_sample_batch, args.size, steps, warmup)


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port), all host threads, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "infer":
        print(json.dumps({"impl": "reference", "unavailable": "the infer workload's CPU arm is the reference's own Trainer.infer_all, "
                          "which needs /root/reference (absent on the GPU box); its parity is pinned by tests/golden/infer_all.*"}))
        return
    steps = max(1, min(args.steps, 1 if args.workload == "full" else 3))
    warmup = 0 if args.workload == "full" else max(1, min(args.warmup, 1))
    rate, t, cores, sample = cpu_rate(args, steps, warmup)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rate, "unit": "img/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": rate, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def build_painter(args, dev, rank, world, dtype):
    from climategan_b200 import ops
    from climategan_b200.generator import OmniGenerator
    from climategan_b200.parallel import GradBucket
    from climategan_b200.utils import default_painter_opts

    B, S = args.batch, args.size
    torch.manual_seed(0)  # identical initial weights on every rank
    G = OmniGenerator(default_painter_opts(), latent_shape=S, storage_dtype=dtype).to(dev).train()
    params = [p for p in G.painter.parameters() if p.requires_grad]
    gen = torch.Generator().manual_seed(1234 + rank)  # each rank its own slice of the global batch
    host = [(torch.rand(B, 3, S, S, generator=gen) * 2 - 1).pin_memory(),
            (torch.rand(B, 1, S, S, generator=gen) > 0.5).float().pin_memory(),
            (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).pin_memory()]
    bucket = GradBucket(params) if world > 1 else None

    def step(x, m, t):
        for p in params:
            p.grad = None
        out = G.paint(m, x)
        loss = ops.l1_loss(out, t)
        loss.backward()
        if bucket is not None:
            bucket.allreduce()  # ONE all-reduce (mean) of the flat painter gradient bucket per step
        return loss

    to_dev = lambda nb: [h.to(dev, non_blocking=nb) for h in host]  # noqa: E731
    h2d = int(sum(h.numel() * h.element_size() for h in host))
    return step, to_dev, h2d, PAINTER_STEP_GFLOP


def build_full(args, dev, rank, world, dtype):
    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts, synth_batch

    B, S = args.batch, args.size
    torch.manual_seed(0)  # identical initial weights on every rank
    opts = full_opts(nblocks=(3, 4, 23, 3), size=S, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3)
    opts.dis.soft_shift, opts.dis.flip_prob = 0.2, 0.05   # defaults.yaml:194-195 (label smoothing / flipping on, as in training)
    t = Trainer(opts, device=dev, storage_dtype=dtype).setup(input_shape=(S, S))
    mdb = synth_batch(opts, B, S, seed=1234 + rank)       # each rank its own slice of the global batch
    host = {dom: {k: v.pin_memory() for k, v in b["data"].items()} for dom, b in mdb.items()}
    if world > 1:
        t.enable_data_parallel()

    def step(batch):
        t.update_G(batch)
        t.update_D(batch)
        t.logger.global_step += 1
        return t.logger.losses.gen.total_loss

    def to_dev(nb):
        return [{dom: {"data": {k: v.to(dev, non_blocking=nb) for k, v in d.items()}, "domain": [dom] * B, "mode": ["train"] * B,
                       "paths": {}} for dom, d in host.items()}]

    h2d = int(sum(v.numel() * v.element_size() for d in host.values() for v in d.values()))
    return step, to_dev, h2d, FULL_STEP_GFLOP


def build_infer(args, dev, rank, world, dtype):
    import random

    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts

    B, S = args.batch, args.size
    torch.manual_seed(0)
    opts = full_opts(nblocks=(3, 4, 23, 3), size=S, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3)
    t = Trainer(opts, device=dev, storage_dtype=dtype).setup(inference=True, input_shape=(S, S))
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [(torch.rand(B, 3, S, S, generator=gen) * 2 - 1).pin_memory()]
    random.seed(0)
    state = {"numpy": False}

    def step(x):
        out = t.infer_all(x, numpy=state["numpy"])
        return out["flood"]

    def to_dev(nb):
        state["numpy"] = nb          # the e2e leg (non_blocking H2D from pinned memory) also returns uint8 NHWC arrays on the host
        return [host[0].to(dev, non_blocking=nb)]

    return step, to_dev, int(host[0].numel() * 4), INFER_GFLOP


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    from climategan_b200 import _lib

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

        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    B = args.batch

    builder = {"full": build_full, "painter": build_painter, "infer": build_infer}[args.workload]
    step, to_dev, h2d_bytes, gflop_per_image = builder(args, dev, rank, world, dtype)
    resident = to_dev(False)

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

    # ---- warm-up (the last warm-up step runs with the per-launch event profiler on, so its event pool exists before timing)
    import ctypes as C

    for i in range(args.warmup):
        if i == args.warmup - 1:
            lib.cgb_prof_enable(1)
        step(*resident)
    barrier()
    lib.cgb_prof_enable(0)
    _scratch = C.create_string_buffer(1 << 22)
    lib.cgb_prof_dump(_scratch, len(_scratch))   # discard the warm-up records (returns their events to the pool)

    # ---- device-resident timed region (value), with clocks + per-launch conv timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.cgb_launch_count_reset()
    lib.cgb_prof_enable(1)
    total_ms = timed(lambda: step(*resident), args.steps)
    lib.cgb_prof_enable(0)
    launches = int(lib.cgb_launch_count())
    clocks = sampler.stop() if rank == 0 else None
    buf = C.create_string_buffer(1 << 22)
    lib.cgb_prof_dump(buf, len(buf))
    prof = [ln.split() for ln in buf.value.decode().strip().splitlines() if ln.strip()]

    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end-to-end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        def e2e_step():
            res = step(*to_dev(True))    # H2D of this step's inputs from pinned host memory
            if isinstance(res, torch.Tensor):
                return float(res.sum().item()) if res.numel() > 1 else float(res.item())   # D2H read of the step result
            return int(res[0, 0, 0, 0])  # infer: the uint8 NHWC events were already copied to the host by infer_all

        e2e_step()
        e2e_ms = timed(e2e_step, args.steps) / args.steps
        d2h = 4 if args.workload != "infer" else 3 * B * args.size * args.size * 3   # three uint8 NHWC events per image
        e2e = {"value": world * B / (e2e_ms / 1e3), "unit": "img/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (largest share of conv time in the timed region)
    peaks, peak_src = load_peaks()
    traffic = load_traffic()
    names = {0: "fwd", 1: "dgrad", 2: "wgrad"}
    rows = []
    for f in prof:
        which, tc, n, hi, wi, ci, ho, wo, co, kh, kw, stride, dil, count = map(int, f[:14])
        tot = float(f[14])
        rows.append(dict(op=names[which], engine="tcgen05" if tc else "simt", n=n, hi=hi, wi=wi, ho=ho, wo=wo, ci=ci, co=co, k=kh,
                         stride=stride, dil=dil, count=count, total_ms=tot))
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
        name = (f"conv {r['op']} [{r['engine']}] {r['ci']}->{r['co']} {r['k']}x{r['k']} s{r['stride']} d{r['dil']} "
                f"@{r['hi']}x{r['wi']} n={r['n']}")
        tkey = f"{r['op']} {r['ci']}->{r['co']} k{r['k']} s{r['stride']} d{r['dil']} @{r['hi']}x{r['wi']} n={r['n']}"
        common = {"kernel": name, "avg_launch_ms": avg_ms, "share_of_conv_time": r["total_ms"] / conv_ms,
                  "conv_time_share_of_step": conv_ms / total_ms, "algorithmic_flops_per_launch": flops,
                  "algorithmic_bytes_per_launch": byts, "traffic": traffic.get(tkey)}
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
    step_tflops = world * B * gflop_per_image * 1e9 / (ms_per_step * 1e-3) / 1e12

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N=1 only
        rate, t, cores, sample = cpu_rate(args, 1 if args.workload == "full" else 2, 0 if args.workload == "full" else 1)
        cpu = {"value": rate, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample + f" ({t:.1f} s/step)"}

    line = {
        "metric": metric_name(args), "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": roofline, "roofline_top_tensor_kernel": roofline_tensor, "cpu_baseline": cpu,
        "step_tflops_algorithmic": step_tflops,
        "step_frac_of_bf16_peak": step_tflops / (world * tpeak),
        "conv_time_share_of_step": conv_ms / total_ms,
        "top_kernels": [
            {"kernel": f"{r['op']}[{r['engine']}] {r['ci']}->{r['co']} k{r['k']} s{r['stride']} d{r['dil']} @{r['hi']}x{r['wi']}",
             "count": r["count"], "total_ms": round(r["total_ms"], 3)} for r in rows[:int(os.environ.get("CGB_TOPK", "8"))]],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _logical_flops(r):
    """Algorithmic FLOPs of one launch: 2*N*Ho*Wo*Cout*Cin*k*k with LOGICAL channel counts.
    Storage channels are logical channels rounded up to 8 (20->24; the fused gamma||beta conv has
    2*round8(C)); the logical counts on the path are 3 (images), 4 (mask+image), 11 (classes), 20, 40, ... and 128."""
    def logical(cs):
        table = {8: 3, 24: 20, 48: 40}  # 8: a 3-channel image; 24: 20 ch; 48: gamma||beta of 20 ch
        return table.get(cs, cs)
    return 2.0 * r["n"] * r["ho"] * r["wo"] * logical(r["co"]) * logical(r["ci"]) * r["k"] * r["k"]


if __name__ == "__main__":
    main()
