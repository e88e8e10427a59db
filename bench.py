#!/usr/bin/env python
"""bench.py — the driver's measurement contract for the ClimateGAN hot path on B200.

Workloads:
  full     (default) BASELINE.json's metric: the full Masker+Painter G+D train step — Trainer.update_G + Trainer.update_D on
           tasks [d, s, m, p] (DeepLab-v2 ResNet-101 masker in train mode, SPADE painter, OmniDiscriminator, VGG loss,
           every masker loss, ExtraAdam), 8 images per domain (r, s, rf) per GPU, 640x640, bf16 storage / fp32 accumulate.
           "images/sec" = per-domain images per second (the reference's batch_size convention, data.py:512).
  painter  C1 / BASELINE.json configs[1]: painter-only OmniGenerator.paint + L1 + backward, 16 images per GPU.
  masker   C2 / configs[2]: Masker (tasks d,s,m) + AdvEnt discriminators update_G + update_D, 32 images per domain (r, s).
  infer    C4 / configs[4]: Trainer.infer_all (masker + painter + flood / wildfire / smog), 16 images per GPU.
One "step" = one such pass over one batch of synthetic input.  The train workloads replay forward + backward from a CUDA graph
(Trainer.enable_cuda_graphs; --no-graphs times the eager step); the per-class kernel table comes from ONE extra eager step
bracketed launch by launch with CUDA events, outside the timed region.  N > 1: one process per GPU (torchrun), each rank its own
batch slice (weak scaling), G and D gradients all-reduced (mean) over NCCL as flat buckets after each backward.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|gpu-eager] [--workload full|painter|masker|infer]

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
MASKER_STEP_GFLOP = 6710.9    # C2: one (r, s) image pair through update_G + update_D of the masker + AdvEnt discriminators


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "gpu-eager"])
    ap.add_argument("--workload", default="full", choices=["full", "painter", "masker", "infer"],
                    help="full = Masker+Painter G+D train step (BASELINE.json metric; SURVEY.md §8d C3, 8 images/domain/GPU); "
                         "painter = C1 painter-only fwd+bwd (configs[1], 16 images/GPU); "
                         "masker = C2 masker + AdvEnt D train step (configs[2], 32 images/domain/GPU); "
                         "infer = C4 Trainer.infer_all (masker + painter + flood/wildfire/smog compositing, 16 images/GPU)")
    ap.add_argument("--batch", type=int, default=0, help="images per domain per GPU per step (default: 8 full, 16 painter)")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32", "fp16"])
    ap.add_argument("--no-graphs", action="store_true", help="time the eager train step instead of the CUDA-graph replay")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager-on-this-GPU (cuDNN) baseline leg")
    ap.add_argument("--eager-mode", default="tf32", choices=["tf32", "bf16"], help="--impl gpu-eager: fp32/TF32 or bf16 autocast")
    ap.add_argument("--topk", type=int, default=int(os.environ.get("CGB_TOPK", "8")))
    ap.add_argument("--cpu-sample-batch", type=int, default=0, help="CPU baseline sample batch (default: 2 full, 1 painter)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.batch <= 0:
        a.batch = {"full": 8, "masker": 32}.get(a.workload, 16)
    if a.workload == "infer":
        a.no_cpu_baseline = True   # the CPU arm of this workload is the reference's own infer_all, which cannot travel
        a.no_gpu_eager = True
    if a.cpu_sample_batch <= 0:
        a.cpu_sample_batch = 1 if a.workload == "painter" else 2
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
            "masker": "masker_train_step_images_per_sec", "infer": "infer_all_images_per_sec"}[args.workload]


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
    if args.workload == "masker":
        return {
            "workload": f"C2 Masker + AdvEnt discriminators train step (Trainer.update_G + update_D, tasks d,s,m; deeplabv2 ResNet-101 "
                        f"encoder + DADA depth + DeepLab-v2 seg + base mask decoders, D[m] + D[s]), {args.batch} images per domain (r,s) "
                        f"per GPU, {args.size}x{args.size}",
            "batch_per_domain_per_gpu": args.batch, "domains": ["r", "s"], "size": args.size,
            "images_per_sec_convention": "per-domain images/s (reference batch_size convention)",
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
# Baseline arms: the oracle port of the reference algorithm — on the host cores (CPU baseline / --impl reference) and, as the
# "PyTorch eager on this GPU" bar SURVEY.md §8(d) asks for, on the B200 through ATen / cuDNN (--impl gpu-eager)
# ------------------------------------------------------------------------------------------------
def _oracle_step_fn(workload: str, batch: int, size: int, device: str):
    """(one_step, sample description): one forward + backward of the workload through oracle/ (plain PyTorch ops on
    reference-layout state_dicts) on `device`; the optimiser's elementwise update is not included (< 1 %)."""
    from climategan_b200.utils import default_painter_opts, full_opts, synth_batch

    torch.manual_seed(0)
    dev = torch.device(device)
    z = size // 2 ** 7

    def grads_on(sd, skip):
        out = {}
        for k, v in sd.items():
            v = v.detach().clone().to(dev)
            if v.dtype.is_floating_point and not k.endswith(skip):
                v.requires_grad_(True)
            out[k] = v
        return out

    if workload == "painter":
        from climategan_b200.painter import PainterSpadeDecoder
        from oracle import painter_oracle as po  # baseline legs only

        sd = grads_on(PainterSpadeDecoder(default_painter_opts()).state_dict(), ("_u", "_v"))   # parameter container only
        x = (torch.rand(batch, 3, size, size) * 2 - 1).to(dev)
        m = (torch.rand(batch, 1, size, size) > 0.5).float().to(dev)
        t = (torch.rand(batch, 3, size, size) * 2 - 1).to(dev)

        def one_step():
            for v in sd.values():
                v.grad = None
            out = po.paint(sd, m, x, z, z, po.n_up_spades_of(sd))
            torch.nn.functional.l1_loss(out, t).backward()

        return one_step, f"oracle (PyTorch fp32 restatement of the reference) paint+L1+backward, batch {batch} at {size}x{size}"

    from climategan_b200.discriminator import OmniDiscriminator
    from climategan_b200.generator import OmniGenerator
    from climategan_b200.losses import Vgg19
    from oracle import full_step_oracle as fo  # baseline legs only
    from oracle.painter_oracle import SNState

    tasks = ("d", "s", "m", "p") if workload == "full" else ("d", "s", "m")
    opts = full_opts(nblocks=(3, 4, 23, 3), size=size, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3, tasks=tasks)
    gsd = grads_on(OmniGenerator(opts, latent_shape=(size, size)).state_dict(), ("_u", "_v", "running_mean", "running_var"))
    dsd = grads_on(OmniDiscriminator(opts).state_dict(), ("_u", "_v"))
    vsd = {k: v.detach().clone().to(dev) for k, v in Vgg19().state_dict().items()} if workload == "full" else None
    mdb = synth_batch(opts, batch, size, 1)
    mdb = {dom: {**b, "data": {k: v.to(dev) for k, v in b["data"].items()}} for dom, b in mdb.items()}

    def one_step():
        for v in list(gsd.values()) + list(dsd.values()):
            v.grad = None
        if workload == "full":
            loss, _ = fo.full_g_loss(gsd, dsd, vsd, mdb, z)
            loss.backward()
            ld, _ = fo.full_d_loss(gsd, dsd, mdb, z)
            ld.backward()
        else:
            loss, _ = fo.masker_g_loss(gsd, dsd, mdb, SNState(gsd), SNState(dsd))
            loss.backward()
            md = fo.masker_d_loss(gsd, dsd, mdb, SNState(gsd), SNState(dsd))
            (md["m"] + md["s"]).backward()

    doms = "(r,s,rf)" if workload == "full" else "(r,s)"
    what = "Trainer.update_G/update_D losses + backward" + ("" if workload == "full" else ", tasks d,s,m")
    return one_step, (f"oracle (PyTorch fp32 restatement of the reference's {what}), {batch} images per domain {doms} at "
                      f"{size}x{size}" + (" (2 is the minimum: train-mode BatchNorm)" if batch == 2 else ""))


def cpu_rate(args, steps, warmup):
    """img/s of the reference algorithm on the host cores (oracle port, all threads): (rate, s/step, cores, sample)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = args.cpu_sample_batch
    one_step, sample = _oracle_step_fn(args.workload, batch, args.size, "cpu")
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    mean_t = sum(times) / len(times)
    return batch / mean_t, mean_t, cores, sample + f", {steps} timed step(s), {warmup} warm-up", batch


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port), all host threads, on a bounded sample of the workload.  The
    line states the batch it ACTUALLY ran (`config.batch_per_domain_per_gpu`), not the GPU arm's.  Under torchrun (N > 1) rank 0
    alone runs: the value is ONE CPU process whatever N is (`n_gpus` repeats the launch's N for the driver's bookkeeping)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "infer":
        print(json.dumps({"impl": "reference", "unavailable": "the infer workload's CPU arm is the reference's own Trainer.infer_all, "
                          "which needs /root/reference (absent on the GPU box); its parity is pinned by tests/golden/infer_all.*"}))
        return
    heavy = args.workload in ("full", "masker")
    steps = max(1, min(args.steps, 2 if heavy else 3))
    warmup = 1
    rate, t, cores, sample, batch = cpu_rate(args, steps, warmup)
    cfg_args = argparse.Namespace(**{**vars(args), "batch": batch})
    cfg = workload_config(cfg_args)
    cfg["note"] = (f"CPU arm: bounded sample of the GPU arm's workload — {batch} image(s) per domain instead of {args.batch}; img/s is "
                   f"per-image and comparable.  One CPU process on {cores} host threads regardless of --gpus.")
    line = {
        "impl": "reference", "metric": metric_name(args), "value": rate, "unit": "img/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": rate, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu_eager(args):
    """--impl gpu-eager: the same oracle port executed by PyTorch eager ON THIS GPU (ATen / cuDNN kernels, the reference's own
    execution model) — the "beat this" bar next to the CPU arm (last row).  fp32 with TF32 tensor cores
    (cudnn.allow_tf32, the PyTorch default for convolutions) or bf16 autocast.  Halves the batch on out-of-memory."""
    dev = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    batch = args.batch
    err = None
    while batch >= 2 or (args.workload == "painter" and batch >= 1):
        try:
            torch.set_default_device(dev)   # the oracle builds its small constant tensors with torch.tensor(...)
            one_step, sample = _oracle_step_fn(args.workload, batch, args.size, dev)
            ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if args.eager_mode == "bf16" else (lambda: torch.autocast("cuda", enabled=False))
            steps, warmup = max(1, min(args.steps, 3)), 2
            for _ in range(warmup):
                with ctx():
                    one_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                with ctx():
                    one_step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            print(json.dumps({"impl": "gpu-eager", "metric": metric_name(args), "value": batch / (ms / 1e3), "unit": "img/s",
                              "ms_per_step": ms, "batch": batch, "steps": steps, "warmup": warmup,
                              "dtype": "fp32 storage, TF32 tensor cores (cudnn.allow_tf32)" if args.eager_mode == "tf32" else "bf16 autocast",
                              "sample": sample + " — PyTorch eager (ATen/cuDNN) on this GPU",
                              "max_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
            return
        except torch.cuda.OutOfMemoryError as e:   # noqa: PERF203
            err = str(e).splitlines()[0]
            batch //= 2
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"impl": "gpu-eager", "unavailable": f"{type(e).__name__}: {str(e)[:300]}"}), flush=True)
            return
    print(json.dumps({"impl": "gpu-eager", "unavailable": f"out of memory down to batch 2: {err}"}), flush=True)


def gpu_eager_baseline(args):
    """Run both eager modes in child processes (their memory is returned before our arm starts) and parse their lines."""
    out = {}
    for mode in ("tf32", "bf16"):
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "gpu-eager", "--workload", args.workload, "--batch", str(args.batch),
               "--size", str(args.size), "--steps", "3", "--eager-mode", mode]
        try:
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0"})
            lines = [ln for ln in res.stdout.strip().splitlines() if ln.startswith("{")]
            out[mode] = json.loads(lines[-1]) if lines else {"unavailable": (res.stderr or "no output")[-300:]}
        except Exception as e:  # noqa: BLE001
            out[mode] = {"unavailable": f"{type(e).__name__}: {e}"}
    return out


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
    """C3 (tasks d,s,m,p) and C2 (`--workload masker`: tasks d,s,m) train steps through the Trainer."""
    from climategan_b200.trainer import Trainer
    from climategan_b200.utils import full_opts, synth_batch

    B, S = args.batch, args.size
    torch.manual_seed(0)  # identical initial weights on every rank
    tasks = ("d", "s", "m", "p") if args.workload == "full" else ("d", "s", "m")
    opts = full_opts(nblocks=(3, 4, 23, 3), size=S, latent=640, n_up=7, ndf=64, n_layers=4, num_d=3, tasks=tasks)
    opts.dis.soft_shift, opts.dis.flip_prob = 0.2, 0.05   # defaults.yaml:194-195 (label smoothing / flipping on, as in training)
    t = Trainer(opts, device=dev, storage_dtype=dtype).setup(input_shape=(S, S))
    mdb = synth_batch(opts, B, S, seed=1234 + rank)       # each rank its own slice of the global batch
    host = {dom: {k: v.pin_memory() for k, v in b["data"].items()} for dom, b in mdb.items()}
    if world > 1:
        t.enable_data_parallel()
    if not args.no_graphs:
        t.enable_cuda_graphs()

    def step(batch):
        t.update_G(batch)
        t.update_D(batch)
        t.logger.global_step += 1
        return t.logger.losses.gen.total_loss

    def to_dev(nb):
        return [{dom: {"data": {k: v.to(dev, non_blocking=nb) for k, v in d.items()}, "domain": [dom] * B, "mode": ["train"] * B,
                       "paths": {}} for dom, d in host.items()}]

    h2d = int(sum(v.numel() * v.element_size() for d in host.values() for v in d.values()))
    step.trainer = t

    def host_batches(count):   # what zip(*loaders) yields: pinned host batches (the same synthetic batch every step)
        for _ in range(count):
            yield [{dom: {"data": dict(d), "domain": [dom] * B, "mode": ["train"] * B, "paths": {}} for dom, d in host.items()}]

    step.host_batches = host_batches
    return step, to_dev, h2d, FULL_STEP_GFLOP if args.workload == "full" else MASKER_STEP_GFLOP


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
    if args.impl == "gpu-eager":
        run_gpu_eager(args)
        return

    from climategan_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device()
    lib = _lib.lib()

    # ---- PyTorch-eager-on-this-GPU bar (child processes, before our arm allocates; rank 0 at N=1 only)
    eager = None
    if world == 1 and not args.no_gpu_eager:
        eager = gpu_eager_baseline(args)

    dist = None
    if world > 1:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    dtype = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[args.dtype]
    B = args.batch

    builder = {"full": build_full, "masker": build_full, "painter": build_painter, "infer": build_infer}[args.workload]
    step, to_dev, h2d_bytes, gflop_per_image = builder(args, dev, rank, world, dtype)
    trainer = getattr(step, "trainer", None)
    graphs_on = trainer is not None and not args.no_graphs
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

    import ctypes as C

    # ---- warm-up: with graphs on, step 0 runs eagerly, step 1 captures + replays, later steps replay
    warm = max(args.warmup, 3) if graphs_on else args.warmup
    for _ in range(warm):
        step(*resident)
    barrier()

    # ---- device-resident timed region (value), clocks sampled during it
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms = timed(lambda: step(*resident), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step / 1e3)

    # ---- end-to-end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        host_batches = getattr(step, "host_batches", None)
        if host_batches is not None:
            # the train workloads: every step's batch goes pinned host -> device through climategan_b200.data.DevicePrefetcher
            # (side-stream copy two batches deep, the loader edge of data.py:506-539 / trainer.py:609-621), INSIDE the timed region
            from climategan_b200.data import DevicePrefetcher

            feed = DevicePrefetcher(host_batches(2 * args.steps + 4), dev, depth=2)

            def e2e_step():
                res = step(*next(feed))
                return float(res.item())                                                   # D2H read of the step result
        else:
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

    # ---- ONE eager step bracketed launch by launch with CUDA events (outside the timed region): the per-class conv table and
    #      the count of library launches a step consists of (what the graph replays)
    if graphs_on:
        trainer.enable_cuda_graphs(False)
    step(*resident)
    barrier()
    lib.cgb_prof_enable(1)
    step(*resident)                       # builds the event pool
    barrier()
    lib.cgb_prof_enable(0)
    _scratch = C.create_string_buffer(1 << 22)
    lib.cgb_prof_dump(_scratch, len(_scratch))
    lib.cgb_launch_count_reset()
    lib.cgb_prof_enable(1)
    prof_ms = timed(lambda: step(*resident), 1)
    lib.cgb_prof_enable(0)
    launches_per_step = int(lib.cgb_launch_count())
    buf = C.create_string_buffer(1 << 22)
    lib.cgb_prof_dump(buf, len(buf))
    prof = [ln.split() for ln in buf.value.decode().strip().splitlines() if ln.strip()]
    eager_ms = timed(lambda: step(*resident), 2) / 2   # the eager (no graph, no per-launch events) step, for the record

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines
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
    tpeak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    hpeak = peaks["hbm_gbs"]
    esz = 4 if args.dtype == "fp32" else 2

    def describe(r):
        """Roofline entry of one (op, shape) class.  The bound is decided by arithmetic intensity: algorithmic FLOPs
        over the bytes the launch must move (operands read once + result written once, storage dtype) against the
        machine ridge (measured bf16 peak / measured HBM bandwidth)."""
        avg_ms = r["total_ms"] / r["count"]
        flops = _logical_flops(r)
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
        common = {"kernel": name, "avg_launch_ms": avg_ms, "launches_per_step": r["count"], "share_of_conv_time": r["total_ms"] / conv_ms,
                  "share_of_step": r["total_ms"] / prof_ms, "algorithmic_flops_per_launch": flops,
                  "algorithmic_bytes_per_launch": byts, "traffic": traffic.get(tkey)}
        if flops / byts >= ridge:
            a = flops / (avg_ms * 1e-3) / 1e12
            return dict(bound="tensor", achieved=a, peak=tpeak, unit="TFLOP/s", frac=a / tpeak,
                        peak_source=f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)", **common)
        a = byts / (avg_ms * 1e-3) / 1e9
        return dict(bound="hbm", achieved=a, peak=hpeak, unit="GB/s", frac=a / hpeak,
                    peak_source=f"{peak_src} hbm_gbs", **common)

    classes = [describe(r) for r in rows[:max(args.topk, 8)]]
    step_tflops = world * B * gflop_per_image * 1e9 / (ms_per_step * 1e-3) / 1e12
    conv_flops = sum(_logical_flops(r) * r["count"] for r in rows)
    conv_tflops = conv_flops / (conv_ms * 1e-3) / 1e12
    roofline = None
    if classes:
        roofline = dict(classes[0])   # the (op, shape) class with the largest share of the step
        roofline.update({
            "note": "dominant = the conv class with the largest share of the step; no class exceeds a few % of it, so the step-level "
                    "and conv-aggregate fractions below are the ones that describe the path",
            "step_frac": step_tflops / (world * tpeak),
            "step_tflops_algorithmic": step_tflops,
            "conv_aggregate": {"tflops": conv_tflops, "frac": conv_tflops / tpeak, "conv_ms_per_step": conv_ms,
                               "share_of_eager_profiled_step": conv_ms / prof_ms},
            "top_classes": [{k: c[k] for k in ("kernel", "bound", "achieved", "unit", "frac", "avg_launch_ms", "launches_per_step",
                                                "share_of_step")} for c in classes[:8]],
        })

    cpu = None
    if not args.no_cpu_baseline and world == 1:   # the CPU baseline is reported at N=1 only
        heavy = args.workload in ("full", "masker")
        rate, t, cores, sample, cb = cpu_rate(args, 1 if heavy else 2, 1)
        cpu = {"value": rate, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample + f" ({t:.1f} s/step)"}

    eager_line = None
    if eager is not None:
        best = max((v for v in eager.values() if "value" in v), key=lambda v: v["value"], default=None)
        eager_line = {"value": best["value"] if best else None, "unit": "img/s", "dtype": best["dtype"] if best else None,
                      "ms_per_step": best["ms_per_step"] if best else None, "batch": best["batch"] if best else None,
                      "what": "the oracle port (plain PyTorch ops, reference layout) run by PyTorch eager on this B200 through "
                              "ATen/cuDNN — the faster of fp32+TF32 and bf16 autocast; no Blackwell kernel ships in the reference",
                      "modes": eager}

    line = {
        "metric": metric_name(args), "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {**workload_config(args), "execution": ("CUDA-graph replay of zero_grad+forward+backward per update; optimiser + "
                                                          "all-reduce eager" if graphs_on else "eager")},
        "clocks": clocks, "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
        "gpu_launches_note": "library kernel launches one step consists of (counted on an eager step); with graphs on, the timed "
                             "region replays exactly these as graph nodes",
        "roofline": roofline, "cpu_baseline": cpu, "gpu_eager_baseline": eager_line,
        "step_tflops_algorithmic": step_tflops,
        "step_frac_of_bf16_peak": step_tflops / (world * tpeak),
        "eager_ms_per_step": eager_ms, "profiled_eager_ms_per_step": prof_ms,
        "top_kernels": [
            {"kernel": f"{r['op']}[{r['engine']}] {r['ci']}->{r['co']} k{r['k']} s{r['stride']} d{r['dil']} @{r['hi']}x{r['wi']}",
             "count": r["count"], "total_ms": round(r["total_ms"], 3)} for r in rows[:args.topk]],
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
