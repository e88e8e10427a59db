#!/usr/bin/env python
"""apply_events.py — the inference CLI of the reference (``/apply_events.py:4-148`` flags, ``:430-640`` main loop) on top of
``climategan_b200.trainer.Trainer``: read a folder of images, resize/crop them to the target size, run
``Trainer.infer_all`` (masker + painter + flood / wildfire / smog compositing on the B200) in batches and write
``{stem}_{event}_{width}{suffix}.png`` (``:616``).

Same flags as the reference.  Host-side image I/O uses PIL (the reference uses skimage: bilinear + anti-aliasing resize, so
the pre-processed pixels can differ in the last bits); ``--half`` runs the generator with fp16 storage (the reference's
``trainer.G.half()``, :465-468; bf16 otherwise, fp32 with ``--fp32``); ``--upload`` (comet.ml) is out of scope and refused.  Everything between the pre-processed batch and the uint8
events runs through libcgb200 — there is no CPU fallback.
"""
from __future__ import annotations

import argparse
import shutil
import sys
import time
from pathlib import Path


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("-b", "--batch_size", type=int, default=4)
    p.add_argument("-i", "--images_paths", type=str, required=True, help="Path to a directory with image files")
    p.add_argument("-o", "--output_path", type=str, default=None)
    p.add_argument("-s", "--save_input", action="store_true", default=False)
    p.add_argument("-r", "--resume_path", type=str, default=None, help="directory with opts.yaml and checkpoints/")
    p.add_argument("--no_time", action="store_true", default=False)
    p.add_argument("-f", "--flood_mask_binarization", type=float, default=0.5)
    p.add_argument("-t", "--target_size", type=int, default=640)
    p.add_argument("--half", action="store_true", default=False)
    p.add_argument("-n", "--n_images", default=-1, type=int)
    p.add_argument("--no_conf", action="store_true", default=False)
    p.add_argument("--overwrite", action="store_true", default=False)
    p.add_argument("--no_cloudy", action="store_true", default=False)
    p.add_argument("--keep_ratio_128", action="store_true", default=False)
    p.add_argument("--fuse", action="store_true", default=False)
    p.add_argument("--cpu_preprocess", action="store_true", default=False,
                   help="resize / crop on the host with PIL instead of on the GPU (not a reference flag)")
    p.add_argument("--save_masks", action="store_true", default=False)
    p.add_argument("-m", "--max_im_width", type=int, default=-1)
    p.add_argument("--upload", action="store_true")
    p.add_argument("--zip_outdir", "-z", action="store_true")
    p.add_argument("--fp32", action="store_true", help="(this implementation) fp32 storage / SIMT engine: the parity mode")
    return p.parse_args()


IMG_EXT = {".png", ".jpg", ".jpeg", ".bmp", ".tif", ".tiff", ".webp"}


def find_images(path):
    p = Path(path).expanduser().resolve()
    assert p.exists(), f"{p} does not exist"
    if p.is_file():
        return [p]
    return sorted(q for q in p.glob("**/*") if q.suffix.lower() in IMG_EXT)


def to_128(h, w, w_max=-1):
    """apply_events.py:150-183: closest multiples of 128 keeping the aspect ratio, width capped by w_max."""
    if 0 < w_max < w:
        h, w = int(h * w_max / w), w_max
    return max(128, round(h / 128) * 128), max(128, round(w / 128) * 128)


def read_image_u8(path):
    """The decoded photograph, uint8 HWC RGB (RGBA is composited on white like skimage.color.rgba2rgb, apply_events.py:491)."""
    import numpy as np
    from PIL import Image

    im = Image.open(path)
    if im.mode == "RGBA":
        bg = Image.new("RGBA", im.size, (255, 255, 255, 255))
        im = Image.alpha_composite(bg, im)
    return np.asarray(im.convert("RGB"), dtype=np.uint8)


def load_image(path, target, keep_ratio, max_w):
    """HOST version of the input edge (PIL), kept for --cpu_preprocess: -> float32 HWC in [-1, 1] (resize_and_crop / to_m1_p1
    semantics: short side to `target`, centre crop).  The default path resizes on the GPU (events.InputEdge)."""
    import numpy as np
    from PIL import Image

    im = Image.open(path).convert("RGB")
    w, h = im.size
    if keep_ratio:
        nh, nw = to_128(h, w, max_w)
        im = im.resize((nw, nh), Image.BILINEAR, reducing_gap=3.0)
    else:
        s = target / min(h, w)
        nh, nw = max(target, round(h * s)), max(target, round(w * s))
        im = im.resize((nw, nh), Image.BILINEAR, reducing_gap=3.0)
        top, left = (nh - target) // 2, (nw - target) // 2
        im = im.crop((left, top, left + target, top + target))
    a = np.asarray(im, dtype=np.float32) / 255.0
    return a * 2.0 - 1.0


def main():
    args = parse_args()
    if args.upload:
        sys.exit("--upload (comet.ml) is not built: out of scope of the hot path")
    import numpy as np
    import torch
    from PIL import Image

    from climategan_b200.bn_fusion import bn_fuse
    from climategan_b200.trainer import Trainer

    outdir = None
    if args.output_path is not None:
        outdir = Path(args.output_path).expanduser().resolve()
        if outdir.exists() and any(outdir.iterdir()) and not args.overwrite:
            sys.exit(f"{outdir} exists and is not empty (use --overwrite)")
        outdir.mkdir(parents=True, exist_ok=True)
    assert args.resume_path, "-r/--resume_path (a directory with opts.yaml and checkpoints/latest_ckpt.pth) is required"
    t0 = time.perf_counter()
    torch.set_grad_enabled(False)
    trainer = Trainer.resume_from_path(args.resume_path, setup=True, inference=True, new_exp=None,
                                       storage_dtype=torch.float32 if args.fp32 else torch.bfloat16,
                                       input_shape=(args.target_size, args.target_size))
    if args.fuse:
        trainer.G = bn_fuse(trainer.G)
    t_setup = time.perf_counter() - t0

    paths = find_images(args.images_paths)
    base = list(paths)
    assert paths, f"no images under {args.images_paths}"
    if 0 < args.n_images < len(paths):
        paths = paths[: args.n_images]
    elif args.n_images > len(paths):
        paths = (base * (args.n_images // len(base) + 1))[: args.n_images]
    t0 = time.perf_counter()
    if args.cpu_preprocess:
        data = [load_image(p, args.target_size, args.keep_ratio_128, args.max_im_width) for p in paths]
        sizes = [d.shape[:2] for d in data]
    else:
        # decode on the host, resize / crop / rescale on the GPU (events.InputEdge: one kernel per image, pinned double-buffered H2D)
        from climategan_b200.events import InputEdge

        edge = InputEdge(trainer.device)
        data = [read_image_u8(p) for p in paths]
        sizes = [to_128(d.shape[0], d.shape[1], args.max_im_width) if args.keep_ratio_128 else (args.target_size, args.target_size)
                 for d in data]
    t_pre = time.perf_counter() - t0
    print("Found", len(base), "images. Inferring on", len(data), "images.")

    # batches must hold images of one size (keep_ratio_128 produces several): group consecutive equal shapes
    all_events, t0 = [], time.perf_counter()
    i = 0
    while i < len(data):
        j = i
        while j < len(data) and j - i < args.batch_size and sizes[j] == sizes[i]:
            j += 1
        if args.cpu_preprocess:
            images = np.stack(data[i:j])
        else:
            images = edge(data[i:j], args.target_size, keep_ratio_sizes=sizes[i:j] if args.keep_ratio_128 else None)
        ev = trainer.infer_all(images, numpy=True, bin_value=args.flood_mask_binarization, half=args.half,
                               cloudy=not args.no_cloudy, return_masks=args.save_masks)
        if args.save_input:
            src = images if args.cpu_preprocess else images.permute(0, 2, 3, 1).cpu().numpy()
            ev["input"] = ((src + 1) / 2 * 255).astype(np.uint8)
        all_events.append(ev)
        i = j
    torch.cuda.synchronize()
    t_inf = time.perf_counter() - t0

    if outdir is not None:
        k = 0
        for ev in all_events:
            names = [n for n in ev if ev[n] is not None]
            for b in range(len(ev[names[0]])):
                stem = Path(paths[k % len(paths)]).stem
                width = sizes[k][1]
                suffix = ("_AR" if args.keep_ratio_128 else "") + ("_no_cloudy" if args.no_cloudy else "")
                for name in names:
                    im = ev[name][b]
                    if name == "mask":
                        im = im[0]
                    Image.fromarray(im).save(outdir / f"{stem}_{name}_{width}{suffix}.png")
                k += 1
        if not args.no_conf:
            (outdir / "apply_events_command.txt").write_text(" ".join(sys.argv) + "\n")
        if args.zip_outdir:
            arch = Path(shutil.make_archive(outdir.name, "zip", root_dir=outdir))
            arch.rename(outdir.parent / arch.name)
    if not args.no_time:
        n = len(data)
        print(f"setup {t_setup:.2f} s | pre-processing {t_pre:.2f} s | inference {t_inf:.3f} s ({n / max(t_inf, 1e-9):.1f} img/s)")


if __name__ == "__main__":
    main()
