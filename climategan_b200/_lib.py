"""ctypes binding of libcgb200.so (the C ABI declared in include/cgb200.h).

The shared library is built in-tree by :func:`build` (nvcc, sm_100a only) and loaded by
:func:`lib`.  There is no fallback of any kind: if the library is missing, or the device is not
an sm_100 part, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_CSRC = _PKG / "csrc"
_SO = _PKG / "libcgb200.so"
_SOURCES = ["api.cu", "ops.cu", "masker_ops.cu", "events.cu", "conv_simt.cu", "conv_tc.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "--threads", "6",
    "-Xcompiler", "-fPIC", "-shared",
]

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH, ACT_SIGMOID, ACT_SELU = 0, 1, 2, 3, 4, 5
PAD_ZERO, PAD_REFLECT = 0, 1
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1, 2


class ConvDesc(C.Structure):
    """Mirror of ``cgb_conv_desc`` (include/cgb200.h)."""

    _fields_ = [
        ("n", C.c_int32), ("hi", C.c_int32), ("wi", C.c_int32), ("ci", C.c_int32),
        ("ho", C.c_int32), ("wo", C.c_int32), ("co", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32),
        ("stride", C.c_int32), ("dil", C.c_int32), ("pad", C.c_int32),
        ("pad_mode", C.c_int32), ("dtype", C.c_int32), ("act", C.c_int32),
        ("slope", C.c_float), ("engine", C.c_int32), ("res_before_act", C.c_int32),
    ]


def _needs_build() -> bool:
    if not _SO.exists():
        return True
    so_m = _SO.stat().st_mtime
    deps = [_CSRC / s for s in _SOURCES] + list(_CSRC.glob("*.cuh")) + [_PKG.parent / "include" / "cgb200.h"]
    return any(d.stat().st_mtime > so_m for d in deps if d.exists())


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/*.cu into climategan_b200/libcgb200.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not _needs_build():
        return _SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(_SO), *[str(_CSRC / s) for s in _SOURCES]]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed:\n{res.stdout}\n{res.stderr}")
    return _SO


_P = C.c_void_p
_I = C.c_int32
_F = C.c_float
_L = C.c_int64
_DP = C.POINTER(ConvDesc)

# name -> argtypes; every symbol declared in include/cgb200.h (tests check the two lists agree)
SIGNATURES = {
    "cgb_version": ([], C.c_char_p),
    "cgb_last_error": ([], C.c_char_p),
    "cgb_device_ok": ([], C.c_int),
    "cgb_launch_count": ([], C.c_int64),
    "cgb_launch_count_reset": ([], None),
    "cgb_prof_enable": ([C.c_int], None),
    "cgb_prof_dump": ([C.c_char_p, C.c_int64], C.c_int),
    "cgb_conv2d_uses_tcgen05": ([_DP, C.c_int], C.c_int),
    "cgb_conv2d_fwd": ([_DP, _P, _P, _P, _P, _P, _P], C.c_int),
    "cgb_conv2d_stats_rows": ([], C.c_int32),
    "cgb_conv2d_fwd_stats": ([_DP, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "cgb_conv2d_dgrad": ([_DP, _P, _P, _P, _I, _P, _P, _P], C.c_int),
    "cgb_conv2d_pack_dgrad_weight": ([_DP, _P, _P, _P], C.c_int),
    "cgb_pack_weight": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_pack_weight_dual": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_conv2d_wgrad": ([_DP, _P, _P, _P, _P, _I, _P], C.c_int),
    "cgb_instnorm_ws_doubles": ([_I, _I, _I], C.c_int64),
    "cgb_instnorm_stats": ([_P, _I, _I, _I, _I, _F, _P, _P, _P, _P], C.c_int),
    "cgb_spade_modulate_fwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_spade_modulate_bwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_spade_modulate_bwd_bias": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_instnorm_bwd": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _P], C.c_int),
    "cgb_instnorm_apply_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_instnorm_apply_bwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_avgpool3s2_fwd": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_avgpool3s2_bwd": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_const_target_loss": ([_P, _P, _P, _L, _I, _F, _F, _P], C.c_int),
    "cgb_const_target_loss_dev": ([_P, _P, _P, _L, _I, _P, _F, _P], C.c_int),
    "cgb_l1_loss_storage": ([_P, _P, _P, _P, _I, _L, _F, _P], C.c_int),
    "cgb_vgg_preprocess_fwd": ([_P, _P, _P, _I, _I, _I, _P], C.c_int),
    "cgb_vgg_preprocess_bwd": ([_P, _P, _P, _I, _I, _I, _P], C.c_int),
    "cgb_maxpool2_fwd": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_maxpool2_bwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_extra_adam": ([_P, _P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _I, _I, _P], C.c_int),
    "cgb_maxpool3s2_ceil_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_resize_bilinear_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_resize_bicubic_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_channel_mean": ([_P, _P, _I, _L, _I, _I, _P], C.c_int),
    "cgb_mul": ([_P, _P, _P, _I, _L, _P], C.c_int),
    "cgb_make_m_cond": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_make_m_cond_bwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_bn_apply_fwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P], C.c_int),
    "cgb_bn_apply_bwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P], C.c_int),
    "cgb_bn_bwd_ws_doubles": ([_L, _I], C.c_int64),
    "cgb_bn_bwd_finalize": ([_P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _P], C.c_int),
    "cgb_bn_train_fwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _F, _I, _F, _P], C.c_int),
    "cgb_bn_train_fwd_partials": ([_P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _F, _I, _F, _P], C.c_int),
    "cgb_bn_train_bwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P], C.c_int),
    "cgb_bn_train_bwd2": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _F, _P], C.c_int),
    "cgb_bn_update_running": ([_P, _P, _P, _P, _I, _L, _F, _F, _P], C.c_int),
    "cgb_maxpool3s2_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_maxpool3s2_bwd": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_maxpool3s2_ceil_bwd": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_resize_bilinear_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_reflect_pad_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_reflect_pad_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_replicate_pad_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_replicate_pad_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_affine_nc_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_affine_nc_bwd": ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_channel_mean_bwd": ([_P, _P, _I, _L, _I, _I, _P], C.c_int),
    "cgb_broadcast_hw": ([_P, _P, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_dropout": ([_P, _P, _I, _L, _F, C.c_uint64, _P], C.c_int),
    "cgb_dropout_dev": ([_P, _P, _I, _L, _F, _P, _P], C.c_int),
    "cgb_softmax_nchw_fwd": ([_P, _P, _I, _I, _I, _P], C.c_int),
    "cgb_softmax_nchw_bwd": ([_P, _P, _P, _I, _I, _I, _P], C.c_int),
    "cgb_cross_entropy_nchw": ([_P, _P, _P, _P, _I, _I, _I, _P], C.c_int),
    "cgb_entropy_nchw": ([_P, _P, _P, _P, _I, _I, _I, _I, _P], C.c_int),
    "cgb_minent_loss": ([_P, _P, _P, _P, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_sigmoid_pair": ([_P, _P, _P, _I, _I, _I, _P], C.c_int),
    "cgb_tv_loss": ([_P, _P, _P, _I, _I, _I, _I, _P], C.c_int),
    "cgb_bce_logits_loss": ([_P, _P, _P, _P, _L, _P], C.c_int),
    "cgb_ground_intersection_loss": ([_P, _P, _P, _L, _P], C.c_int),
    "cgb_sigm_loss": ([_P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P], C.c_int),
    "cgb_dada_depth_loss": ([_P, _P, _P, _P, _L, _P], C.c_int),
    "cgb_diff_aug_sum": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_diff_aug_fwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_diff_aug_bwd": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_argmax_confusion": ([_P, _P, _P, _P, _I, _I, _L, _P], C.c_int),
    "cgb_minmax_per_sample": ([_P, _P, _I, _L, _P], C.c_int),
    "cgb_fire_tone": ([_P, _P, _P, _P, _I, _I, _F, _F, _P], C.c_int),
    "cgb_sky_mask": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_plane_resize_nearest": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_box_dilate": ([_P, _P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_gauss_blur": ([_P, _P, _P, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_fire_paste": ([_P, _P, _P, _I, _I, _I, _F, _F, _F, _F, _F, _P], C.c_int),
    "cgb_smog": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _F, _F, _F, _F, _P], C.c_int),
    "cgb_perlin_noise": ([_P, _P, _I, _I, _I, _I, _P], C.c_int),
    "cgb_cloudy_mix": ([_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P], C.c_int),
    "cgb_to_uint8_nhwc": ([_P, _P, _P, _I, _I, _P], C.c_int),
    "cgb_resize_crop_u8": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_mask_to_uint8": ([_P, _P, _F, _L, _P], C.c_int),
    "cgb_resize_nearest_fwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_upsample_nearest_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_resize_nearest_bwd": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_im2col_strided": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_im2col": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_col2im_strided": ([_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_nchw_to_nhwc": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_nhwc_to_nchw": ([_P, _P, _I, _I, _I, _I, _I, _P], C.c_int),
    "cgb_act_bwd": ([_P, _P, _P, _I, _L, _I, _F, _P], C.c_int),
    "cgb_act_bwd_bias": ([_P, _P, _P, _P, _I, _L, _I, _I, _F, _P], C.c_int),
    "cgb_act_fwd": ([_P, _P, _I, _L, _I, _F, _P], C.c_int),
    "cgb_mask_cond": ([_P, _P, _P, _I, _I, _I, _I, _P], C.c_int),
    "cgb_paste_fwd": ([_P, _P, _P, _P, _I, _I, _P], C.c_int),
    "cgb_paste_bwd": ([_P, _P, _P, _I, _I, _P], C.c_int),
    "cgb_mask_cond_bwd": ([_P, _P, _P, _I, _I, _I, _I, _P], C.c_int),
    "cgb_paste_bwd_mask": ([_P, _P, _P, _P, _I, _I, _P], C.c_int),
    "cgb_l1_loss": ([_P, _P, _P, _P, _L, _F, _P], C.c_int),
    "cgb_spectral_power_iter": ([_P, _P, _P, _P, _I, _I, _P], C.c_int),
}

_lib = None


def lib() -> C.CDLL:
    """Load (building first if the sources are newer) and return the ctypes handle."""
    global _lib
    if _lib is not None:
        return _lib
    if not _SO.exists():
        # building needs nvcc; on a GPU box the prebuilt .so travels with the snapshot
        build()
    handle = C.CDLL(str(_SO))
    for name, (argtypes, restype) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = handle
    return _lib


class CgbError(RuntimeError):
    pass


def check(status: int, what: str = "") -> None:
    """Raise :class:`CgbError` for a negative cgb_status."""
    if status != 0:
        msg = lib().cgb_last_error().decode("utf-8", "replace")
        raise CgbError(f"libcgb200 {what} failed (status {status}): {msg}")


def require_device() -> None:
    """Fail loudly when the CUDA extension cannot run (no sm_100 device)."""
    if not lib().cgb_device_ok():
        raise CgbError(
            "libcgb200 requires an sm_100 (B200) CUDA device; there is no CPU or eager fallback"
        )
