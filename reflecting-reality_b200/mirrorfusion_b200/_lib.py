"""ctypes binding of libmfb200.so (the C ABI declared in include/mfb200.h).

There is no fallback: if the shared library is missing or the device is not sm_100 every
entry point raises.  Build it with `python __graft_entry__.py build` (nvcc, in-tree).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MFB200_LIB selects another build of the SAME library (kernel A/B variants, tools/ab_build.sh); never a fallback
LIB_PATH = os.environ.get("MFB200_LIB") or os.path.join(_HERE, "lib", "libmfb200.so")

_lib = None
_inited_device = None

vp, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("H", i32), ("W", i32), ("Cin", i32), ("Cout", i32), ("ksize", i32), ("stride", i32),
        ("x", vp), ("n_extra", i32), ("extra_x", vp * 3), ("extra_C", i32 * 3), ("w", vp), ("bias", vp),
        ("rowbias", vp), ("rowbias_ld", i32), ("alpha", vp), ("res1", vp), ("res2", vp), ("out", vp),
        ("geglu", i32), ("block_n", i32), ("up2x", i32), ("igemm_mode", i32), ("pad0", i32), ("dtype", i32),
    ]


_SIGS = {
    "mfb_abi_version": (i32, []),
    "mfb_init": (i32, [i32]),
    "mfb_last_error": (C.c_char_p, []),
    "mfb_program_begin": (i32, [C.POINTER(vp)]),
    "mfb_program_end": (i32, []),
    "mfb_program_size": (i32, [vp]),
    "mfb_program_run": (i32, [vp, vp]),
    "mfb_program_destroy": (i32, [vp]),
    "mfb_copy_f32": (i32, [vp, vp, i64, vp]),
    "mfb_conv_plan_create": (i32, [C.POINTER(ConvDesc), C.POINTER(vp)]),
    "mfb_plan_run": (i32, [vp, vp]),
    "mfb_plan_destroy": (i32, [vp]),
    "mfb_plan_flops": (f64, [vp]),
    "mfb_plan_ktotal": (i32, [vp]),
    "mfb_plan_igemm_mode": (i32, [vp]),
    "mfb_plan_launches": (i32, [vp]),
    "mfb_plan_stats_floats": (i64, [vp]),
    "mfb_plan_stats_tiles": (i32, [vp]),
    "mfb_plan_set_stats": (i32, [vp, vp]),
    "mfb_groupnorm_prestat": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, f32, vp, vp, i32, vp, vp, vp]),
    "mfb_groupnorm": (i32, [vp, i32, vp, i32, i32, i32, i32, f32, vp, vp, i32, vp, vp, vp]),
    "mfb_layernorm": (i32, [vp, i32, i32, f32, vp, vp, vp, vp]),
    "mfb_softmax_rows": (i32, [vp, i32, i32, vp, vp]),
    "mfb_attention": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mfb_transpose_tokens": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "mfb_conv_in": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp]),
    "mfb_conv_out": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "mfb_upsample2x": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "mfb_nchw_f32_to_nhwc_bf16": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "mfb_nhwc_bf16_to_nchw_f32": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "mfb_f32_to_bf16": (i32, [vp, i64, vp, vp]),
    "mfb_timestep_sinusoid": (i32, [vp, i32, i32, vp, vp]),
    "mfb_linear_small": (i32, [vp, i32, i32, vp, vp, i32, i32, i32, vp, vp]),
    "mfb_prep_image_u8": (i32, [vp, i32, i32, i32, vp, vp]),
    "mfb_prep_mask_depth": (i32, [vp, vp, i32, i32, i32, i32, f32, vp, vp, vp, vp]),
    "mfb_post_image_u8": (i32, [vp, i32, i32, i32, vp, vp]),
    "mfb_resize_crop_bicubic": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "mfb_depth_normalize": (i32, [vp, vp, i32, i32, i32, f32, vp, vp, vp]),
    "mfb_latent_sample": (i32, [vp, vp, vp, f32, vp, i64, vp]),
    "mfb_cfg_sched_step": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, i64, vp]),
    # fp32 parity mode
    "mfb_groupnorm_f32": (i32, [vp, i32, vp, i32, i32, i32, i32, f32, vp, vp, i32, vp, vp]),
    "mfb_layernorm_f32": (i32, [vp, i32, i32, f32, vp, vp, vp, vp]),
    "mfb_attention_f32": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mfb_conv_in_f32": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp]),
    "mfb_conv_out_f32": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "mfb_linear_small_f32": (i32, [vp, i32, i32, vp, vp, i32, i32, i32, vp, vp]),
    # training-step glue (config 4)
    "mfb_add_noise": (i32, [vp, vp, vp, vp, i32, i32, i64, vp, vp, vp]),
    "mfb_mse_loss": (i32, [vp, vp, vp, i32, i64, vp, vp, vp, vp, vp]),
    "mfb_grad_sqnorm": (i32, [vp, i64, vp, vp, i32, vp]),
    "mfb_adamw_step": (i32, [vp, vp, vp, vp, vp, i64, vp, vp, f32, vp]),
    "mfb_conv_wgrad": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp]),
    "mfb_conv_wgrad_tc_ws_floats": (i64, [i32, i32, i32, i32, i32, i32, i32]),
    "mfb_conv_wgrad_tc": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp, i64, vp]),
    "mfb_groupnorm_bwd": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, i32, f32, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp]),
    "mfb_rowsum_per_image": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "mfb_silu_bwd": (i32, [vp, vp, i32, vp, vp, i64, vp]),
    # bf16 backward of the bandwidth-bound ops (csrc/train_bf16.cu)
    "mfb_attention_lse": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
    "mfb_attention_bwd": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "mfb_groupnorm_bwd2_ws_floats": (i64, [i32, i32, i32]),
    "mfb_groupnorm_stats": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp]),
    "mfb_groupnorm_bwd2": (i32, [vp, i32, vp, i32, vp, i32, i32, i32, f32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, i32, vp]),
    "mfb_layernorm_bwd": (i32, [vp, vp, i32, i32, f32, vp, vp, vp, vp]),
    "mfb_geglu": (i32, [vp, i64, i32, vp, vp, vp, vp]),
    "mfb_conv_out_bwd": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "mfb_sumpool2x2": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "mfb_dgrad_repack": (i32, [vp, i32, i32, i32, vp, vp]),
    "mfb_conv_out_bwd_f32": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
    "mfb_sumpool2x2_f32": (i32, [vp, i32, i32, i32, i32, vp, vp]),
    "mfb_add_f32": (i32, [vp, vp, i64, vp]),
    # fp32 parity-mode backward of attention / LayerNorm / GEGLU
    "mfb_attention_bwd_f32": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "mfb_layernorm_bwd_f32": (i32, [vp, vp, i32, i32, f32, vp, vp, vp]),
    "mfb_geglu_f32": (i32, [vp, i64, i32, vp, vp, vp, vp]),
}
EXPORTS = tuple(_SIGS)


class MfbError(RuntimeError):
    pass


def load(init_device: int | None = None):
    """Load the library (no GPU needed) and optionally bind it to a device (GPU needed)."""
    global _lib, _inited_device
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MfbError(f"{LIB_PATH} not built; run `python __graft_entry__.py build`. There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.mfb_abi_version() != 1:
            raise MfbError("libmfb200.so ABI version mismatch")
        _lib = lib
    if init_device is not None and _inited_device != init_device:
        check(_lib.mfb_init(init_device))
        _inited_device = init_device
    return _lib


def check(rc: int):
    if rc != 0:
        msg = _lib.mfb_last_error().decode() if _lib is not None else "?"
        raise MfbError(f"libmfb200 error {rc}: {msg}")
