"""ctypes binding of libvcd_b200.so (include/vcd.h).  There is no CPU fallback: a missing library
or a failing kernel raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VCD_LIB_PATH: load another build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("VCD_LIB_PATH") or os.path.join(_HERE, "libvcd_b200.so")

F32, BF16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_UMMA = 0, 1, 2
WGRAD_OVERLAP_PREV = 0x100   # flag OR-ed into `impl` of vcd_conv2d_wgrad (include/vcd.h)
ACC_PREZEROED = 0x200        # flag: accumulator arguments are already zero, the call enqueues no memset (include/vcd.h)

_p, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); kept in the order of include/vcd.h
SIGNATURES = {
    "vcd_last_error": (C.c_char_p, []),
    "vcd_version": (_i, []),
    "vcd_conv_umma_supported": (_i, [_i] * 5),
    "vcd_pack_conv_weight": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vcd_pack_tile_co": (_i, []),
    "vcd_pack_tile_ci": (_i, []),
    "vcd_multi_pack_weights": (_i, [_p, _p, _p, _p, _i, _p]),
    "vcd_conv2d_fprop_ws_bytes": (_i64, [_i] * 8),
    "vcd_conv2d_fprop": (_i, [_p] * 6 + [_i] * 14 + [_p, _i, _p]),
    "vcd_conv2d_dgrad_ws_bytes": (_i64, [_i] * 8),
    "vcd_conv2d_dgrad": (_i, [_p] * 5 + [_i] * 14 + [_p]),
    "vcd_conv2d_dgrad_gn_supported": (_i, [_i] * 8),
    "vcd_conv2d_dgrad_gn": (_i, [_p] * 3 + [_i] * 9 + [_p] * 4 + [_i, _i, _f, _i, _p, _p, _p]),
    "vcd_conv2d_wgrad_ws_bytes": (_i64, [_i] * 8),
    "vcd_conv2d_wgrad": (_i, [_p] * 5 + [_i] + [_p] + [_i] * 14 + [_p]),
    "vcd_conv2d_wgrad_prepare": (_i, [_p, _i, _i, _i, _i, _p]),
    "vcd_pack_upconv_weight": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "vcd_upconv2d_fprop": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "vcd_upconv2d_dgrad": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vcd_upconv2d_wgrad_ws_bytes": (_i64, [_i, _i]),
    "vcd_upconv2d_wgrad": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _p]),
    "vcd_pair_kernel_launches": (_i64, []),
    "vcd_set_pair_kernels": (_i, [_i]),
    "vcd_space_to_planes": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vcd_planes_to_space": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vcd_upsample2x_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vcd_upsample2x_bwd": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vcd_nchw_to_nhwc": (_i, [_p, _i, _p, _i, _i, _i, _i, _p]),
    "vcd_nhwc_to_nchw": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "vcd_add": (_i, [_p, _p, _p, _i64, _p]),
    "vcd_gn_stats": (_i, [_p, _p, _p, _f, _i, _i, _i, _i, _p]),
    "vcd_gn_apply_fwd": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _f, _f, _i, _i, _i, _i, _i, _p]),
    "vcd_gn_bwd_reduce": (_i, [_p, _p, _p, _p, _p, _i, _p, _f, _i, _i, _i, _i, _i, _p]),
    "vcd_gn_bwd_apply": (_i, [_p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _p]),
    "vcd_gn_param_grad": (_i, [_p, _p, _p, _p, _i, _f, _i, _i, _i, _i, _p]),
    "vcd_silu_fwd": (_i, [_p, _p, _i64, _p]),
    "vcd_silu_bwd": (_i, [_p, _p, _p, _i64, _p]),
    "vcd_softmax_fwd": (_i, [_p, _p, _i64, _i, _p]),
    "vcd_softmax_bwd": (_i, [_p, _p, _p, _f, _i64, _i, _p]),
    "vcd_transpose_bf16": (_i, [_p, _p, _i, _i, _i, _p]),
    "vcd_gemm_nt": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _p]),
    "vcd_gemm_tn": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _p]),
    "vcd_gauss_sample_kl_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "vcd_gauss_sample_kl_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "vcd_mse_fwd_bwd": (_i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _p]),
    "vcd_chan_stats": (_i, [_p, _i, _p, _f, _i, _i, _i, _i, _p]),
    "vcd_stats_finalize": (_i, [_p, _p, _p, _i64, _i, _p]),
    "vcd_classify_mask": (_i, [_p, _f, _p, _p, _i, _p]),
    "vcd_nudge_gamma": (_i, [_p, _i, _i, _p, _i, _d, _d, _i, _p, _p]),
    "vcd_dead_weight_count": (_i, [_p, _p, _p, _i, _d, _d, _i, _p, _p, _p]),
    "vcd_optim_chunk_elems": (_i, []),
    "vcd_multi_sqnorm": (_i, [_p, _p, _p, _p, _p, _i, _p, _p]),
    "vcd_clip_adamw_step": (_i, [_p] * 8 + [_i, _p, _d, _d, _d, _d, _d, _d, _p, _i64, _p]),
    "vcd_ssim_psnr_update": (_i, [_p, _p, _i, _i, _i, _i, _f, _i, _f, _p, _p, _p]),
    "vcd_preprocess_u8": (_i, [_p, _p, _i, _i, _i, _i, _p]),
}

_lib = None
launches = 0  # number of C-ABI compute calls issued (bench.py reports it as gpu_launches)
# Measurement hook (bench.py instep_rooflines): when `profile` is a list, every call is bracketed by two CUDA events
# recorded on the call's own stream (the last argument of every entry point) and (name, args, start, end) is appended.
profile = None


class VcdError(RuntimeError):
    pass


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VcdError(
                f"{LIB_PATH} is missing: build it with `python vae-channel-dynamics_b200/build.py` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point, raising VcdError with vcd_last_error() on failure."""
    global launches
    l = lib()
    if profile is not None:
        import torch
        st = torch.cuda.ExternalStream(args[-1]) if args[-1] else torch.cuda.default_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        rc = getattr(l, name)(*args)
        e1.record(st)
        profile.append((name, args, e0, e1))
    else:
        rc = getattr(l, name)(*args)
    launches += 1
    if rc != 0:
        raise VcdError(f"{name} failed ({rc}): {l.vcd_last_error().decode()}")
