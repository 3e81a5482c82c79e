"""ctypes binding of libb2s.so (C ABI in include/b2s.h). Fails loudly when the CUDA library is absent."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb2s.so")

B2S_FLAG_RELU = 1
B2S_FLAG_STATS = 2

_lib = None

P = c_void_p
I = c_int
F = c_float
LL = c_longlong

# name -> (restype, argtypes); order follows include/b2s.h
SIGNATURES = {
    "b2s_last_error": (c_char_p, []),
    "b2s_launch_count": (LL, []),
    "b2s_version": (I, []),
    "b2s_conv_fwd": (I, [P, I, P, P, P, I, P, I, I, I, I, I, I, I, I, P]),
    "b2s_conv_stats_rows": (I, [I, I, I, I, I]),
    "b2s_conv_fwd_affine": (I, [P, I, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P]),
    "b2s_maxpool2x2": (I, [P, I, P, I, I, I, I, P]),
    "b2s_convt2x2_fwd": (I, [P, I, P, P, P, I, I, I, I, I, I, I, P]),
    "b2s_convt2x2_dgrad": (I, [P, I, P, P, I, I, I, I, I, I, I, P]),
    "b2s_conv_wgrad_workspace": (LL, [I, I, I, I, I, I, I, I, POINTER(c_int)]),
    "b2s_conv3x3_wgrad": (I, [P, I, P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_convt2x2_wgrad": (I, [P, I, P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_wgrad_reduce": (I, [P, I, I, I, I, P, I, P]),
    "b2s_pack_conv_weight": (I, [P, P, P, I, I, I, P]),
    "b2s_pack_convt_weight": (I, [P, P, P, I, I, P]),
    "b2s_pack_weights_all": (I, [I, P, P, P, P, P, P, P, P]),
    "b2s_conv3x3_c1_fwd": (I, [P, P, P, P, P, I, I, I, I, I, P]),
    "b2s_c1_rows": (I, [I, I, I]),
    "b2s_conv3x3_c1_fwd_affine": (I, [P, P, P, P, P, P, I, I, I, I, I, P]),
    "b2s_conv3x3_c1_wgrad": (I, [P, P, P, I, I, I, I, P]),
    "b2s_reduce_rows": (I, [P, I, I, P, P, P]),
    "b2s_bn_finalize": (I, [P, I, I, c_double, P, P, P, P, P, F, F, P, P, P, P, P, P]),
    "b2s_bn_eval_affine": (I, [P, P, P, P, F, P, P, I, P]),
    "b2s_bn_apply": (I, [P, I, P, P, P, I, P, I, I, I, I, P]),
    "b2s_ew_rows": (I, []),
    "b2s_bn_bwd_reduce": (I, [P, I, P, P, I, P, P, P, P, P, I, I, I, I, P]),
    "b2s_bn_bwd_finalize": (I, [P, I, I, c_double, P, P, P, P, P, P, P]),
    "b2s_bn_bwd_apply": (I, [P, I, P, P, I, P, P, P, P, P, P, I, P, I, I, I, I, P]),
    "b2s_head_fwd": (I, [P, I, P, P, P, P, P, P, I, LL, I, I, P]),
    "b2s_head_bwd": (I, [P, P, I, P, P, P, P, I, P, I, LL, I, I, P]),
    "b2s_loss_chunks": (I, [LL]),
    "b2s_seg_loss_fwd": (I, [P, P, I, LL, P, P, P, F, F, F, F, F, F, F, F, P]),
    "b2s_seg_loss_bwd": (I, [P, P, P, P, I, LL, LL, I, P, P, F, F, F, F, F, F, F, F, P]),
    "b2s_metrics_blocks": (I, [LL]),
    "b2s_seg_metrics": (I, [P, P, LL, P, P, P]),
    "b2s_adamw_step": (I, [P, P, P, P, LL, F, F, F, F, F, I, F, P]),
    "b2s_adamw_step_dev": (I, [P, P, P, P, LL, P, P]),
    "b2s_copy_channels": (I, [P, I, P, I, LL, I, P]),
    "b2s_conv3x3_s2_fwd": (I, [P, I, P, P, P, I, I, I, I, I, I, I, P]),
    "b2s_conv3x3_s2_dgrad": (I, [P, I, P, P, I, I, I, I, I, I, I, P]),
    "b2s_conv3x3_s2_wgrad": (I, [P, I, P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_upsample_zero2x": (I, [P, I, P, I, I, I, I, I, P]),
    "b2s_conv1x1_wgrad": (I, [P, I, P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_bn_act_apply": (I, [P, I, P, P, P, I, P, I, LL, I, I, F, ctypes.c_uint, P, P]),
    "b2s_bn_act_bwd_reduce": (I, [P, I, P, I, P, P, P, P, P, LL, I, I, F, ctypes.c_uint, P, P]),
    "b2s_bn_act_bwd_apply": (I, [P, I, P, I, P, P, P, P, P, P, I, P, LL, I, I, F, ctypes.c_uint, P, P]),
    "b2s_relu_bwd": (I, [P, I, P, I, P, I, LL, I, P]),
    "b2s_maxpool2x2_bwd": (I, [P, I, P, I, P, I, I, I, I, I, P]),
    "b2s_channel_sums": (I, [P, I, P, LL, I, P]),
    "b2s_se_chunks": (I, [LL]),
    "b2s_se_pool": (I, [P, I, P, I, P, I, LL, I, P]),
    "b2s_se_fc_fwd": (I, [P, I, LL, P, P, P, P, P, P, P, I, I, I, P]),
    "b2s_se_scale": (I, [P, I, P, P, F, P, I, I, LL, I, P]),
    "b2s_se_fc_bwd": (I, [P, I, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, P]),
    "b2s_conv_dgrad_bnred": (I, [P, I, P, P, I, P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_convt2x2_dgrad_bnred": (I, [P, I, P, P, I, P, I, P, I, I, I, I, I, I, P]),
    "b2s_convt2x2_dgrad_rows": (I, [I, I, I, I, I]),
    "b2s_bn_bwd_finalize_raw": (I, [P, I, I, c_double, P, P, P, P, P, P, P, P]),
    "b2s_psi_rows": (I, [LL]),
    "b2s_psi_stats": (I, [P, I, LL, P, P]),
    "b2s_psi_fwd": (I, [P, I, P, P, P, LL, P]),
    "b2s_psi_bwd_reduce": (I, [P, I, P, P, P, P, LL, P, P]),
    "b2s_psi_bwd_apply": (I, [P, I, P, P, P, P, P, P, LL, P]),
    "b2s_pixel_scale_fwd": (I, [P, I, P, P, I, LL, I, P]),
    "b2s_pixel_scale_bwd": (I, [P, I, P, P, I, P, I, P, LL, I, P]),
    "b2s_bilinear_fwd": (I, [P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_bilinear_bwd": (I, [P, I, P, I, I, I, I, I, I, I, P]),
    "b2s_image_to_nhwc": (I, [P, P, I, I, LL, I, P]),
}


class B2SError(RuntimeError):
    pass


def lib():
    """Load libb2s.so once. Raises B2SError (never falls back) when it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2SError(
            f"{LIB_PATH} not found: build it with csrc/build.sh (or __graft_entry__.build()). "
            "There is no CPU fallback for the B200 hot path.")
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:  # pragma: no cover
            raise B2SError(f"libb2s.so does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().b2s_last_error()
        raise B2SError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().b2s_launch_count())
