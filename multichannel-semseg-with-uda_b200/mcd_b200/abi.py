"""ctypes binding of libmcd_sm100.so (include/mcd_sm100.h).

The library is the product: there is no CPU or PyTorch fallback.  `lib()` raises if the shared object
has not been built (`python multichannel-semseg-with-uda_b200/build.py`) and every call raises
`McdError` with the library's own message on a non-zero status.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# MCD_LIB_PATH: an alternative build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("MCD_LIB_PATH") or os.path.join(PKG_DIR, "libmcd_sm100.so")
ABI_VERSION = 21

ALGO_AUTO, ALGO_DIRECT, ALGO_UMMA = 0, 1, 2
OUT_NHWC_BF16, OUT_PLANAR_F32 = 0, 1
FMT_F16, FMT_BF16 = 0, 1      # 16-bit element formats: forward tensors IEEE half, gradients / wgrad twins bfloat16


class McdError(RuntimeError):
    pass


class ConvGeom(Structure):
    """mirror of mcd_conv_geom"""
    _fields_ = [(n, c_int32) for n in ("N", "H", "W", "Cin", "Cout", "Cin_s", "Cout_s", "R", "S",
                                       "stride", "dil", "pad", "Ho", "Wo")]

    def key(self):
        return tuple(getattr(self, n) for n, _ in self._fields_)


P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "mcd_last_error": (c_char_p, []),
    "mcd_version": (c_int, []),
    "mcd_launch_count": (c_int64, []),
    "mcd_check_device": (c_int, [c_int]),
    "mcd_nchw_f32_to_nhwc": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_nhwc_to_nchw_f32": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_convert16": (c_int, [P, c_int, P, c_int64, c_int, P]),
    "mcd_pack_weight": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_pack_weight_rows": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_pack_weight_rowconv": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_sgd_pack_multi": (c_int, [P, c_int, P, c_int, c_int, P]),
    "mcd_pack_weights_multi": (c_int, [P, c_int, c_int, c_int, P]),
    "mcd_conv2d_pack_kind": (c_int, [POINTER(ConvGeom), c_int, c_int]),
    "mcd_conv2d_kernel_id": (c_int, [POINTER(ConvGeom), c_int, c_int, c_int]),
    "mcd_conv2d_fprop": (c_int, [P, P, P, P, c_int, P, P, P, POINTER(ConvGeom), c_int, c_int, P]),
    "mcd_conv2d_streamk_workspace": (c_size_t, [POINTER(ConvGeom), c_int, c_int, c_int, POINTER(c_int)]),
    "mcd_conv2d_dgrad": (c_int, [P, P, P, P, P, P, P, P, P, POINTER(ConvGeom), c_int, c_int, P]),
    "mcd_conv2d_fprop_act_supported": (c_int, [POINTER(ConvGeom)]),
    "mcd_conv2d_fprop_act": (c_int, [P, P, P, P, c_int, P, P, P, POINTER(ConvGeom), c_int, c_int, P]),
    "mcd_conv2d_wgrad_workspace": (c_size_t, [POINTER(ConvGeom), c_int]),
    "mcd_conv2d_wgrad_partials": (c_int, [POINTER(ConvGeom), c_int, POINTER(c_int32)]),
    "mcd_conv2d_wgrad": (c_int, [P, P, P, P, P, c_size_t, POINTER(ConvGeom), c_int, c_int, c_int, P]),
    "mcd_bn_stats": (c_int, [P, P, c_int64, c_int, c_int, c_int, P]),
    "mcd_bn_finalize": (c_int, [P, c_int64, P, P, P, P, c_float, c_float, c_int, P, P, P, P, P, c_int,
                                c_int, P]),
    "mcd_bn_forward": (c_int, [P, P, P, P, P, P, P, c_float, c_float, c_int, P, P, P, P, P, P, P, P, c_float,
                               c_float, c_int, P, c_int, P, P, c_int64, c_int, c_int, c_int, P]),
    "mcd_bn_apply": (c_int, [P, P, P, P, P, P, c_int, P, P, c_int64, c_int, c_int, c_int, P]),
    "mcd_bn_bwd_reduce": (c_int, [P, P, P, P, P, P, P, P, c_int, P, c_int64, c_int, c_int, c_int, P]),
    "mcd_bn_bwd_apply": (c_int, [P, P, P, P, P, P, P, c_int, c_int, P, P, P, P, P, P, P, c_int, P, P,
                                 P, c_int, c_int64, c_int, c_int, c_int, P]),
    "mcd_deconv16s8_fwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_deconv16s8_bwd": (c_int, [P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_head_loss": (c_int, [c_int, c_int, c_int, POINTER(P), POINTER(P), POINTER(P), POINTER(P), P, P, c_int64, P,
                              c_float, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_label_weight_sum": (c_int, [P, P, c_int64, c_int, P, c_int64, c_int, P]),
    "mcd_bilinear_up_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_bilinear_up_bwd": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_ce2d_fwd": (c_int, [P, c_int, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_ce2d_bwd": (c_int, [P, c_int, P, P, c_int64, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_diff2d_fwd": (c_int, [P, P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_diff2d_bwd": (c_int, [P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_mse_fwd": (c_int, [P, c_int, P, P, c_int64, c_int, P]),
    "mcd_mse_bwd": (c_int, [P, c_int, P, P, P, c_int64, c_int, P]),
    "mcd_sum_f32": (c_int, [P, P, c_int64, c_int, P]),
    "mcd_sigmoid3_bce_fwd": (c_int, [P, P, P, c_int, P, P, P, P, c_int64, c_int64, c_int, P]),
    "mcd_sigmoid3_bce_bwd": (c_int, [P, P, P, c_int, P, P, P, P, P, P, c_int64, c_int64, c_int, P]),
    "mcd_bce2d_fwd": (c_int, [P, P, P, P, c_int64, c_int64, c_int, P]),
    "mcd_bce2d_bwd": (c_int, [P, P, P, P, P, c_int64, c_int64, c_int, P]),
    "mcd_label_boundary": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, P]),
    "mcd_pairdist": (c_int, [c_int, P, P, c_int, P, P, P, P, c_float, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_input_transform": (c_int, [P, P, P, P, P, c_int, P, P, c_int, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_relabel_u8": (c_int, [P, P, c_int, c_int, c_int64, c_int, P]),
    "mcd_resize_nearest_u8": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_fast_hist": (c_int, [P, c_int, P, c_int, c_int, c_int64, P, c_int, P]),
    "mcd_unnormalize_u8": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "mcd_argmax_entropy": (c_int, [P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_add_nhwc": (c_int, [P, P, P, P, c_int64, c_int, P]),
    "mcd_gate_fuse_fwd": (c_int, [P, P, P, P, c_int64, c_int, P]),
    "mcd_gate_fuse_bwd": (c_int, [P, P, P, P, P, P, P, c_int64, c_int, P]),
    "mcd_softmax_ch_fwd": (c_int, [P, P, c_int, c_int, c_int64, c_int, P]),
    "mcd_softmax_ch_bwd": (c_int, [P, P, P, c_int, c_int, c_int64, c_int, P]),
    "mcd_cat2_f32": (c_int, [P, c_int, P, c_int, P, c_int, c_int64, c_int, P]),
    "mcd_split2_f32": (c_int, [P, P, c_int, P, c_int, c_int, c_int64, c_int, P]),
    "mcd_sigmoid_fwd": (c_int, [P, P, c_int64, c_int, P]),
    "mcd_sigmoid_bwd": (c_int, [P, P, P, c_int64, c_int, P]),
    "mcd_add3_f32": (c_int, [P, P, P, P, c_int64, c_int, P]),
    "mcd_prob_ce2d_fwd": (c_int, [P, P, P, c_int64, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_prob_ce2d_bwd": (c_int, [P, P, P, c_int64, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_bilinear_ac_up_fwd": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_bilinear_ac_up_bwd": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mcd_sgd_step": (c_int, [P, P, P, c_int64, c_float, c_float, c_float, c_int, c_int, P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load (once) and return the bound library.  Raises when it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McdError(
            "libmcd_sm100.so not found at %s - build it with `python %s` (there is no CPU fallback)"
            % (LIB_PATH, os.path.join(PKG_DIR, "build.py")))
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if handle.mcd_version() != ABI_VERSION:
        raise McdError("libmcd_sm100.so ABI %d != binding ABI %d - rebuild" %
                       (handle.mcd_version(), ABI_VERSION))
    _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        msg = lib().mcd_last_error()
        raise McdError("%s failed (%d): %s" % (what or "libmcd_sm100 call", status,
                                               msg.decode() if msg else "?"))


def launch_count():
    return int(lib().mcd_launch_count())
