"""ctypes binding of libhallucidet_b200.so (the C ABI declared in include/hallucidet_b200.h).

There is no CPU fallback: importing works on a CPU-only host (symbol checks), but every compute
call raises ``RuntimeError`` unless the CUDA library is present and a B200 is the current device.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HD_LIB", os.path.join(HERE, "libhallucidet_b200.so"))   # HD_LIB: A/B builds while tuning

c_int, c_float, c_double, c_void_p, c_int64 = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p, ctypes.c_int64


class HdAct(ctypes.Structure):
    _fields_ = [("ptr", c_void_p), ("n", ctypes.c_int32), ("h", ctypes.c_int32), ("w", ctypes.c_int32), ("c", ctypes.c_int32)]


class HdPackDesc(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("scale", c_void_p), ("w_fwd", c_void_p), ("w_dgrad", c_void_p), ("w_t", c_void_p),
                ("cout", ctypes.c_int32), ("cin", ctypes.c_int32), ("kh", ctypes.c_int32), ("kw", ctypes.c_int32),
                ("cout_pad", ctypes.c_int32), ("k_pad", ctypes.c_int32), ("cin_pad", ctypes.c_int32), ("first_block", ctypes.c_int32),
                ("g", c_void_p), ("m", c_void_p), ("v", c_void_p)]


class HdAdamArgs(ctypes.Structure):
    _fields_ = [("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("bias_correction1", c_float),
                ("bias_correction2", c_float), ("grad_scale", c_float), ("clip", c_float), ("one_minus_beta1", c_float),
                ("one_minus_beta2", c_float)]


class HdAdamDesc(ctypes.Structure):
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", ctypes.c_int32), ("first_block", ctypes.c_int32)]


class HdUnpackDesc(ctypes.Structure):
    _fields_ = [("dw", c_void_p), ("g", c_void_p),
                ("cout", ctypes.c_int32), ("cin", ctypes.c_int32), ("taps", ctypes.c_int32), ("tap_stride", ctypes.c_int32),
                ("row_stride", ctypes.c_int32), ("first_block", ctypes.c_int32), ("scale", ctypes.c_float), ("pad_", ctypes.c_int32)]


class HdRoiGatherArgs(ctypes.Structure):
    _fields_ = [("flat", c_void_p), ("counts", c_void_p), ("props", c_void_p), ("gt", c_void_p), ("labels", c_void_p),
                ("matched", c_void_p), ("batch", ctypes.c_int32), ("slots", ctypes.c_int32), ("n_gt", ctypes.c_int32),
                ("rows", ctypes.c_int32), ("weights", ctypes.c_float * 4), ("canonical_scale", ctypes.c_float),
                ("canonical_level", ctypes.c_float), ("eps", ctypes.c_float), ("k_min", ctypes.c_float), ("k_max", ctypes.c_float),
                ("out_props", c_void_p), ("out_labels", c_void_p), ("out_matched", c_void_p), ("out_image", c_void_p),
                ("out_targets", c_void_p), ("out_rois", c_void_p), ("out_levels", c_void_p), ("out_n_drawn", c_void_p),
                ("out_per_image", c_void_p)]


class HdRoiLevel(ctypes.Structure):
    _fields_ = [("feat_nhwc", c_void_p), ("grad_nhwc", c_void_p), ("h", ctypes.c_int32), ("w", ctypes.c_int32),
                ("scale", ctypes.c_float), ("pad_", ctypes.c_int32)]


P_ = ctypes.POINTER


class HdBnFin(ctypes.Structure):
    _fields_ = [("count", c_double), ("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p), ("running_var", c_void_p),
                ("mean", c_void_p), ("invstd", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("counter", c_void_p),
                ("eps", c_float), ("momentum", c_float)]


class HdConvArgs(ctypes.Structure):
    _fields_ = [
        ("x0", HdAct), ("x1", HdAct), ("y0", HdAct), ("y1", HdAct),
        ("w", c_void_p),
        ("kh", ctypes.c_int32), ("kw", ctypes.c_int32), ("stride", ctypes.c_int32),
        ("bias", c_void_p), ("add", c_void_p), ("mask", c_void_p),
        ("relu", ctypes.c_int32), ("sigmoid", ctypes.c_int32),
        ("stats", c_void_p), ("stats_replicas", ctypes.c_int32),
        ("out_f32_nchw", c_void_p), ("out_f32_channels", ctypes.c_int32),
        ("store_bf16", ctypes.c_int32), ("phase_mask", ctypes.c_int32),
        ("dw", c_void_p), ("split_k", ctypes.c_int32), ("out_f32_nhwc", ctypes.c_int32),
        ("bn_fin", P_(HdBnFin)),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
    ]


P = ctypes.POINTER
# name -> argtypes (restype is int unless listed in _RESTYPES)
PROTOTYPES = {
    "hd_version": [],
    "hd_last_error": [],
    "hd_device_ok": [],
    "hd_conv_fwd": [P(HdConvArgs), c_void_p],
    "hd_conv_fwd_tiles": [P(HdConvArgs)],
    "hd_conv_dgrad": [P(HdConvArgs), c_void_p],
    "hd_conv_workspace_bytes": [],
    "hd_conv_has_streamk": [],
    "hd_conv_debug_timestamps": [c_void_p],
    "hd_conv_wgrad": [P(HdConvArgs), c_void_p],
    "hd_pack_conv_weight": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "hd_unpack_wgrad": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p],
    "hd_stem_im2col": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "hd_stem_col2im": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "hd_stem_fwd": [c_void_p, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                    P_(HdBnFin), c_void_p],
    "hd_stem_fwd_rows": [c_int, c_int, c_int],
    "hd_stem_im2col_1ch": [c_void_p, c_int, c_float, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "hd_bn_finalize": [c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "hd_bn_apply": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p],
    "hd_bn_bwd_reduce": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "hd_bn_bwd_apply": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                        c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p],
    "hd_bn_bwd_fused": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double,
                        c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p],
    "hd_maxpool_fwd": [P(HdAct), P(HdAct), c_void_p, c_int, c_void_p],
    "hd_maxpool_bwd": [P(HdAct), P(HdAct), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "hd_upsample2x_fwd": [P(HdAct), P(HdAct), c_void_p],
    "hd_upsample2x_bwd": [P(HdAct), P(HdAct), c_void_p],
    "hd_add_nearest_fwd": [P(HdAct), P(HdAct), c_void_p],
    "hd_add_nearest_bwd": [P(HdAct), P(HdAct), c_int, c_void_p],
    "hd_pad_hw": [P(HdAct), P(HdAct), c_void_p],
    "hd_crop_add_mask": [P(HdAct), c_void_p, c_void_p, P(HdAct), c_void_p],
    "hd_nchw_f32_to_nhwc_bf16": [c_void_p, P(HdAct), c_int, c_int, c_void_p],
    "hd_nhwc_bf16_to_nchw_f32": [P(HdAct), c_void_p, c_int, c_void_p],
    "hd_sigmoid_bwd_pack": [c_void_p, c_void_p, P(HdAct), c_int, c_void_p, c_void_p],
    "hd_resize_nearest_fwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "hd_resize_nearest_bwd": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p],
    "hd_regulariser": [c_int, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int, c_int, c_void_p, c_void_p,
                       c_float, c_int, c_void_p],
    "hd_pack_blocks": [c_void_p],
    "hd_multi_blocks": [ctypes.c_int64],
    "hd_pack_conv_weights": [c_void_p, c_int, c_int, c_void_p],
    "hd_unpack_wgrads": [c_void_p, c_int, c_int, c_void_p],
    "hd_adam_pack_conv_weights": [c_void_p, c_int, c_int, P(HdAdamArgs), c_void_p],
    "hd_adam_multi": [c_void_p, c_int, c_int, P(HdAdamArgs), c_void_p],
    "hd_roi_align_bwd_nhwc": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "hd_roi_align_fwd_nhwc": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p],
    "hd_roi_align_ml_fwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hd_roi_align_ml_fwd_bf16": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hd_roi_align_ml_bwd": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p],
    "hd_nchw_to_nhwc_f32": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "hd_nhwc_to_nchw_f32": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p],
    "hd_nms": [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p],
    "hd_nms_valid": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p],
    "hd_roi_match_labels": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p,
                            c_void_p],
    "hd_roi_gather_samples": [P(HdRoiGatherArgs), c_void_p],
    "hd_fastrcnn_loss": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "hd_rpn_loss": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p,
                    c_void_p, c_void_p],
    "hd_rpn_concat_preds": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p],
    "hd_unpack_blocks": [c_int, c_int, c_int],
    "hd_rpn_assign_targets": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p],
    "hd_rpn_decode_selected": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_float, c_float, c_float,
                               c_float, c_void_p, c_void_p, c_void_p, c_void_p],
    "hd_sample_balanced_workspace_bytes": [c_int],
    "hd_sample_balanced": [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_void_p, ctypes.c_int64, c_void_p],
}
_RESTYPES = {"hd_last_error": ctypes.c_char_p, "hd_conv_workspace_bytes": ctypes.c_int64,
             "hd_sample_balanced_workspace_bytes": ctypes.c_int64}

_lib = None


def load():
    """Load the shared library (raises if it has not been built; never falls back to another implementation)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m hallucidet_b200.build` "
                "(hallucidet_b200 has no CPU / PyTorch fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(lib, name)           # AttributeError = ABI mismatch, fail loudly
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, c_int)
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = load().hd_last_error()
        raise RuntimeError(f"{what} failed with hd_status {status}: {msg.decode() if msg else ''}")
