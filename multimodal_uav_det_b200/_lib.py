"""ctypes binding of libuavdet_b200.so (the C-ABI in include/uavdet_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails the error is
raised to the caller (north_star: "no CPU fallback")."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libuavdet_b200.so")

OK = 0
ACT = {"none": 0, None: 0, "leaky": 1, "silu": 2, "relu": 3, "gelu": 4}
EPI_AFFINE, EPI_STATS, EPI_HEAD = 0, 1, 2


class Act(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
                ("ld", C.c_int)]


class Epilogue(C.Structure):
    _fields_ = [("epi", C.c_int), ("act", C.c_int), ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("res", C.c_void_p), ("res_ld", C.c_int), ("sum", C.c_void_p), ("sumsq", C.c_void_p),
                ("head_obj", C.c_void_p), ("head_bbox", C.c_void_p), ("head_anchors", C.c_int),
                ("shift_per_sample", C.c_int), ("sample_affine", C.c_void_p)]


class UavdetError(RuntimeError):
    pass


_P = C.c_void_p
_AP = C.POINTER(Act)
_EP = C.POINTER(Epilogue)
_i, _f, _d, _sz, _i64 = C.c_int, C.c_float, C.c_double, C.c_size_t, C.c_int64

# name -> (restype, argtypes).  Must list every symbol declared in include/uavdet_b200.h
SIGNATURES = {
    "uavdet_last_error": (C.c_char_p, []),
    "uavdet_version": (_i, []),
    "uavdet_launch_count": (C.c_uint64, []),
    "uavdet_check_device": (_i, [_P, C.POINTER(_i)]),
    "uavdet_set_sm_margin": (_i, [_i, _i]),
    "uavdet_timestamp": (_i, [_P, _P]),
    "uavdet_nms_workspace_bytes": (_sz, [_i, _i]),
    "uavdet_nms": (_i, [_P, _P, _i, _i, _d, _f, _P, _P, _P, _sz, _P]),
    "uavdet_decode_yolo": (_i, [_P, _P, _i, _i, _i, _i, C.POINTER(_f), _i, _P, _P, _i, _i, _P]),
    "uavdet_encode_targets": (_i, [_P, _P, _i, C.POINTER(_f), _i, _i, C.POINTER(_i), _f, C.POINTER(_P), _P, _P]),
    "uavdet_decode_rtm": (_i, [_P, _i, _i, _i, _i, C.POINTER(_f), _P, _P]),
    "uavdet_cxcywh_to_xyxy": (_i, [_P, _P, _i64, _P]),
    "uavdet_yolo_head_loss_workspace_bytes": (_sz, [_i]),
    "uavdet_yolo_head_loss": (_i, [_P, _P, _P, _i, _i, _i, _i, C.POINTER(_f), _i, _f, _f, _f, _f, _P, _P, _P, _P, _P, _P]),
    "uavdet_conv_fwd": (_i, [_AP, _P, _i, _i, _i, _i, _i, _i, _AP, _EP, _P]),
    "uavdet_conv_dgrad": (_i, [_AP, _P, _i, _i, _i, _i, _i, _AP, _EP, _P]),
    "uavdet_conv3x3_pair_fwd": (_i, [_AP, _P, _i, _AP, _EP, _P]),
    "uavdet_conv_dgrad_s2d": (_i, [_AP, _P, _i, _i, _i, _i, _AP, _EP, _P]),
    "uavdet_pack_dgrad_s2_fused": (_i, [_P, _i, _i, _P, _P]),
    "uavdet_conv_dgrad_s2_fused": (_i, [_AP, _P, _i, _AP, _EP, _P]),
    "uavdet_conv_wgrad": (_i, [_AP, _AP, _i, _i, _i, _i, _P, _i, _P]),
    "uavdet_pack_weight": (_i, [_P, _i, _i, _i, _i, _P, _P]),
    "uavdet_pack_weights_batched": (_i, [_P, _i, C.c_longlong, _P]),
    "uavdet_unpack_wgrad": (_i, [_P, _i, _i, _i, _P, _i, _P]),
    "uavdet_stem_fwd": (_i, [_P, _i, _i, _i, _i, _P, _i, _i, _i, _i, _AP, _EP, _P]),
    "uavdet_stem_s2d_pack": (_i, [_P, _i, _i, _i, _AP, _P]),
    "uavdet_im2col_stem": (_i, [_P, _i, _i, _i, _i, _i, _i, _i, _AP, _P]),
    "uavdet_stem_wgrad": (_i, [_P, _i, _i, _i, _i, _AP, _i, _i, _i, _P, _P]),
    "uavdet_stem_mma_supported": (_i, [_i, _i, _i]),
    "uavdet_stem_mma_fwd": (_i, [_P, _i, _i, _i, _i, _P, _i, _i, _i, _i, _AP, _EP, _P]),
    "uavdet_stem_mma_wgrad": (_i, [_P, _i, _i, _i, _i, _AP, _i, _i, _i, _P, _i, _P]),
    "uavdet_bn_finalize": (_i, [_P, _P, _i, _d, _f, _f, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "uavdet_bn_act_fwd": (_i, [_AP, _P, _P, _i, _AP, _AP, _P]),
    "uavdet_bn_act_bwd_reduce": (_i, [_AP, _AP, _P, _P, _i, _P, _P, _P]),
    "uavdet_bn_bwd_finalize": (_i, [_P, _P, _P, _P, _P, _i, _d, _P, _P, _P, _P, _P]),
    "uavdet_bn_act_bwd_apply": (_i, [_AP, _AP, _P, _P, _P, _P, _i, _AP, _P]),
    "uavdet_bn_train_fwd": (_i, [_AP, _P, _P, _d, _f, _f, _P, _P, _P, _P, _P, _P, _P, _P, _i, _AP, _AP, _P]),
    "uavdet_bn_act_bwd_apply_fused": (_i, [_AP, _AP, _P, _P, _P, _P, _P, _P, _d, _i, _P, _P, _i, _AP, _P]),
    "uavdet_act_bwd": (_i, [_AP, _AP, _P, _P, _i, _AP, _P]),
    "uavdet_upsample2x_fwd": (_i, [_AP, _AP, _P]),
    "uavdet_upsample2x_bwd": (_i, [_AP, _AP, _i, _P]),
    "uavdet_upsample2x_add": (_i, [_AP, _AP, _f, _AP, _P]),
    "uavdet_add": (_i, [_AP, _AP, _AP, _P]),
    "uavdet_nhwc_to_nchw_f32": (_i, [_AP, _P, _P]),
    "uavdet_nchw_f32_to_nhwc": (_i, [_P, _AP, _P]),
    "uavdet_gap": (_i, [_AP, _i, _P, _P]),
    "uavdet_gap_nchw": (_i, [_P, _i, _i, _i, _P, _P]),
    "uavdet_attn_mlp_softmax": (_i, [_P, _i, _i, _P, _P, _i, _P, _P, _i, _f, _P, _P, _P]),
    "uavdet_head_grad_pack": (_i, [_P, _P, _i, _i, _i, _i, _AP, _P, _P, _P]),
    "uavdet_attn_mlp_bwd": (_i, [_P, _P, _P, _P, _i, _i, _P, _i, _P, _i, _f, _f, _P, _P, _P, _P, _P, _P, _P]),
    "uavdet_dyn_aggregate": (_i, [_P, _i, _i, _P, _i, _i, _i, _i, _P, _P, _P, _P]),
    "uavdet_dyn_bwd_contract": (_i, [_P, _i, _i, _P, _P, _i, _i, _i, _i, _P, _P, _P]),
    "uavdet_dyn_bias_bwd": (_i, [_P, _f, _i, _i, _i, _P, _P, _P, _P, _P]),
    "uavdet_dwdynconv_fwd": (_i, [_AP, _P, _P, _i, _i, _AP, _P]),
    "uavdet_dwdynconv_res_stats_fwd": (_i, [_AP, _P, _P, _i, _i, _AP, _P, _AP, _P]),
    "uavdet_linear": (_i, [_P, _i, _i, _P, _P, _i, _i, _P, _P]),
    "uavdet_groupnorm1": (_i, [_AP, _AP, _P, _P, _f, _P, _AP, _P]),
    "uavdet_groupnorm1_stats": (_i, [_AP, _AP, _P, _P]),
    "uavdet_groupnorm1_fold": (_i, [_P, _i, _d, _f, _P, _P]),
    "uavdet_bilinear2x_fwd": (_i, [_AP, _AP, _P]),
    "uavdet_rtm_head_post": (_i, [_P, _P, _i, _i, _i, _i, C.POINTER(_f), _P, _P, _P]),
    "uavdet_sgd_momentum": (_i, [_P, _P, _P, _i64, _f, _f, _f, _i, _P]),
    "uavdet_sgd_momentum_dev": (_i, [_P, _P, _P, _i64, _P, _i, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (raises if it has not been built — no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UavdetError(
            f"{LIB_PATH} not found: build it with `python -m multimodal_uav_det_b200.build` "
            "(or __graft_entry__.build()); this package has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale / missing a symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        msg = load().uavdet_last_error().decode("utf-8", "replace")
        raise UavdetError(f"{what or 'uavdet call'} failed (code {rc}): {msg}")
