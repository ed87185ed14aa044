"""ctypes binding of libnbp_b200.so (the C ABI declared in include/nbp_b200.h).

There is no fallback: if the shared library is missing or does not load, importing any operator
raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NBP_B200_LIB") or os.path.join(_HERE, "libnbp_b200.so")   # env override: A/B kernel experiments only

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float
_z = C.c_size_t

# name -> (restype, argtypes); mirrors include/nbp_b200.h one to one
SIGNATURES = {
    "nbp_version": (_i, []),
    "nbp_last_error": (C.c_char_p, []),
    "nbp_launch_count": (C.c_uint64, []),
    "nbp_count_launches": (None, [C.c_uint64]),
    "nbp_raster_workspace_bytes": (_z, [_i, _l]),
    "nbp_raster_depth_batched": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _i, _l, _i, _i, _i, _f, _f, _p, _p, _p, _z, _p]),
    "nbp_backproject_workspace_bytes": (_z, [_i]),
    "nbp_backproject_key_cache_bytes": (_z, [_i, _i, _i]),
    "nbp_backproject_append": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _f, C.c_double, C.c_uint64,
                                    _p, _p, _l, _i, _p, _p, _p, _p, _z, _p]),
    "nbp_grid_scatter": (_i, [_p, _p, _l, _p, _p, _l, _p, _p, _p, _i, _i, _i, _i, _f, _f, _l, _p, _p]),
    "nbp_map_points": (_i, [_p, _p, _i, _l, _i, _i, _f, _f, _p, _p]),
    "nbp_point_cells": (_i, [_p, _l, _i, _i, _f, _f, _p, _p]),
}



class ConvDesc(C.Structure):
    """struct nbp_conv_desc (include/nbp_b200.h)."""
    _fields_ = [("precise", _i), ("src0", _p), ("c0", _i), ("ld0", _i), ("lo0", _i),
                ("src1", _p), ("c1", _i), ("ld1", _i), ("lo1", _i),
                ("n", _i), ("h", _i), ("w", _i), ("taps", _i), ("up2x", _i), ("weight", _p), ("c_out", _i),
                ("scale", _p), ("shift", _p), ("relu", _i), ("dst", _p), ("dst_ld", _i), ("dst_c_off", _i),
                ("dst_lo_off", _i), ("out_f32", _i), ("k_chunk", _i), ("dst_fmt", _i), ("pool_fmt", _i), ("w_lo_scale", _f),
                ("sat_count", _p), ("pool_dst", _p), ("pool_ld", _i), ("pool_lo_off", _i),
                ("dot_w", _p), ("dot_out", _p), ("dot_scale", _f), ("dot_shift", _f), ("dot_sigmoid", _i),
                ("gate_src", _p), ("gate_c", _i), ("gate_ld", _i), ("gate_lo", _i),
                ("dot_n", _i), ("dot_bias", _p), ("dot_max", _p)]


SIGNATURES.update({
    "nbp_conv_fwd": (_i, [C.POINTER(ConvDesc), _p]),
    "nbp_conv_profile_begin": (_i, [_i]),
    "nbp_conv_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "nbp_conv_first": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _i, _i, _p, _i, _i, _i, _p]),
    "nbp_maxpool2x2": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "nbp_upsample2x": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "nbp_att_gate": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _p, _f, _f, _p, _i, _i, _i, _l, _i, _p]),
    "nbp_att_scale": (_i, [_p, _p, _i, _i, _i, _p, _i, _i, _i, _l, _i, _p]),
    "nbp_conv1x1_head": (_i, [_p, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _l, _i, _p]),
})

_d = C.c_double
SIGNATURES.update({
    "nbp_bn_train_stats": (_i, [_p, _i, _i, _l, _i, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p]),
    "nbp_affine_act": (_i, [_p, _i, _i, _l, _i, _p, _p, _i, _p, _i, _i, _p]),
    "nbp_att_pre": (_i, [_p, _p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _i, _i, _p]),
    "nbp_psi_train": (_i, [_p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p]),
    "nbp_att_apply": (_i, [_p, _p, _p, _p, _i, _i, _l, _i, _p, _i, _i, _i, _p, _p]),
    "nbp_bn_bwd": (_i, [_p, _i, _p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p]),
    "nbp_bn_bwd_split": (_i, [_p, _i, _p, _i, _i, _l, _i, _p, _p, _p, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _i, _i, _p, _p, _i, _p]),
    "nbp_to_split_nhwc": (_i, [_p, _i, _l, _i, _p, _p, _i, _i, _p, _i, _p]),
    "nbp_to_split_cnhw": (_i, [_p, _p, _i, _i, _l, _i, _i, _i, _i, _p, _p, _l, _l, _p, _p]),
    "nbp_conv_wgrad": (_i, [_p, _i, _i, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p]),
    "nbp_maxpool2x2_bwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "nbp_upsample2x_bwd": (_i, [_p, _i, _i, _i, _i, _p, _i, _p]),
    "nbp_att_bwd": (_i, [_p, _i, _p, _i, _i, _p, _p, _l, _i, _p, _p, _p, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "nbp_head_bwd": (_i, [_p, _p, _p, _i, _i, _i, _p, _i, _i, _l, _p, _p, _p, _p, _p]),
    "nbp_stem_wgrad": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "nbp_add_f32": (_i, [_p, _p, _i, _l, _i, _p]),
    "nbp_debug_attach_wgrad": (_i, [_p]),
    "nbp_obstacle_fuse": (_i, [_p, _p, _l, _p, _l, _p, _l, _i, _i, _f, _p, _p, _p]),
    "nbp_candidate_scores": (_i, [_p, _p, _i, _p, _p, _p, _i, _i, _p, _i, _i, _f, _f, _i, _p, _p, _p, _p, _p]),
    "nbp_gt_obstacle_map": (_i, [_p, _p, _p, _p, _p, _p, _i, _l, _i, _f, _f, _p, _p]),
    "nbp_segments_hit_mesh": (_i, [_p, _p, _p, _p, _i, _p, _p, _i, _i, _p, _p, _p]),
    "nbp_coverage_percentage": (_i, [_p, _l, _p, _p, _l, _p, _p, _p, _p, _p, _p, _i, _l, _l, _f, _f, _i, C.c_uint64, _p, _p, _p, _p]),
})

_lib = None


def lib():
    """The loaded library; raises RuntimeError with build instructions if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().nbp_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
