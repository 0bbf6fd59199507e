"""ctypes binding of the C ABI in include/cvb200.h (libcvb200.so).  Product path:
fails loudly when the CUDA library is missing -- there is no CPU / PyTorch fallback."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_C", "libcvb200.so")

_f = ctypes.c_void_p   # device float*
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_i32 = ctypes.c_int32
_F3 = ctypes.c_float * 3
_I3 = ctypes.c_int32 * 3

# name -> (restype, argtypes); must list every symbol include/cvb200.h declares
SIGNATURES = {
    "cvb200_abi_version": (ctypes.c_int, []),
    "cvb200_last_error": (ctypes.c_char_p, []),
    "cvb200_hv_grid_dims_work_bytes": (ctypes.c_size_t, []),
    "cvb200_hv_grid_dims": (ctypes.c_int, [_f, _i64, ctypes.c_float, _vp, ctypes.POINTER(ctypes.c_float),
                                            ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i32), _vp]),
    "cvb200_hv_forward_work_bytes": (ctypes.c_size_t, [ctypes.POINTER(_i32)]),
    "cvb200_hv_forward": (ctypes.c_int, [_f, _f, _f, _f, _i64, ctypes.c_float, _i32, ctypes.POINTER(ctypes.c_float),
                                          ctypes.POINTER(_i32), _f, _f, _f, _vp, ctypes.c_size_t, _vp]),
    "cvb200_hv_backward": (ctypes.c_int, [_f, _f, _f, _f, _f, _i64, ctypes.c_float, _i32,
                                           ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_i32), _f, _f, _f, _vp]),
    "cvb200_hv_vote_indices": (ctypes.c_int, [_f, _f, _f, _i64, ctypes.c_float, _i32, ctypes.POINTER(ctypes.c_float),
                                               ctypes.POINTER(_i32), _vp, _vp]),
    "cvb200_hv_theta_table": (ctypes.c_int, [_i32, _f, _f, _vp]),
    "cvb200_hv_project_y": (ctypes.c_int, [_f, ctypes.POINTER(_i32), _f, _vp, _vp]),
    "cvb200_hv_proposals_work_bytes": (ctypes.c_size_t, [_i32]),
    "cvb200_hv_proposals": (ctypes.c_int, [_vp, _i32, _vp, _f, ctypes.POINTER(_i32), ctypes.c_float, ctypes.POINTER(ctypes.c_float),
                                            _f, _i32, ctypes.c_float, _i32, _f, _f, _vp, _vp, ctypes.c_size_t, _vp]),
}

ABI_VERSION = 14


class BpParams(ctypes.Structure):
    """cvb200_bp_params (include/cvb200.h)."""
    _fields_ = [("thresh_high", ctypes.c_float), ("thresh_low", _i32), ("valid_ratio", ctypes.c_float),
                ("elimination", _i32), ("elim_hi_inclusive", _i32), ("prob_thresh", ctypes.c_float),
                ("err_thresh", ctypes.c_float), ("max_boxes", _i32), ("max_iters", _i32), ("max_trace", _i32)]


SIGNATURES.update({
    "cvb200_bp_default_params": (None, [ctypes.POINTER(BpParams)]),
    "cvb200_bp_work_bytes": (ctypes.c_size_t, [ctypes.POINTER(_i32)]),
    "cvb200_back_project": (ctypes.c_int, [_f, _f, _f, ctypes.POINTER(_i32), ctypes.POINTER(ctypes.c_float), ctypes.c_float,
                                            _f, _f, _f, _vp, _i64, ctypes.POINTER(BpParams), _f, _f, _vp, _vp, _vp, _vp,
                                            ctypes.c_size_t, _vp]),
})

SIGNATURES.update({
    "cvb200_obb_nms": (ctypes.c_int, [_f, _f, _vp, _i32, _i32, ctypes.c_double, _vp, _vp, _vp]),
    "cvb200_obb_iou_matrix": (ctypes.c_int, [_f, _i32, _f, _i32, _vp, _vp]),
})

SIGNATURES.update({
    "cvb200_sc_hash_capacity": (_i64, [_i64]),
    "cvb200_sc_build_table": (ctypes.c_int, [_vp, _i64, _vp, _vp, _i64, _vp]),
    "cvb200_sc_down_flags": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp]),
    "cvb200_sc_down_finish": (ctypes.c_int, [_vp, _i64, _i32, _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cvb200_sc_kernel_map": (ctypes.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "cvb200_sc_conv_forward": (ctypes.c_int, [_f, _i32, _f, _i32, _vp, _i64, _i32, _f, _f, _vp]),
    "cvb200_sc_conv_forward_tc": (ctypes.c_int, [_f, _i64, _i32, _f, _i32, _vp, _i64, _i32, _f, _f, _vp]),
    "cvb200_sc_set_conv_options": (ctypes.c_int, [_i32, _i32]),
    "cvb200_sc_set_conv_debug": (ctypes.c_int, [_i32]),
    "cvb200_sc_set_conv_trace": (ctypes.c_int, [_vp]),
    "cvb200_sc_conv_plan": (ctypes.c_int, [_i64, _i32, _i32, _i32, ctypes.POINTER(_i32), ctypes.POINTER(_i32), _i32]),
    "cvb200_sc_conv_wgrad": (ctypes.c_int, [_f, _i32, _f, _i32, _vp, _i64, _i32, _i32, _f, _vp]),
    "cvb200_sc_conv_wgrad_tc": (ctypes.c_int, [_f, _i32, _f, _i32, _vp, _i64, _i32, _f, _vp]),
})



class DecodeArgs(ctypes.Structure):
    """cvb200_decode_args (include/cvb200.h)."""
    _fields_ = [("xyz", _vp), ("scale", _vp), ("class_pred", _vp), ("prob", _vp), ("coords", _vp), ("points", _vp),
                ("res", ctypes.c_float), ("nclasses", _i32), ("log_scale", _i32)]


class ScOp(ctypes.Structure):
    """cvb200_sc_op (include/cvb200.h)."""
    _fields_ = [("kind", _i32), ("cin", _i32), ("cout", _i32), ("k3", _i32), ("ldi", _i32), ("ldo", _i32), ("ldr", _i32),
                ("relu", _i32), ("n_out", _i64), ("n_in", _i64), ("in_", _vp), ("w", _vp), ("bias", _vp), ("residual", _vp), ("table", _vp),
                ("out", _vp), ("n_out_dev", _vp), ("decode", ctypes.POINTER(DecodeArgs))]


class MapsLayout(ctypes.Structure):
    """cvb200_sc_maps_layout_t (include/cvb200.h)."""
    _fields_ = [("total_bytes", _i64), ("capacity", _i64), ("counts", _i64), ("arange", _i64), ("stem_table", _i64),
                ("coords", _i64 * 5), ("keys", _i64 * 5), ("vals", _i64 * 5), ("nbr3", _i64 * 5),
                ("children", _i64 * 4), ("up_table", _i64 * 4), ("parent", _i64 * 4), ("koff", _i64 * 4),
                ("flag", _i64), ("scan", _i64), ("cub_temp", _i64), ("cub_temp_bytes", _i64), ("fill_ff", _i64), ("fill_ff_bytes", _i64)]


SIGNATURES.update({
    "cvb200_sc_quantize": (ctypes.c_int, [_f, _i64, ctypes.c_float, _i32, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "cvb200_sc_maps_layout": (ctypes.c_int, [_i64, _i32, _i32, ctypes.POINTER(MapsLayout)]),
    "cvb200_sc_build_maps": (ctypes.c_int, [_vp, _i64, _i32, _i32, _vp, ctypes.POINTER(MapsLayout), _vp, _vp]),
})

SIGNATURES.update({
    "cvb200_sc_run_program": (ctypes.c_int, [ctypes.POINTER(ScOp), _i32, _vp]),
    "cvb200_head_decode": (ctypes.c_int, [_f, _i32, _i64, _i32, _i32, _f, _f, _vp, _f, _vp]),
    "cvb200_head_decode_points": (ctypes.c_int, [_f, _i32, _i64, _i32, _i32, _f, _f, _vp, _f, _vp, ctypes.c_float, _f, _vp]),
})

SIGNATURES.update({
    "cvb200_graph_begin": (ctypes.c_int, [_vp]),
    "cvb200_graph_end": (ctypes.c_int, [_vp, _i32, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(_i64)]),
    "cvb200_graph_abort": (ctypes.c_int, [_vp]),
    "cvb200_graph_launch": (ctypes.c_int, [_vp, _vp]),
    "cvb200_graph_destroy": (ctypes.c_int, [_vp]),
})

_lib = None


class CVB200Error(RuntimeError):
    pass


def load():
    """Load libcvb200.so (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CVB200Error(
            "canonicalvoting_b200: %s is missing -- build it with `python -m canonicalvoting_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.cvb200_abi_version() != ABI_VERSION:
        raise CVB200Error("libcvb200.so ABI version %d != %d (stale build?)" % (L.cvb200_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = load().cvb200_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg))


def f3(v):
    return _F3(float(v[0]), float(v[1]), float(v[2]))


def i3(v):
    return _I3(int(v[0]), int(v[1]), int(v[2]))
