"""ctypes binding of the C ABI in include/mllp_b200.h.

There is NO fallback: if libmllp_b200.so is missing or cannot be loaded this module raises,
and every compute entry raises RuntimeError on any non-zero status from the library.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libmllp_b200.so")

NUM_SCALARS = 16
F_DEFAULT, F_NO_SMEM_RESIDENT, F_GRAPH_MODE, F_NO_TUNE, F_PRECONDITION = 0, 1, 2, 4, 8

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64
_dbl = ctypes.c_double

# name -> (restype, argtypes); mirrors include/mllp_b200.h one to one
SIGNATURES = {
    "mllp_last_error": (ctypes.c_char_p, []),
    "mllp_version": (ctypes.c_int, []),
    "mllp_launch_count": (ctypes.c_longlong, []),
    "mllp_format_selfcheck": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "mllp_rowpart_selfcheck": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _i32, _i32, _vp]),
    "mllp_ell_selfcheck": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp]),
    "mllp_format_gather_lines": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "mllp_device_info": (ctypes.c_int, [ctypes.c_int, _vp]),
    "mllp_lp_create": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int,
                                      ctypes.c_uint32, ctypes.POINTER(_vp)]),
    "mllp_graph_edges": (ctypes.c_int, [_i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mllp_norm_scale_work_bytes": (ctypes.c_int64, [_i32]),
    "mllp_norm_scale": (ctypes.c_int, [_i32, _i32, _i64, _i32] + [_vp] * 15),
    "mllp_nccl_unique_id": (ctypes.c_int, [_vp]),
    "mllp_lp_create_rowpart": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int,
                                              ctypes.c_uint32, _i32, _i32, _vp, ctypes.POINTER(_vp)]),
    "mllp_rowpart_ipc_export": (ctypes.c_int, [_vp, _vp]),
    "mllp_rowpart_ipc_import": (ctypes.c_int, [_vp, _vp]),
    "mllp_rowpart_error": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    "mllp_rowpart_mc_supported": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    "mllp_rowpart_mc_create": (ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int32)]),
    "mllp_rowpart_mc_attach": (ctypes.c_int, [_vp, ctypes.c_int32]),
    "mllp_rowpart_mc_bind": (ctypes.c_int, [_vp]),
    "mllp_lp_destroy": (ctypes.c_int, [_vp]),
    "mllp_lp_info": (ctypes.c_int, [_vp, _vp]),
    "mllp_lp_tune_info": (ctypes.c_int, [_vp, _vp]),
    "mllp_lp_geometry": (ctypes.c_int, [_vp, _vp]),
    "mllp_lp_blocks_info": (ctypes.c_int, [_vp, _vp]),
    "mllp_blocks_selfcheck": (ctypes.c_int, [_i32, _i32, _i64, _vp, _vp, _vp, _i32, _vp]),
    "mllp_lp_scaling": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "mllp_spmv": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp]),
    "mllp_estimate_norm": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(_dbl), _vp]),
    "mllp_pdhg_run": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _i32, _vp, _vp]),
    "mllp_pdhg_run_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _i32, _vp, _vp]),
    "mllp_pdhg_solve": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _i32, _i32, _dbl, _vp, _vp]),
    "mllp_batch_create": (ctypes.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int,
                                         ctypes.c_uint32, ctypes.POINTER(_vp)]),
    "mllp_batch_destroy": (ctypes.c_int, [_vp]),
    "mllp_batch_info": (ctypes.c_int, [_vp, _vp]),
    "mllp_batch_estimate_norm": (ctypes.c_int, [_vp, _i32, _vp, _vp]),
    "mllp_batch_run": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "mllp_batch_solve": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i32, _i32, _dbl, _vp, _vp]),
    "mllp_gnn_workspace_floats": (ctypes.c_int64, [_i32, _i32]),
    "mllp_gnn_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mllp_gnn_plan_create": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_vp)]),
    "mllp_gnn_plan_run": (ctypes.c_int, [_vp, _vp]),
    "mllp_gnn_plan_destroy": (ctypes.c_int, [_vp]),
    "mllp_gnn_conv_param_floats": (ctypes.c_int64, [_i32]),
    "mllp_gnn_conv": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "mllp_gnn_flat_param_floats": (ctypes.c_int64, []),
    "mllp_gnn_packed_param_floats": (ctypes.c_int64, []),
    "mllp_gnn_pack_params": (ctypes.c_int, [_vp, _vp, _vp]),
    "mllp_gnn_backward_workspace_floats": (ctypes.c_int64, [_i32, _i32]),
    "mllp_gnn_backward": (ctypes.c_int, [_vp] * 11),
    "mllp_gnn_train_plan_create": (ctypes.c_int, [_vp] * 8 + [ctypes.POINTER(_vp)]),
    "mllp_gnn_backward_plan_create": (ctypes.c_int, [_vp] * 10 + [ctypes.POINTER(_vp)]),
}



class GnnSide(ctypes.Structure):
    """mllp_gnn_side of include/mllp_b200.h"""
    _fields_ = [("nd", _i32), ("ns", _i32), ("group", _i32), ("chunk", _i32), ("indptr", _vp), ("indices", _vp), ("values", _vp),
                ("nlong", _i32), ("nitems", _i32), ("long_rows", _vp), ("long_first", _vp), ("items", _vp), ("scratch", _vp)]


_lib = None


def lib():
    """Load the library (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                "mllp_b200: %s is missing; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)" % SO_PATH)
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    msg = lib().mllp_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what):
    if rc != 0:
        raise RuntimeError("mllp_b200: %s failed (status %d): %s" % (what, rc, last_error()))
