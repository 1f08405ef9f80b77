"""ctypes binding of the C ABI declared in include/sp_b200.h.

The enum values are parsed from the header so Python, the CUDA library and the oracle share one
definition.  Loading fails loudly when the CUDA library has not been built: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(ROOT, "include", "sp_b200.h")
LIB_PATH = os.path.join(_HERE, "libsp_b200.so")


def _parse_enums(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    vals = {}
    for m in re.finditer(r"\b(SP_[A-Z0-9_]+)\s*=\s*(\d+)", text):
        vals[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"#define\s+(SP_[A-Z0-9_]+)\s+(\d+)", text):
        vals[m.group(1)] = int(m.group(2))
    return vals


K = _parse_enums(HEADER)
globals().update(K)

KERNEL_IDS = {
    "wendland1": K["SP_KERNEL_WENDLAND1"], "wendland2": K["SP_KERNEL_WENDLAND2"],
    "wendland3": K["SP_KERNEL_WENDLAND3"], "spline23": K["SP_KERNEL_SPLINE23"],
    "spline24": K["SP_KERNEL_SPLINE24"],
}

# every symbol include/sp_b200.h declares: name -> (restype, argtypes)
_p = C.c_void_p
_i32, _i64, _f64 = C.c_int32, C.c_int64, C.c_double
_pi32, _pi64, _pf64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)


class ShapeNode(C.Structure):
    """sp_shape_node of include/sp_b200.h"""
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("p", C.c_double * 8)]


SIGNATURES = {
    "sp_version": (_i32, []),
    "sp_last_error": (C.c_char_p, [_p]),
    "sp_device_count": (_i32, [_pi32]),
    "sp_create": (_i32, [C.POINTER(_p), _pf64, _pf64, _f64, _i32]),
    "sp_destroy": (_i32, [_p]),
    "sp_key_params": (_i32, [_p, _pi64, _pi64, _pi64, _pi32, _pi64]),
    "sp_add_field": (_i32, [_p, C.c_char_p, _i32, _pi32]),
    "sp_find_field": (_i32, [_p, C.c_char_p, _pi32]),
    "sp_resize": (_i32, [_p, _i64]),
    "sp_num_particles": (_i32, [_p, _pi64]),
    "sp_upload": (_i32, [_p, _i32, _pf64, _i64, _i32]),
    "sp_download": (_i32, [_p, _i32, _pf64, _i64, _i32]),
    "sp_synchronize": (_i32, [_p]),
    "sp_create_cell_list": (_i32, [_p]),
    "sp_apply": (_i32, [_p, _i32, _pi32, _i32, _pf64, _i32, _i32]),
    "sp_sum_at_points": (_i32, [_p, _i32, _pi32, _i32, _pf64, _i32, _pf64, _i64, _pf64]),
    "sp_reduce": (_i32, [_p, _i32, _pi32, _i32, _pf64, _i32, _pf64]),
    "sp_poisson_apply": (_i32, [_p, _pi32, _i32, _pf64, _i32]),
    "sp_poisson_cg": (_i32, [_p, _pi32, _i32, _pf64, _i32, _f64, _f64, _i64, _pi64, _pf64]),
    "sp_assemble_matrix": (_i32, [_p, _pi32, _i32, _pf64, _i32, _pi64, _pi64, _pf64, _i64, _pi64]),
    "sp_run_program": (_i32, [_p, _i32, _pi32, _i32, _pf64, _i32, _i64]),
    "sp_graph_begin": (_i32, [_p]),
    "sp_graph_end": (_i32, [_p, _pi32]),
    "sp_graph_launch": (_i32, [_p, _i32, _i64]),
    "sp_graph_destroy": (_i32, [_p, _i32]),
    "sp_kernel_eval": (_i32, [_i32, _i32, _f64, _pf64, _pf64, _i64, _i32]),
    "sp_get_cell_keys": (_i32, [_p, _pi64, _i64]),
    "sp_get_cell_list": (_i32, [_p, _pi64, _pi64]),
    "sp_get_neighbour_lists": (_i32, [_p, _pi64, _pi64, _i64]),
    "sp_get_sweep_neighbour_lists": (_i32, [_p, _pi64, _pi64, _i64]),
    "sp_build_neighbour_lists": (_i32, [_p]),
    "sp_neighbour_list_capacity": (_i32, [_p, _pi32]),
    "sp_generate_particles": (_i32, [_p, _i32, _f64, C.POINTER(ShapeNode), _i32, _pf64, _i32, _pi64, _pi32, _pf64, _i32,
                                     _pi64]),
    "sp_respawn": (_i32, [_p, _i32, _f64, _f64, _f64, _f64, _pi32, _pf64, _i32, _pi64]),
    "sp_num_removed": (_i32, [_p, _pi64]),
    "sp_last_call_ms": (_i32, [_p, C.POINTER(C.c_float)]),
    "sp_timer_start": (_i32, [_p]),
    "sp_timer_stop": (_i32, [_p, C.POINTER(C.c_float)]),
    "sp_launch_count": (_i32, [_p, _pi64]),
    "sp_slab_unique_id": (_i32, [C.POINTER(C.c_uint8)]),
    "sp_slab_init": (_i32, [_p, C.POINTER(C.c_uint8), _i32, _i32, _i32]),
    "sp_slab_init_cuts": (_i32, [_p, C.POINTER(C.c_uint8), _i32, _i32, _i32, _pi64]),
    "sp_slab_range": (_i32, [_p, _pi64, _pi64, _pf64, _pf64, _pi32]),
    "sp_slab_create_cell_list": (_i32, [_p]),
    "sp_slab_halo_refresh": (_i32, [_p, _pi32, _i32]),
    "sp_slab_num_owned": (_i32, [_p, _pi64]),
    "sp_slab_allreduce": (_i32, [_p, _pf64, _i32, _i32]),
}


def declared_symbols():
    """Function names declared in the header (used by the CPU tests to check the exports)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sp_[a-z0-9_]+)\s*\(", text)))


_lib = None


class SpError(RuntimeError):
    """Raised for any non-zero status (the Julia shim rethrows as ErrorException)."""

    def __init__(self, code, text):
        super().__init__(f"sp_b200 status {code}: {text}")
        self.code = code


def load():
    """Load libsp_b200.so (built in-tree by build.py).  No fallback: raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, handle=None):
    if code != 0:
        lib = load()
        msg = lib.sp_last_error(handle)
        raise SpError(code, msg.decode() if msg else "?")


def as_f64(a):
    import numpy as np
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr_f64(a):
    return a.ctypes.data_as(_pf64)


def ptr_i64(a):
    return a.ctypes.data_as(_pi64)


def ptr_i32(a):
    return a.ctypes.data_as(_pi32)
