"""ctypes binding of include/qoc_b200.h.  Loading never triggers a compute call; every compute call needs a
B200 (the library returns QOC_ERR_NO_DEVICE otherwise -- there is no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

OK, ERR_INVALID, ERR_DIMENSION, ERR_STALE_CACHE, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_NOT_FINITE, \
    ERR_SINGULAR = range(9)
ORDER_FRECHET = 0
COST_INFIDELITY, COST_ABS_TRACE, COST_NONE, COST_ZCAL = 0, 1, 2, 3


class Problem(C.Structure):
    _fields_ = [("d", C.c_int32), ("m", C.c_int32), ("nc", C.c_int32), ("nt", C.c_int32), ("batch", C.c_int32),
                ("order", C.c_int32), ("cost", C.c_int32), ("n", C.c_int32), ("device", C.c_int32),
                ("n_pen_rows", C.c_int32), ("n_pen_cols", C.c_int32),
                ("pen_rows", C.POINTER(C.c_int32)), ("pen_cols", C.POINTER(C.c_int32)), ("mu", C.c_double),
                ("store_costates", C.c_int32), ("reserved", C.c_int32)]


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

# name -> (restype, argtypes); exactly the symbols include/qoc_b200.h declares
SYMBOLS = {
    "qoc_create": (C.c_int, [C.POINTER(Problem), _dp, _dp, _dp, _dp, C.POINTER(_vp)]),
    "qoc_destroy": (C.c_int, [_vp]),
    "qoc_set_order": (C.c_int, [_vp, C.c_int]),
    "qoc_set_cost": (C.c_int, [_vp, C.c_int, _dp, C.c_int]),
    "qoc_set_eager_jacobians": (C.c_int, [_vp, C.c_int]),
    "qoc_set_control_bounds": (C.c_int, [_vp, _dp]),
    "qoc_create_sharded": (C.c_int, [C.POINTER(Problem), _dp, _dp, _dp, _dp, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(_vp)]),
    "qoc_sharded_destroy": (C.c_int, [_vp]),
    "qoc_sharded_set_order": (C.c_int, [_vp, C.c_int]),
    "qoc_sharded_eval": (C.c_int, [_vp, _dp, _dp, _dp]),
    "qoc_sharded_ranks": (C.c_int, [_vp]),
    "qoc_sharded_last_ms": (C.c_double, [_vp]),
    "qoc_sharded_last_error": (C.c_char_p, [_vp]),
    "qoc_propagate": (C.c_int, [_vp, _dp, _dp, _dp]),
    "qoc_gradient": (C.c_int, [_vp, _dp, _dp, _dp]),
    "qoc_eval": (C.c_int, [_vp, _dp, _dp, _dp]),
    "qoc_eval_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "qoc_set_basis": (C.c_int, [_vp, _dp, C.c_int]),
    "qoc_eval_coeffs": (C.c_int, [_vp, _dp, _dp, _dp]),
    "qoc_shard_phase1_device": (C.c_int, [_vp, _vp, _vp, _vp]),
    "qoc_shard_forward_device": (C.c_int, [_vp, _vp, _vp, _vp]),
    "qoc_shard_backward_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "qoc_shard_affine_device": (C.c_int, [_vp, _vp, _vp, _vp]),
    "qoc_shard_phase2_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "qoc_get_states": (C.c_int, [_vp, _dp]),
    "qoc_get_costates": (C.c_int, [_vp, _dp]),
    "qoc_get_propagators": (C.c_int, [_vp, _dp]),
    "qoc_get_jacobians": (C.c_int, [_vp, _dp]),
    "qoc_status_string": (C.c_char_p, [C.c_int]),
    "qoc_last_error": (C.c_char_p, [_vp]),
    "qoc_last_launch_count": (C.c_int, [_vp]),
    "qoc_set_profiling": (C.c_int, [_vp, C.c_int]),
    "qoc_stage_ms": (C.c_double, [_vp, C.c_int]),
    "qoc_last_alg_flops": (C.c_double, [_vp]),
    "qoc_last_exec_flops": (C.c_double, [_vp]),
    "qoc_version": (C.c_int, []),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def _dev_override():
    """Developer aid for A/B builds of the library with other -D switches: QOC_LIB_PATH=<.so> loads that file as is."""
    import os
    return os.environ.get("QOC_LIB_PATH")


def load(build_if_missing: bool = True):
    """Load libqoc_b200.so (building it with nvcc if absent).  Raises if it cannot be had: loudly, no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if _dev_override():
        path = _dev_override()
    elif _build.needs_build():   # missing, or built from other sources than the ones in the tree (content hash)
        if not build_if_missing:
            raise FileNotFoundError(path + " missing or stale: run `python __graft_entry__.py build`")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
