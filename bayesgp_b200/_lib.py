"""ctypes binding of libbgp.so (the C ABI declared in include/bgp.h).

The library is built in-tree by ``bayesgp_b200/build.py`` (``__graft_entry__.build()``).
There is no CPU fallback: if the shared object is missing the import fails loudly, and
every entry point returns ``BGP_ERR_CUDA`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BGP_LIB_PATH: diagnostics only (A/B runs of two builds of the same library)
LIB_PATH = os.environ.get("BGP_LIB_PATH") or os.path.join(_HERE, "libbgp.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class BgpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbgp error %d: %s" % (code, msg))
        self.code = code


_lib = None

# name -> (restype, argtypes); must list every symbol include/bgp.h declares
SIGNATURES = {
    "bgp_last_error": (C.c_char_p, []),
    "bgp_version": (C.c_int, []),
    "bgp_kernel_launch_count": (C.c_int64, []),
    "bgp_profiler_range": (C.c_int, [C.c_int]),
    "bgp_model_new": (C.c_int, [C.c_int64, C.c_int, c_double_p, c_double_p, C.c_int, C.POINTER(C.c_void_p)]),
    "bgp_model_add_random": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int, C.c_double, C.c_double,
                                       C.c_double]),
    "bgp_model_add_boundary": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_double, C.c_double]),
    "bgp_model_add_fixed": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_double, C.c_double]),
    "bgp_model_set_noise_prior": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "bgp_model_add_iwp": (C.c_int, [C.c_void_p, c_double_p, C.c_double, c_double_p, C.c_int, C.c_int, C.c_double,
                                    C.c_double, C.c_double, C.c_double]),
    "bgp_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "bgp_model_set_shard": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "bgp_model_finalize": (C.c_int, [C.c_void_p]),
    "bgp_model_destroy": (None, [C.c_void_p]),
    "bgp_model_dims": (C.c_int, [C.c_void_p, c_int64_p, c_int_p, c_int_p]),
    "bgp_objective": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "bgp_laplace_eval": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]),
    "bgp_laplace_eval_batch": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p]),
    "bgp_model_set_start": (C.c_int, [C.c_void_p, c_double_p]),
    "bgp_model_get_tangent": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bgp_model_set_start_at": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "bgp_model_set_newton": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int]),
    "bgp_aghq_fit": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.POINTER(C.c_void_p)]),
    "bgp_aghq_fit_at": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.POINTER(C.c_void_p)]),
    "bgp_fit_destroy": (None, [C.c_void_p]),
    "bgp_fit_dims": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "bgp_fit_get_opt": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p]),
    "bgp_fit_get_grid": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "bgp_fit_get_modes": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bgp_fit_get_marginal": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p]),
    "bgp_sample": (C.c_int, [C.c_void_p, C.c_int64, c_double_p, c_int32_p, c_double_p]),
    "bgp_sample_draw": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, c_double_p, c_int32_p]),
    "bgp_predict_iwp": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, C.c_int, C.c_int, C.c_int,
                                  c_double_p, C.c_int64, C.c_double, C.c_int, c_double_p, c_double_p, c_double_p,
                                  c_double_p]),
    "bgp_predict_sgp": (C.c_int, [c_double_p, c_double_p, c_double_p, C.c_int64, C.c_double, C.c_int, C.c_int,
                                  c_double_p, C.c_int, c_double_p, C.c_int64, C.c_double, C.c_int, c_double_p,
                                  c_double_p, c_double_p, c_double_p]),
    "bgp_basis_iwp": (C.c_int, [c_double_p, C.c_int, C.c_int, c_double_p, C.c_int64, C.c_int, c_double_p]),
    "bgp_model_hessian_flops": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bgp_model_lik_bytes": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "bgp_predict_last_occupancy": (C.c_int, [c_double_p, c_double_p]),
    "bgp_predict_last_timing": (C.c_int, [c_double_p, c_double_p, c_double_p]),
    "bgp_model_gradient_timing": (C.c_int, [C.c_void_p, c_double_p, c_int64_p, c_double_p, c_double_p]),
    "bgp_model_counters": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_int64_p]),
    "bgp_model_set_factor_reuse": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "bgp_model_add_sgp": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_int, c_double_p, c_double_p,
                                    C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "bgp_model_set_node_group": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "bgp_model_set_hessian_retry": (C.c_int, [C.c_void_p, C.c_int]),
    "bgp_model_set_ospline": (C.c_int, [C.c_void_p, C.c_int]),
    "bgp_model_get_ospline": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "bgp_model_ospline_bytes": (C.c_int, [C.c_void_p, c_double_p]),
    "bgp_model_set_lanes": (C.c_int, [C.c_void_p, C.c_int]),
    "bgp_model_get_lanes": (C.c_int, [C.c_void_p, c_int_p]),
    "bgp_fit_get_diagnostics": (C.c_int, [C.c_void_p, c_int_p, c_int64_p, c_double_p, c_double_p]),
    "bgp_fit_host_arrays": (C.c_int, [C.c_void_p, C.POINTER(c_double_p), C.POINTER(c_double_p)]),
    "bgp_fit_node_owner": (C.c_int, [C.c_void_p, c_int32_p]),
    "bgp_fit_predict_iwp": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int, C.c_int,
                                      c_double_p, C.c_int64, C.c_double, c_double_p, c_double_p, c_double_p]),
    "bgp_fit_predict_sgp": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p,
                                      C.c_int, c_double_p, C.c_int64, C.c_double, c_double_p, c_double_p, c_double_p]),
    "bgp_sgp_precision": (C.c_int, [C.c_double, C.c_int, C.c_int, c_double_p, C.c_double, C.c_int, c_double_p, c_double_p]),
    "bgp_model_add_sgp_auto": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_double, C.c_int, C.c_int, c_double_p,
                                         C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "bgp_model_last_timing": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int64_p,
                                        c_int64_p, c_int64_p]),
}


def load():
    """dlopen libbgp.so and attach prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "bayesgp_b200: %s is missing — build it with `python -m bayesgp_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise BgpError(code, load().bgp_last_error().decode("utf-8", "replace"))


def dptr(a):
    """pointer to a C-contiguous/F-contiguous float64 ndarray (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64
    return a.ctypes.data_as(c_double_p)


def fmat(a):
    """float64 column-major (R layout) copy/view of a 2-D array."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def fvec(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
