"""ctypes binding of libfungp_b200.so (the C ABI declared in include/fungp_b200.h).

There is no CPU fallback: if the shared library has not been built (``python -c "import __graft_entry__ as g;
g.build()"`` or ``make -C fungp_b200/csrc``) or no sm_100 GPU is visible, loading / context creation raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfungp_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_dpp = C.POINTER(c_dp)

# name -> (restype, argtypes); every symbol declared in include/fungp_b200.h
SIGNATURES = {
    "fgp_ctx_create": (C.c_int, [C.c_int, c_ip, C.POINTER(C.c_void_p)]),
    "fgp_ctx_destroy": (None, [C.c_void_p]),
    "fgp_ctx_device_count": (C.c_int, [C.c_void_p]),
    "fgp_last_error": (C.c_char_p, [C.c_void_p]),
    "fgp_version": (C.c_char_p, []),
    "fgp_bspline_basis": (C.c_int, [C.c_int, C.c_int, c_dp]),
    "fgp_project": (C.c_int, [C.c_void_p, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    "fgp_test_dag_order": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_ip, c_ip, c_ip]),
    "fgp_test_project_host": (C.c_int, [c_dp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    "fgp_problem_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_ip, c_dpp, c_dpp, c_dp,
                                     C.POINTER(C.c_void_p)]),
    "fgp_problem_destroy": (None, [C.c_void_p]),
    "fgp_problem_n": (C.c_int, [C.c_void_p]),
    "fgp_problem_nls": (C.c_int, [C.c_void_p]),
    "fgp_problem_set_y": (C.c_int, [C.c_void_p, c_dp]),
    "fgp_dist_max": (C.c_int, [C.c_void_p, c_dp]),
    "fgp_dist_materialize": (C.c_int, [C.c_void_p, C.c_int, c_dp]),
    "fgp_assemble": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp, c_dp]),
    "fgp_nll_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "fgp_premats": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp]),
    "fgp_problem_restore": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp]),
    "fgp_update_var": (C.c_int, [C.c_void_p, C.c_double, c_dp, c_dp]),
    "fgp_update_y": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "fgp_premats_reuse": (C.c_int, [C.c_void_p, C.c_void_p, c_dp, c_dp, c_ip]),
    "fgp_predict": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dpp, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    "fgp_simulate": (C.c_int, [C.c_void_p, C.c_int, c_dp, c_dpp, C.c_int, c_dp, C.c_double, c_dp, c_dp, c_dp]),
    "fgp_loocv": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "fgp_fit_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_dpp, c_dp, C.c_int, c_ip,
                                C.c_double, C.c_int, C.c_int, c_dp, C.POINTER(C.c_longlong), C.c_int, c_dp, c_dp,
                                c_dp, c_dp, c_ip]),
    "fgp_fit_batch_holdout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_ip, c_dpp, c_dp, C.c_int, c_ip,
                                        C.c_double, C.c_int, C.c_int, c_dp, C.POINTER(C.c_longlong), C.c_int, C.c_int, C.c_int,
                                        c_ip, c_dp, c_dp, c_dp, c_dp, c_ip]),
    "fgp_get_timing": (C.c_int, [C.c_void_p, c_dp]),
    "fgp_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "fgp_launch_count": (C.c_longlong, []),
    "fgp_work_count": (C.c_longlong, [C.c_int]),
    "fgp_last_factor_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "fgp_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "fgp_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp]),
    "fgp_measure_peaks": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp]),
}

_lib = None


def load():
    """Load the shared library and attach the signatures.  Raises if it is missing -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C fungp_b200/csrc` (nvcc, sm_100a). "
            "fungp_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
