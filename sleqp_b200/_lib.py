"""ctypes loader for libsleqp_b200.so (the C-ABI declared in include/sleqp_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `sleqp_b200/csrc/Makefile`. There is
no fallback: if it is missing, importing the compute classes raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SLEQP_B200_LIB", os.path.join(_HERE, "libsleqp_b200.so"))


class Stats(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("n_elim", C.c_int32),
        ("n_reduced", C.c_int32),
        ("nnz_K", C.c_int64),
        ("nnz_S", C.c_int64),
        ("nnz_L", C.c_int64),
        ("nnz_L_stored", C.c_int64),
        ("n_row_idx", C.c_int64),
        ("n_supernodes", C.c_int32),
        ("n_levels", C.c_int32),
        ("n_stages", C.c_int32),
        ("max_front", C.c_int32),
        ("flops_factor", C.c_double),
        ("flops_factor_stored", C.c_double),
        ("update_ws_doubles", C.c_int64),
        ("pattern_hash", C.c_uint64),
        ("perm_hash", C.c_uint64),
        ("symbolic_cached", C.c_int32),
        ("n_perturbed", C.c_int32),
        ("refine_steps", C.c_int32),
        ("probe_residual", C.c_double),
        ("ms_symbolic", C.c_double),
        ("ms_numeric", C.c_double),
        ("ms_solve", C.c_double),
        ("n_scratch_slots", C.c_int32),
        ("reserved", C.c_int32),
        ("pattern_hash2", C.c_uint64),
        ("flops_update", C.c_double),
        ("flops_inv", C.c_double),
        ("panel_doubles", C.c_int64),
        ("n_demoted", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

# every symbol include/sleqp_b200.h declares (tests check they are all exported)
SYMBOLS = [
    "b200_fact_create", "b200_fact_set_matrix", "b200_fact_set_kkt", "b200_fact_solve", "b200_fact_solve_offset", "b200_fact_solution",
    "b200_fact_solution_ptr", "b200_fact_solution_sparse", "b200_fact_solve_device", "b200_fact_refactor_device", "b200_fact_profile_solve", "b200_fact_profile_numeric", "b200_fact_rcond", "b200_fact_stats",
    "b200_fact_structure", "b200_fact_pivots", "b200_fact_stream", "b200_fact_device", "b200_fact_device_buffers", "b200_fact_free", "b200_last_error",
    "b200_symbolic_analyze", "b200_symbolic_analyze_kkt", "b200_symbolic_stats", "b200_symbolic_structure", "b200_symbolic_export",
    "b200_symbolic_free", "b200_mat_create", "b200_mat_set", "b200_mat_mult_vec", "b200_mat_mult_vec_trans", "b200_mat_mult_vec_trans_sparse",
    "b200_mat_mult_vec_device", "b200_mat_mult_vec_device_if", "b200_mat_mult_vec_trans_device", "b200_mat_stream", "b200_mat_set_stream", "b200_mat_free",
    "b200_cg_create", "b200_cg_set_hess_callback", "b200_cg_solve", "b200_cg_solve_ex", "b200_cg_solve_sparse", "b200_cg_free", "b200_device_count", "b200_launch_count", "b200_host_pin", "b200_host_unpin",
]


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    L = C.CDLL(LIB_PATH)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.b200_last_error.restype = C.c_char_p
    L.b200_fact_create.argtypes = [C.POINTER(vp), C.c_int]
    L.b200_fact_set_matrix.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, dp, C.c_int]
    L.b200_fact_set_kkt.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, dp, ip, ip, C.c_int]
    L.b200_fact_solve.argtypes = [vp, C.c_int, ip, dp, C.c_int]
    L.b200_fact_solve_offset.argtypes = [vp, C.c_int, ip, dp, C.c_int, C.c_int]
    L.b200_fact_solution.argtypes = [vp, C.c_int, C.c_int, dp]
    L.b200_fact_solution_ptr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(dp)]
    L.b200_fact_solution_sparse.argtypes = [vp, C.c_int, C.c_int, C.c_double, ip, dp, ip]
    L.b200_fact_solve_device.argtypes = [vp, vp, vp]
    L.b200_fact_refactor_device.argtypes = [vp, vp]
    L.b200_fact_profile_solve.argtypes = [vp, C.c_int, dp]
    L.b200_fact_profile_numeric.argtypes = [vp, dp]
    L.b200_fact_rcond.argtypes = [vp, dp]
    L.b200_fact_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b200_fact_structure.argtypes = [vp, ip, ip, ip, ip, ip]
    L.b200_fact_pivots.argtypes = [vp, dp]
    L.b200_fact_stream.argtypes = [vp]
    L.b200_fact_stream.restype = vp
    L.b200_fact_device.argtypes = [vp]
    L.b200_fact_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.b200_fact_free.argtypes = [C.POINTER(vp)]
    L.b200_symbolic_analyze.argtypes = [C.POINTER(vp), C.c_int, C.c_int, ip, ip, dp, C.c_int]
    L.b200_symbolic_analyze_kkt.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, ip, ip, dp, ip, ip, C.c_int]
    L.b200_symbolic_stats.argtypes = [vp, C.POINTER(Stats)]
    L.b200_symbolic_structure.argtypes = [vp, ip, ip, ip, ip, ip]
    L.b200_symbolic_export.argtypes = [vp, C.c_char_p, vp, C.POINTER(C.c_int64)]
    L.b200_symbolic_free.argtypes = [C.POINTER(vp)]
    L.b200_mat_create.argtypes = [C.POINTER(vp), C.c_int]
    L.b200_mat_set.argtypes = [vp, C.c_int, C.c_int, C.c_int, ip, ip, dp]
    L.b200_mat_mult_vec.argtypes = [vp, C.c_int, ip, dp, dp]
    L.b200_mat_mult_vec_trans.argtypes = [vp, C.c_int, ip, dp, dp]
    L.b200_mat_mult_vec_trans_sparse.argtypes = [vp, C.c_int, ip, dp, C.c_double, ip, dp, ip]
    L.b200_mat_mult_vec_device.argtypes = [vp, vp, vp]
    L.b200_mat_mult_vec_trans_device.argtypes = [vp, vp, vp]
    L.b200_mat_mult_vec_device_if.argtypes = [vp, vp, vp, vp]
    L.b200_mat_stream.argtypes = [vp]
    L.b200_mat_stream.restype = vp
    L.b200_mat_set_stream.argtypes = [vp, vp]
    L.b200_mat_free.argtypes = [C.POINTER(vp)]
    L.b200_cg_create.argtypes = [C.POINTER(vp), vp, vp]
    L.b200_cg_solve.argtypes = [vp, C.c_int, C.c_int, ip, dp, C.c_double, C.c_double, C.c_int, dp, ip, ip]
    L.b200_cg_solve_ex.argtypes = [vp, C.c_int, C.c_int, ip, dp, C.c_double, C.c_double, C.c_int, dp, ip, ip, dp, dp, dp]
    L.b200_cg_solve_sparse.argtypes = [vp, C.c_int, C.c_int, ip, dp, C.c_double, C.c_double, C.c_int, C.c_double, ip, dp, ip, ip, ip, dp, dp, dp]
    L.b200_cg_set_hess_callback.argtypes = [vp, vp, vp]
    L.b200_cg_free.argtypes = [C.POINTER(vp)]
    L.b200_host_pin.argtypes = [vp, C.c_size_t]
    L.b200_host_unpin.argtypes = [vp]
    L.b200_device_count.restype = C.c_int
    L.b200_launch_count.restype = C.c_int64
    _lib = L
    return L


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[b200 error {code}] {msg}")
        self.code = code


def check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode("utf-8", "replace"))
