"""ctypes driver for the UNMODIFIED reference built by oracle/build_ref.sh -- TEST INFRASTRUCTURE
ONLY (see oracle/sleqp_oracle.py for the import rule).

Wraps exactly the reference entry points on the hot path:
  SleqpMat / SleqpVec        sparse/pub_mat.h:35-151, sparse/pub_vec.h:16-146
  sleqp_mat_mult_vec{,_trans} sparse/mat.c:282-363
  SleqpFact                  fact/fact.h:23-68 (fact_create_default -> the backend linked in)

`RefLib("lapack")` loads oracle/_ref/libsleqp_ref_lapack.so (reference LAPACK backend, the dense
oracle). `RefLib("b200")` loads oracle/_ref/libsleqp_ref_b200.so: the same reference objects with
sleqp_b200/host/fact_b200.c linked in place of a reference backend -- the drop-in test.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class SleqpVec(C.Structure):  # sparse/pub_vec.h:16-25
    _fields_ = [("data", C.POINTER(C.c_double)), ("indices", C.POINTER(C.c_int)), ("dim", C.c_int), ("nnz", C.c_int), ("nnz_max", C.c_int)]


def available(kind="lapack"):
    return os.path.exists(os.path.join(HERE, "_ref", f"libsleqp_ref_{kind}.so"))


class RefLib:
    def __init__(self, kind="lapack"):
        path = os.path.join(HERE, "_ref", f"libsleqp_ref_{kind}.so")
        self.L = L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        L.sleqp_mat_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int]
        L.sleqp_mat_push.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.sleqp_mat_push_col.argtypes = [vp, C.c_int]
        L.sleqp_mat_release.argtypes = [C.POINTER(vp)]
        L.sleqp_mat_mult_vec.argtypes = [vp, C.POINTER(SleqpVec), C.POINTER(C.c_double)]
        L.sleqp_mat_mult_vec_trans.argtypes = [vp, C.POINTER(SleqpVec), C.c_double, C.POINTER(SleqpVec)]
        L.sleqp_mat_data.argtypes = [vp]
        L.sleqp_mat_data.restype = C.POINTER(C.c_double)
        L.sleqp_mat_cols.argtypes = [vp]
        L.sleqp_mat_cols.restype = C.POINTER(C.c_int)
        L.sleqp_mat_rows.argtypes = [vp]
        L.sleqp_mat_rows.restype = C.POINTER(C.c_int)
        L.sleqp_mat_set_nnz.argtypes = [vp, C.c_int]
        L.sleqp_vec_create.argtypes = [C.POINTER(C.POINTER(SleqpVec)), C.c_int, C.c_int]
        L.sleqp_vec_create_empty.argtypes = [C.POINTER(C.POINTER(SleqpVec)), C.c_int]
        L.sleqp_vec_push.argtypes = [C.POINTER(SleqpVec), C.c_int, C.c_double]
        L.sleqp_vec_free.argtypes = [C.POINTER(C.POINTER(SleqpVec))]
        L.sleqp_vec_set_from_raw.argtypes = [C.POINTER(SleqpVec), C.POINTER(C.c_double), C.c_int, C.c_double]
        L.sleqp_settings_create.argtypes = [C.POINTER(vp)]
        L.sleqp_settings_release.argtypes = [C.POINTER(vp)]
        L.sleqp_fact_create_default.argtypes = [C.POINTER(vp), vp]
        L.sleqp_fact_set_matrix.argtypes = [vp, vp]
        L.sleqp_fact_solve.argtypes = [vp, C.POINTER(SleqpVec)]
        L.sleqp_fact_solution.argtypes = [vp, C.POINTER(SleqpVec), C.c_int, C.c_int, C.c_double]
        L.sleqp_fact_cond.argtypes = [vp, C.POINTER(C.c_double)]
        L.sleqp_fact_flags.argtypes = [vp]
        L.sleqp_fact_name.argtypes = [vp]
        L.sleqp_fact_name.restype = C.c_char_p
        L.sleqp_fact_release.argtypes = [C.POINTER(vp)]
        L.sleqp_error_msg.restype = C.c_char_p

    def call(self, rc):
        if rc != 0:
            raise RuntimeError("reference call failed: " + self.L.sleqp_error_msg().decode("utf-8", "replace"))

    # ---- containers -----------------------------------------------------------------------
    def mat(self, num_rows, num_cols, colptr, rows, data):
        """SleqpMat filled through the public push API (sleqp_mat_push_col / sleqp_mat_push)."""
        L = self.L
        m = C.c_void_p()
        nnz = int(len(rows))
        self.call(L.sleqp_mat_create(C.byref(m), int(num_rows), int(num_cols), max(nnz, 1)))
        if nnz > 64:
            # bulk fill through the accessors (pub_mat.h:102-114): push every column start, then
            # write rows/data/cols in place and set nnz -- same final layout as repeated pushes
            for j in range(1):
                pass
            cols_p, rows_p, data_p = L.sleqp_mat_cols(m), L.sleqp_mat_rows(m), L.sleqp_mat_data(m)
            C.memmove(rows_p, np.ascontiguousarray(rows, dtype=np.int32).ctypes.data, 4 * nnz)
            C.memmove(data_p, np.ascontiguousarray(data, dtype=np.float64).ctypes.data, 8 * nnz)
            C.memmove(cols_p, np.ascontiguousarray(colptr, dtype=np.int32).ctypes.data, 4 * (num_cols + 1))
            self.call(L.sleqp_mat_set_nnz(m, nnz))
        else:
            for j in range(num_cols):
                self.call(L.sleqp_mat_push_col(m, j))
                for q in range(int(colptr[j]), int(colptr[j + 1])):
                    self.call(L.sleqp_mat_push(m, int(rows[q]), j, float(data[q])))
        return m

    def vec(self, dim, idx, val):
        v = C.POINTER(SleqpVec)()
        self.call(self.L.sleqp_vec_create(C.byref(v), int(dim), max(int(len(idx)), 1)))
        n = int(len(idx))
        if n:
            C.memmove(v.contents.indices, np.ascontiguousarray(idx, dtype=np.int32).ctypes.data, 4 * n)
            C.memmove(v.contents.data, np.ascontiguousarray(val, dtype=np.float64).ctypes.data, 8 * n)
            v.contents.nnz = n
        return v

    @staticmethod
    def vec_to_numpy(v):
        n = v.contents.nnz
        idx = np.ctypeslib.as_array(v.contents.indices, shape=(max(n, 1),))[:n].copy()
        val = np.ctypeslib.as_array(v.contents.data, shape=(max(n, 1),))[:n].copy()
        return idx.astype(np.int32), val, int(v.contents.dim)

    # ---- SpMV --------------------------------------------------------------------------------
    def mat_mult_vec(self, num_rows, num_cols, colptr, rows, data, x_idx, x_val):
        m = self.mat(num_rows, num_cols, colptr, rows, data)
        v = self.vec(num_cols, x_idx, x_val)
        out = np.empty(num_rows, dtype=np.float64)
        self.call(self.L.sleqp_mat_mult_vec(m, v, out.ctypes.data_as(C.POINTER(C.c_double))))
        self.L.sleqp_vec_free(C.byref(v))
        self.L.sleqp_mat_release(C.byref(m))
        return out

    def mat_mult_vec_trans(self, num_rows, num_cols, colptr, rows, data, v_idx, v_val, eps):
        m = self.mat(num_rows, num_cols, colptr, rows, data)
        v = self.vec(num_rows, v_idx, v_val)
        r = C.POINTER(SleqpVec)()
        self.call(self.L.sleqp_vec_create_empty(C.byref(r), int(num_cols)))
        self.call(self.L.sleqp_mat_mult_vec_trans(m, v, float(eps), r))
        idx, val, _ = self.vec_to_numpy(r)
        for p in (v, r):
            self.L.sleqp_vec_free(C.byref(p))
        self.L.sleqp_mat_release(C.byref(m))
        return idx, val

    def vec_set_from_raw(self, values, zero_eps):
        values = np.ascontiguousarray(values, dtype=np.float64)
        r = C.POINTER(SleqpVec)()
        self.call(self.L.sleqp_vec_create_empty(C.byref(r), int(len(values))))
        self.call(self.L.sleqp_vec_set_from_raw(r, values.ctypes.data_as(C.POINTER(C.c_double)), int(len(values)), float(zero_eps)))
        idx, val, _ = self.vec_to_numpy(r)
        self.L.sleqp_vec_free(C.byref(r))
        return idx, val

    # ---- factorization plugin --------------------------------------------------------------
    def fact(self):
        return RefFact(self)


class RefFact:
    """sleqp_fact_create_default -> set_matrix -> solve -> solution, as the aug_jac drives it."""

    def __init__(self, lib: RefLib):
        self.lib = lib
        L = lib.L
        self.settings = C.c_void_p()
        lib.call(L.sleqp_settings_create(C.byref(self.settings)))
        self.h = C.c_void_p()
        lib.call(L.sleqp_fact_create_default(C.byref(self.h), self.settings))
        self.mat = None
        self.N = 0

    def name(self):
        return self.lib.L.sleqp_fact_name(self.h).decode()

    def flags(self):
        return int(self.lib.L.sleqp_fact_flags(self.h))

    def set_matrix(self, N, colptr, rows, data):
        L = self.lib.L
        if self.mat is not None:
            L.sleqp_mat_release(C.byref(self.mat))
        self.mat = self.lib.mat(N, N, colptr, rows, data)
        self.N = N
        self.lib.call(L.sleqp_fact_set_matrix(self.h, self.mat))

    def solve(self, idx, val):
        v = self.lib.vec(self.N, idx, val)
        self.lib.call(self.lib.L.sleqp_fact_solve(self.h, v))
        self.lib.L.sleqp_vec_free(C.byref(v))

    def solution(self, begin, end, zero_eps=1e-20):
        r = C.POINTER(SleqpVec)()
        self.lib.call(self.lib.L.sleqp_vec_create_empty(C.byref(r), int(end - begin)))
        self.lib.call(self.lib.L.sleqp_fact_solution(self.h, r, int(begin), int(end), float(zero_eps)))
        idx, val, dim = self.lib.vec_to_numpy(r)
        self.lib.L.sleqp_vec_free(C.byref(r))
        return idx, val

    def cond(self):
        c = C.c_double()
        self.lib.call(self.lib.L.sleqp_fact_cond(self.h, C.byref(c)))
        return c.value

    def release(self):
        L = self.lib.L
        if self.h:
            self.lib.call(L.sleqp_fact_release(C.byref(self.h)))
            self.h = C.c_void_p()
        if self.mat is not None:
            L.sleqp_mat_release(C.byref(self.mat))
            self.mat = None
        if self.settings:
            L.sleqp_settings_release(C.byref(self.settings))
            self.settings = C.c_void_p()
