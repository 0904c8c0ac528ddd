"""Host-side mirror of the reference's factorization plugin and sparse mat-vec interface.

Names, argument meaning and error behaviour follow the reference (chrhansk/sleqp v1.0.2):

  Fact.set_matrix / solve / solution / cond / release  <->  sleqp_fact_set_matrix / _solve / _solution /
      _cond / _release (src/main/fact/fact.h:36-68); flags() returns SLEQP_FACT_FLAGS_LOWER only
      (fact.h:9-14), so callers hand over the lower triangle and the indefinite standard form
      (trial_point.c:94-109).
  Mat.mult_vec / mult_vec_trans  <->  sleqp_mat_mult_vec / sleqp_mat_mult_vec_trans (sparse/mat.c:282-363)

Every method goes through the C-ABI (include/sleqp_b200.h); there is no Python or CPU compute path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Stats, check, lib

SLEQP_FACT_FLAGS_NONE = 0
SLEQP_FACT_FLAGS_PSD = 1 << 0
SLEQP_FACT_FLAGS_LOWER = 1 << 1

_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pi(a):
    return a.ctypes.data_as(_ip)


def _pd(a):
    return a.ctypes.data_as(_dp)


PLAN_FIELDS = {
    # name: dtype
    **{k: np.int32 for k in (
        "e_of_k r_of_k k_of_e k_of_r dE_src Acsc_ptr Acsc_row Acsc_src Acsr_ptr Acsr_col Acsr_src "
        "Gsym_ptr Gsym_col Gsym_src perm pinv parent colcount sn_first sn_of_col sn_parent sn_level "
        "Ridx rel child_ptr child_idx Sgsrc Sterm_a Sterm_b Sterm_d sn_base sn_nt zero_sn lvl_ptr lvl_sn "
        "inv_phase_ptr Ksrc sst_colptr sst_rows sst_blob sst_ea_dst sst_gen_ptr sn_sparse"
    ).split()},
    **{k: np.int64 for k in "Rptr Lptr Wptr Sdest Sterm_ptr Uoff Tptr sst_ea_src".split()},
}
PLAN_STRUCTS = {"stages": 10, "ea_tasks": 2, "pan_tasks": 4, "upd_tasks": 7, "inv_tasks": 6,
                "ffl_tasks": 16, "bfl_tasks": 16, "tr_tasks": 3, "sst": 30}  # int32 columns


class Symbolic:
    """Host-only symbolic analysis (no GPU): ordering, elimination tree, supernodes, schedule."""

    def __init__(self, n, colptr, rowidx, val, lower_only=True):
        self._h = C.c_void_p()
        colptr, rowidx, val = _i32(colptr), _i32(rowidx), _f64(val)
        check(lib().b200_symbolic_analyze(C.byref(self._h), int(n), int(len(rowidx)), _pi(colptr), _pi(rowidx), _pd(val), int(bool(lower_only))))
        self.n = int(n)

    @classmethod
    def from_kkt(cls, num_vars, num_cons, jac_cols, jac_rows, jac_data, var_index, cons_index, working_set_size):
        """Analysis of the KKT system b200_fact_set_kkt builds from (constraint Jacobian, working-set index maps)."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        jac_cols, jac_rows, jac_data = _i32(jac_cols), _i32(jac_rows), _f64(jac_data)
        var_index, cons_index = _i32(var_index), _i32(cons_index)
        check(lib().b200_symbolic_analyze_kkt(C.byref(self._h), int(num_vars), int(num_cons), int(len(jac_rows)), _pi(jac_cols), _pi(jac_rows), _pd(jac_data),
                                              _pi(var_index), _pi(cons_index), int(working_set_size)))
        self.n = int(num_vars + working_set_size)
        return self

    def stats(self) -> dict:
        s = Stats()
        check(lib().b200_symbolic_stats(self._h, C.byref(s)))
        return s.as_dict()

    def structure(self):
        """(perm, parent, colcount, super_first) over the full order of K."""
        n = self.n
        perm, parent, cc = (np.empty(n, dtype=np.int32) for _ in range(3))
        ns = C.c_int()
        sf = np.empty(n + 2, dtype=np.int32)
        check(lib().b200_symbolic_structure(self._h, _pi(perm), _pi(parent), _pi(cc), C.byref(ns), _pi(sf)))
        return perm, parent, cc, sf[: ns.value + 1].copy()

    def export(self, field):
        cnt = C.c_int64()
        check(lib().b200_symbolic_export(self._h, field.encode(), None, C.byref(cnt)))
        if field in PLAN_STRUCTS:
            out = np.empty((cnt.value, PLAN_STRUCTS[field]), dtype=np.int32)
        else:
            out = np.empty(cnt.value, dtype=PLAN_FIELDS[field])
        if cnt.value:
            check(lib().b200_symbolic_export(self._h, field.encode(), out.ctypes.data_as(C.c_void_p), C.byref(cnt)))
        return out

    def plan(self) -> dict:
        d = {k: self.export(k) for k in list(PLAN_FIELDS) + list(PLAN_STRUCTS)}
        d.update(self.stats())
        return d

    def close(self):
        if self._h:
            lib().b200_symbolic_free(C.byref(self._h))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Fact:
    """The B200 factorization backend behind SleqpFactCallbacks (fact_types.h:25-32)."""

    name = "B200"

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        check(lib().b200_fact_create(C.byref(self._h), int(device)))
        self._n = 0

    @staticmethod
    def flags() -> int:
        return SLEQP_FACT_FLAGS_LOWER

    def set_matrix(self, n, colptr, rowidx, val, lower_only=True):
        """sleqp_fact_set_matrix (fact.c:59-75): symbolic (cached per pattern) + numeric LDL^T."""
        colptr, rowidx, val = _i32(colptr), _i32(rowidx), _f64(val)
        check(lib().b200_fact_set_matrix(self._h, int(n), int(n), int(len(rowidx)), _pi(colptr), _pi(rowidx), _pd(val), int(bool(lower_only))))
        self._n = int(n)

    def set_kkt(self, num_vars, num_cons, jac_cols, jac_rows, jac_data, var_index, cons_index, working_set_size):
        """sleqp_aug_jac_set_iterate with the KKT assembly on the device (b200_fact_set_kkt): the constraint Jacobian
        (CSC) and the working set as index maps (-1 = not in the working set)."""
        jac_cols, jac_rows, jac_data = _i32(jac_cols), _i32(jac_rows), _f64(jac_data)
        var_index, cons_index = _i32(var_index), _i32(cons_index)
        check(lib().b200_fact_set_kkt(self._h, int(num_vars), int(num_cons), int(len(jac_rows)), _pi(jac_cols), _pi(jac_rows), _pd(jac_data),
                                      _pi(var_index), _pi(cons_index), int(working_set_size)))
        self._n = int(num_vars + working_set_size)

    def solve(self, idx, val, dim=None, offset=0):
        """sleqp_fact_solve (fact.c:83-89) with a sparse right-hand side of dimension `dim`."""
        idx, val = _i32(idx), _f64(val)
        check(lib().b200_fact_solve_offset(self._h, int(len(idx)), _pi(idx), _pd(val), int(offset), int(self._n if dim is None else dim)))

    def solution_dense(self, begin, end) -> np.ndarray:
        out = np.empty(int(end - begin), dtype=np.float64)
        check(lib().b200_fact_solution(self._h, int(begin), int(end), _pd(out)))
        return out

    def solution(self, begin, end, zero_eps=1e-20):
        """sleqp_fact_solution (fact.c:91-102): sparse slice, entries with |v| <= zero_eps dropped
        (sleqp_vec_set_from_raw, vec.c:72-104). Returns (indices, values)."""
        n = int(end - begin)
        if getattr(self, "_sol_cap", 0) < n:  # grown on demand and reused, like the caller-owned SleqpVec
            self._unpin_solution_buffers()
            self._sol_cap = max(n, 1)
            self._sol_idx = np.empty(self._sol_cap, dtype=np.int32)
            self._sol_val = np.empty(self._sol_cap, dtype=np.float64)
            # page-locked once: the sparsified slice is then DMA'd straight into them (b200_host_pin)
            self._sol_pinned = [a for a in (self._sol_idx, self._sol_val) if lib().b200_host_pin(a.ctypes.data_as(C.c_void_p), a.nbytes) == 0]
        idx, val = self._sol_idx, self._sol_val
        nnz = C.c_int()
        check(lib().b200_fact_solution_sparse(self._h, int(begin), int(end), float(zero_eps), _pi(idx), _pd(val), C.byref(nnz)))
        return idx[: nnz.value], val[: nnz.value]

    def _unpin_solution_buffers(self):
        for a in getattr(self, "_sol_pinned", []):
            lib().b200_host_unpin(a.ctypes.data_as(C.c_void_p))
        self._sol_pinned = []

    def solve_device(self, d_rhs_ptr: int, d_sol_ptr: int):
        check(lib().b200_fact_solve_device(self._h, C.c_void_p(d_rhs_ptr), C.c_void_p(d_sol_ptr)))

    def refactor_device(self, d_val_ptr: int):
        """Numeric refactorization from values resident in device memory (same pattern)."""
        check(lib().b200_fact_refactor_device(self._h, C.c_void_p(d_val_ptr)))

    def profile_solve(self, reps=20):
        """Mean device ms of (E-elimination, forward sweep, backward sweep, back-substitution)."""
        out = np.zeros(4, dtype=np.float64)
        check(lib().b200_fact_profile_solve(self._h, int(reps), _pd(out)))
        return out

    def profile_numeric(self) -> dict:
        """Device ms of one eager numeric factorization by kernel class."""
        out = np.zeros(8, dtype=np.float64)
        check(lib().b200_fact_profile_numeric(self._h, _pd(out)))
        return dict(zip(("assemble", "zero", "extend_add", "panel", "update", "inv_gemm", "transpose", "rest"), out.tolist()))

    def cond(self) -> float:
        """sleqp_fact_cond (fact.c:104-118): 1 / rcond."""
        r = C.c_double()
        check(lib().b200_fact_rcond(self._h, C.byref(r)))
        return float("inf") if r.value == 0.0 else 1.0 / r.value

    def stats(self) -> dict:
        s = Stats()
        check(lib().b200_fact_stats(self._h, C.byref(s)))
        return s.as_dict()

    def structure(self):
        n = self._n
        perm, parent, cc = (np.empty(n, dtype=np.int32) for _ in range(3))
        ns = C.c_int()
        sf = np.empty(n + 2, dtype=np.int32)
        check(lib().b200_fact_structure(self._h, _pi(perm), _pi(parent), _pi(cc), C.byref(ns), _pi(sf)))
        return perm, parent, cc, sf[: ns.value + 1].copy()

    def pivots(self) -> np.ndarray:
        d = np.empty(self._n, dtype=np.float64)
        check(lib().b200_fact_pivots(self._h, _pd(d)))
        return d

    @property
    def stream(self) -> int:
        return int(lib().b200_fact_stream(self._h) or 0)

    def release(self):
        """sleqp_fact_release (fact.c:143-161) -> callbacks.free."""
        if self._h:
            self._unpin_solution_buffers()
            check(lib().b200_fact_free(C.byref(self._h)))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Mat:
    """Device-resident CSC matrix for the Jacobian/Hessian products of the EQP loop."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        check(lib().b200_mat_create(C.byref(self._h), int(device)))
        self.num_rows = self.num_cols = 0

    def set(self, num_rows, num_cols, cols, rows, data):
        cols, rows, data = _i32(cols), _i32(rows), _f64(data)
        check(lib().b200_mat_set(self._h, int(num_rows), int(num_cols), int(len(rows)), _pi(cols), _pi(rows), _pd(data)))
        self.num_rows, self.num_cols = int(num_rows), int(num_cols)

    def mult_vec(self, idx, val, out=None) -> np.ndarray:
        """sleqp_mat_mult_vec (mat.c:282-310): dense result of length num_rows."""
        idx, val = _i32(idx), _f64(val)
        out = np.empty(self.num_rows, dtype=np.float64) if out is None else out
        check(lib().b200_mat_mult_vec(self._h, int(len(idx)), _pi(idx), _pd(val), _pd(out)))
        return out

    def mult_vec_trans(self, idx, val, eps=0.0, out=None):
        """sleqp_mat_mult_vec_trans (mat.c:312-363): sparse result, |s| <= eps dropped."""
        idx, val = _i32(idx), _f64(val)
        out = np.empty(self.num_cols, dtype=np.float64) if out is None else out
        check(lib().b200_mat_mult_vec_trans(self._h, int(len(idx)), _pi(idx), _pd(val), _pd(out)))
        keep = np.nonzero(np.abs(out) > eps)[0].astype(np.int32)
        return keep, out[keep]

    def mult_vec_trans_sparse(self, idx, val, eps=0.0):
        """Same product, sparsified on the device: only the kept entries come back (b200_mat_mult_vec_trans_sparse)."""
        idx, val = _i32(idx), _f64(val)
        oi = np.empty(max(self.num_cols, 1), dtype=np.int32)
        ov = np.empty(max(self.num_cols, 1), dtype=np.float64)
        nnz = C.c_int()
        check(lib().b200_mat_mult_vec_trans_sparse(self._h, int(len(idx)), _pi(idx), _pd(val), float(eps), _pi(oi), _pd(ov), C.byref(nnz)))
        return oi[: nnz.value].copy(), ov[: nnz.value].copy()

    def set_stream(self, stream_ptr: int):
        check(lib().b200_mat_set_stream(self._h, C.c_void_p(stream_ptr)))

    def mult_vec_device(self, d_x: int, d_y: int):
        check(lib().b200_mat_mult_vec_device(self._h, C.c_void_p(d_x), C.c_void_p(d_y)))

    def mult_vec_trans_device(self, d_v: int, d_y: int):
        check(lib().b200_mat_mult_vec_trans_device(self._h, C.c_void_p(d_v), C.c_void_p(d_y)))

    @property
    def stream(self) -> int:
        return int(lib().b200_mat_stream(self._h) or 0)

    def release(self):
        if self._h:
            check(lib().b200_mat_free(C.byref(self._h)))

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class ProjectedCG:
    """Device-resident Steihaug projected CG (SleqpTRSolver, src/main/tr/steihaug_solver.c:223-496) over a
    factorized KKT system and a Hessian held in device memory: min g^T p + 1/2 p^T H p, A_W p = 0, |p| <= radius."""

    INTERIOR, BOUNDARY, NEG_CURVATURE, MAX_ITER = range(4)

    HESS_PROD = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _dp, _dp)

    def __init__(self, fact: Fact, hess: Mat | None = None, hess_prod=None):
        """hess: the Hessian of the Lagrangian as a device matrix, or hess_prod: a host function d -> H d (the
        reference's matrix-free SLEQP_FUNC_HESS_PROD callback)."""
        self._h = C.c_void_p()
        self._fact, self._hess = fact, hess  # keep the borrowed handles alive
        check(lib().b200_cg_create(C.byref(self._h), fact._h, hess._h if hess is not None else None))
        self._cb = None
        if hess is None:
            def _cb(_ctx, n, d, out):
                try:
                    np.ctypeslib.as_array(out, shape=(n,))[:] = hess_prod(np.ctypeslib.as_array(d, shape=(n,)))
                    return 0
                except Exception:  # noqa: BLE001
                    return 1

            self._cb = self.HESS_PROD(_cb)
            check(lib().b200_cg_set_hess_callback(self._h, C.cast(self._cb, C.c_void_p), None))

    def solve_ex(self, n, g_idx, g_val, trust_radius, rel_tol=1e-8, max_iter=100):
        """Returns (step[n], iterations, termination, tr_dual, min_rayleigh, max_rayleigh): everything
        SleqpTRCallbacks.solve / .rayleigh hand back (tr/tr_types.h:9-30)."""
        g_idx, g_val = _i32(g_idx), _f64(g_val)
        step = np.empty(int(n), dtype=np.float64)
        it, term = C.c_int(), C.c_int()
        dual, rmin, rmax = C.c_double(), C.c_double(), C.c_double()
        check(lib().b200_cg_solve_ex(self._h, int(n), int(len(g_idx)), _pi(g_idx), _pd(g_val), float(trust_radius), float(rel_tol),
                                     int(max_iter), _pd(step), C.byref(it), C.byref(term), C.byref(dual), C.byref(rmin), C.byref(rmax)))
        return step, it.value, term.value, dual.value, rmin.value, rmax.value

    def solve_sparse(self, n, g_idx, g_val, trust_radius, rel_tol=1e-8, max_iter=100, zero_eps=0.0, pinned=False):
        """The step as (indices, values) with the contract of sleqp_vec_set_from_raw (|v| > zero_eps, ascending), plus
        (iterations, termination, tr_dual, min_rayleigh, max_rayleigh). pinned=True page-locks the arrays first: the
        step is then sparsified on the device and DMA'd into them (what host/tr/tr_b200.c does)."""
        g_idx, g_val = _i32(g_idx), _f64(g_val)
        idx = np.empty(int(n), dtype=np.int32)
        val = np.empty(int(n), dtype=np.float64)
        bufs = (g_idx, g_val, idx, val) if pinned else ()
        for b in bufs:
            check(lib().b200_host_pin(b.ctypes.data_as(C.c_void_p), b.nbytes))
        try:
            it, term, nnz = C.c_int(), C.c_int(), C.c_int()
            dual, rmin, rmax = C.c_double(), C.c_double(), C.c_double()
            check(lib().b200_cg_solve_sparse(self._h, int(n), int(len(g_idx)), _pi(g_idx), _pd(g_val), float(trust_radius), float(rel_tol),
                                             int(max_iter), float(zero_eps), _pi(idx), _pd(val), C.byref(nnz), C.byref(it), C.byref(term),
                                             C.byref(dual), C.byref(rmin), C.byref(rmax)))
        finally:
            for b in bufs:
                lib().b200_host_unpin(b.ctypes.data_as(C.c_void_p))
        return idx[: nnz.value].copy(), val[: nnz.value].copy(), it.value, term.value, dual.value, rmin.value, rmax.value

    def solve(self, n, g_idx, g_val, trust_radius, rel_tol=1e-8, max_iter=100):
        """Returns (step[n], iterations, termination)."""
        g_idx, g_val = _i32(g_idx), _f64(g_val)
        step = np.empty(int(n), dtype=np.float64)
        it, term = C.c_int(), C.c_int()
        check(lib().b200_cg_solve(self._h, int(n), int(len(g_idx)), _pi(g_idx), _pd(g_val), float(trust_radius), float(rel_tol),
                                  int(max_iter), _pd(step), C.byref(it), C.byref(term)))
        return step, it.value, term.value

    def release(self):
        if self._h:
            check(lib().b200_cg_free(C.byref(self._h)))
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
