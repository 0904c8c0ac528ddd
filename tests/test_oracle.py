"""The CPU oracle (oracle/sleqp_oracle.py) pinned against (a) the reference's own known-answer
tests, (b) fixtures produced by running the unmodified reference (tests/golden/make_golden.py),
and (c) the reference library itself when oracle/_ref is present."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import ref_lib
from oracle import sleqp_oracle as orc
from sleqp_b200 import problems


def test_spmv_known_answer():
    # src/test/sparse/sleqp_sparse_matrix_test.c:12-56: 2x3 CSC times (2,4,3) = (8,17), tol 1e-8
    colptr, rows, data = [0, 1, 2, 4], [0, 1, 0, 1], [1.0, 2.0, 2.0, 3.0]
    y = orc.mat_mult_vec(2, colptr, rows, data, [0, 1, 2], [2.0, 4.0, 3.0])
    assert abs(y[0] - 8.0) <= 1e-8 and abs(y[1] - 17.0) <= 1e-8


def test_spmv_against_reference_fixtures(golden):
    g = golden("spmv_reference.npz")
    for name in g["names"]:
        m, n = g[f"{name}_shape"]
        cp, ri, d = g[f"{name}_colptr"], g[f"{name}_rows"], g[f"{name}_data"]
        y = orc.mat_mult_vec(m, cp, ri, d, g[f"{name}_xi"], g[f"{name}_xv"])
        assert np.array_equal(y, g[f"{name}_y"]), name  # same accumulation order => bit-exact
        ti, tv = orc.mat_mult_vec_trans(n, cp, ri, d, g[f"{name}_vi"], g[f"{name}_vv"], m, 1e-12)
        assert np.array_equal(ti, g[f"{name}_ti"]), name
        assert np.array_equal(tv, g[f"{name}_tv"]), name


def test_vec_set_from_raw_fixture(golden):
    g = golden("vec_reference.npz")
    for eps in (0.0, 1e-20, 1e-3):
        i, x = orc.vec_set_from_raw(g["values"], eps)
        assert np.array_equal(i, g[f"idx_{eps}"]) and np.array_equal(x, g[f"val_{eps}"])


def test_newton_known_answer():
    # src/test/constrained_newton_test.c:204-275: K = [I J^T; J 0] with J = [0 1] (3x3); the
    # projection of (2,4) onto null(J) is (2,0)
    colptr, rows, data = orc.fill_aug_jac(2, [0, 0, 1], [0], [1.0], np.array([-1, -1]), np.array([0]), 1, lower_only=True)
    assert colptr.tolist() == [0, 1, 3, 3] and rows.tolist() == [0, 1, 2] and data.tolist() == [1.0, 1.0, 1.0]
    idx, val, dim, b, e = orc.aug_jac_rhs("project_nullspace", 2, 1, [0, 1], [2.0, 4.0])
    x = orc.kkt_solve_dense(3, colptr, rows, data, orc.vec_to_raw(idx, val, dim))
    si, sv = orc.vec_set_from_raw(x[b:e], 1e-20)
    assert si.tolist() == [0] and abs(sv[0] - 2.0) <= 1e-8


def test_dual_estimation_known_answer():
    # src/test/dual_estimation_test.c:15-103: quadfunc fixture, two active lower bounds,
    # gradient (2,4) at x=(1,2): LSQ multipliers of min |g + A^T lam| with A = I give -g.
    n = 2
    vi, ci, ws = orc.working_set_indices(2, 0, [0, 1], [])
    colptr, rows, data = orc.fill_aug_jac(n, [0, 0, 0], [], [], vi, ci, ws)
    # solve_lsq([r;0]) returns (A A^T)^-1 A r; the reference negates the gradient first
    # (dual_estimation_lsq.c:37-47), so with r = -g the duals are (-2,-4)
    idx, val, dim, b, e = orc.aug_jac_rhs("solve_lsq", n, ws, [0, 1], [-2.0, -4.0])
    x = orc.kkt_solve_dense(n + ws, colptr, rows, data, orc.vec_to_raw(idx, val, dim))
    assert np.allclose(x[b:e], [-2.0, -4.0], atol=1e-8)


def test_fact_against_reference_lapack_fixtures(golden):
    g = golden("fact_reference_lapack.npz")
    for name in g["names"]:
        n, ws = g[f"{name}_n"]
        N = n + ws
        cp, ri, d = g[f"{name}_colptr"], g[f"{name}_rows"], g[f"{name}_data"]
        lu = orc.SparseLU()
        lu.set_matrix(N, cp, ri, d)
        for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
            idx, val = g[f"{name}_{kind}_idx"], g[f"{name}_{kind}_val"]
            b, e = g[f"{name}_{kind}_range"]
            xd = orc.kkt_solve_dense(N, cp, ri, d, orc.vec_to_raw(idx, val, N))
            ref = orc.vec_to_raw(g[f"{name}_{kind}_si"], g[f"{name}_{kind}_sv"], e - b)
            scale = max(1.0, np.abs(ref).max())
            assert np.abs(xd[b:e] - ref).max() <= 1e-9 * scale, (name, kind)
            lu.solve(idx, val)
            si, sv = lu.solution(b, e, 1e-20)
            assert np.abs(orc.vec_to_raw(si, sv, e - b) - ref).max() <= 1e-9 * scale, (name, kind)


@pytest.mark.parametrize("lower_only", [True, False])
def test_fill_aug_jac_matches_independent_assembly(lower_only):
    for p in (problems.config(0), problems.poisson_control(6, 2, seed=2), problems.chain_rosenbrock(60, 0.3, seed=9)):
        vi, ci, ws = orc.working_set_indices(p.n, p.m, p.active_vars, p.active_cons)
        J = p.J
        cp, ri, d = orc.fill_aug_jac(p.n, J.indptr, J.indices, J.data, vi, ci, ws, lower_only)
        K = sp.csc_matrix((d, ri, cp), shape=(p.N, p.N))
        full = p.kkt_full()
        want = sp.tril(full) if lower_only else full
        assert (abs(K - want)).max() == 0.0
        # rows strictly increasing per column (sleqp_mat_is_valid, mat.c:797-804)
        for j in range(p.N):
            assert np.all(np.diff(ri[cp[j]:cp[j + 1]]) > 0)
        if lower_only:
            cp2, ri2, d2 = p.kkt_lower()
            assert np.array_equal(cp, cp2) and np.array_equal(ri, ri2) and np.array_equal(d, d2)


def test_partial_working_set_rows():
    # only some constraints active: inactive Jacobian rows must be dropped (standard_aug_jac.c:189-216)
    p = problems.chain_rosenbrock(30, 0.2, seed=1)
    p.active_cons = p.active_cons[::2]
    vi, ci, ws = orc.working_set_indices(p.n, p.m, p.active_vars, p.active_cons)
    cp, ri, d = orc.fill_aug_jac(p.n, p.J.indptr, p.J.indices, p.J.data, vi, ci, ws, True)
    K = sp.csc_matrix((d, ri, cp), shape=(p.N, p.N))
    assert (abs(K - sp.tril(p.kkt_full()))).max() == 0.0


@pytest.mark.skipif(not ref_lib.available("lapack"), reason="oracle/_ref not built (reference absent)")
def test_oracle_against_live_reference():
    ref = ref_lib.RefLib("lapack")
    rng = np.random.default_rng(5)
    A = sp.random(50, 35, density=0.1, random_state=np.random.RandomState(3), format="csc")
    A.sort_indices()
    xi = np.sort(rng.choice(35, 20, replace=False)).astype(np.int32)
    xv = rng.standard_normal(20)
    assert np.array_equal(ref.mat_mult_vec(50, 35, A.indptr, A.indices, A.data, xi, xv), orc.mat_mult_vec(50, A.indptr, A.indices, A.data, xi, xv))
    vi = np.sort(rng.choice(50, 17, replace=False)).astype(np.int32)
    vv = rng.standard_normal(17)
    ri, rv = ref.mat_mult_vec_trans(50, 35, A.indptr, A.indices, A.data, vi, vv, 1e-10)
    oi, ov = orc.mat_mult_vec_trans(35, A.indptr, A.indices, A.data, vi, vv, 50, 1e-10)
    assert np.array_equal(ri, oi) and np.array_equal(rv, ov)
    p = problems.chain_rosenbrock(80, 0.25, seed=12)
    cp, r, d = p.kkt_lower()
    f = ref.fact()
    f.set_matrix(p.N, cp, r, d)
    idx, val = p.rhs("project_nullspace", 3)
    f.solve(idx, val)
    si, sv = f.solution(0, p.n, 1e-20)
    x = orc.kkt_solve_dense(p.N, cp, r, d, orc.vec_to_raw(idx, val, p.N))
    assert np.abs(orc.vec_to_raw(si, sv, p.n) - x[: p.n]).max() <= 1e-10
    f.release()
