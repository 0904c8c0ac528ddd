"""Device KKT assembly: the KKT matrix is built from the constraint
Jacobian and the working-set index maps the way the reference's fill_aug_jac does it (standard_aug_jac.c:135-237,
restated and pinned in oracle.fill_aug_jac). The host lays the pattern out only when (Jacobian pattern, working set)
is new; the values are gathered on the device."""
import numpy as np
import pytest

from oracle import sleqp_oracle as orc
from sleqp_b200 import B200Error, Fact, Symbolic, problems

CASES = {
    "config1": lambda: problems.config(0),
    "poisson2d_g16": lambda: problems.poisson_control(16, 2, seed=2),
    "chain_n3000_active": lambda: problems.chain_rosenbrock(3000, 0.2, seed=4),
    "poisson3d_g6": lambda: problems.poisson_control(6, 3, seed=1),
}


def _maps(p):
    J = p.J.tocsc()
    J.sort_indices()
    vi, ci, ws = orc.working_set_indices(p.n, p.m, p.active_vars, p.active_cons)
    return J, vi, ci, ws


@pytest.mark.parametrize("name", list(CASES))
def test_kkt_layout_equals_fill_aug_jac(name):
    p = CASES[name]()
    J, vi, ci, ws = _maps(p)
    cp, ri, v = orc.fill_aug_jac(p.n, J.indptr, J.indices, J.data, vi, ci, ws)
    a = Symbolic.from_kkt(p.n, p.m, J.indptr, J.indices, J.data, vi, ci, ws)
    b = Symbolic(p.N, cp, ri, v)
    # same pattern -> same analysis, bit for bit
    for x, y in zip(a.structure(), b.structure()):
        assert np.array_equal(x, y)
    assert a.stats()["perm_hash"] == b.stats()["perm_hash"] and a.stats()["nnz_K"] == len(ri)
    # the value map reproduces fill_aug_jac's values in its order
    src = a.export("Ksrc")
    assert len(src) == len(v)
    got = np.where(src < 0, 1.0, J.data[np.maximum(src, 0)])
    assert np.array_equal(got, v)
    # and equals what problems.kkt_lower hands to set_matrix
    cp2, ri2, v2 = p.kkt_lower()
    assert np.array_equal(cp, cp2) and np.array_equal(ri, ri2) and np.array_equal(v, v2)


def test_malformed_maps_are_rejected():
    p = problems.config(0)
    J, vi, ci, ws = _maps(p)
    bad = ci.copy()
    bad[3] = ws + 5  # index outside the working set
    with pytest.raises(B200Error):
        Symbolic.from_kkt(p.n, p.m, J.indptr, J.indices, J.data, vi, bad, ws)
    swapped = ci.copy()
    swapped[[0, 1]] = swapped[[1, 0]]  # rows of a column would not come out increasing
    with pytest.raises(B200Error):
        Symbolic.from_kkt(p.n, p.m, J.indptr, J.indices, J.data, vi, swapped, ws)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_set_kkt_equals_set_matrix(name):
    p = CASES[name]()
    J, vi, ci, ws = _maps(p)
    f, g = Fact(), Fact()
    f.set_matrix(p.N, *p.kkt_lower())
    g.set_kkt(p.n, p.m, J.indptr, J.indices, J.data, vi, ci, ws)
    assert np.abs(f.pivots() - g.pivots()).max() <= 1e-12 * np.abs(f.pivots()).max()
    K = p.kkt_full()
    for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
        idx, val = p.rhs(kind, 3)
        g.solve(idx, val, p.N)
        x = g.solution_dense(0, p.N)
        b = orc.vec_to_raw(idx, val, p.N)
        assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b)
    # min-norm right-hand side given in working-set coordinates + offset (standard_aug_jac.c:328-345 without the mutation)
    idx, val = p.rhs("solve_min_norm", 4)
    g.solve(idx - p.n, val, p.N, offset=p.n)
    x = g.solution_dense(0, p.N)
    assert np.linalg.norm(K @ x - orc.vec_to_raw(idx, val, p.N)) <= 1e-10 * np.linalg.norm(val)


@pytest.mark.gpu
def test_set_kkt_walks_a_sequence_of_working_sets():
    """20 working sets in a row on one handle (active bounds move, the Jacobian values change every time): every
    factorization solves its own K; a working set seen before reuses its cached analysis, new values only cost the
    device gather + numeric factorization."""
    rng = np.random.default_rng(5)
    base = problems.poisson_control(20, 2, seed=1)
    J = base.J.tocsc()
    J.sort_indices()
    q = base.m
    f = Fact()
    seen = {}
    for step in range(20):
        na = int(rng.integers(0, q // 4))
        active = np.sort(rng.choice(q, size=na, replace=False)) + q if step % 5 else np.sort(base.active_vars)
        data = J.data * rng.uniform(0.9, 1.1, size=len(J.data))
        p = problems.KKTProblem(name="walk", n=base.n, m=base.m, J=type(J)((data, J.indices, J.indptr), shape=J.shape), H=base.H,
                                active_vars=active.astype(np.int64), active_cons=base.active_cons)
        vi, ci, ws = orc.working_set_indices(p.n, p.m, p.active_vars, p.active_cons)
        f.set_kkt(p.n, p.m, J.indptr, J.indices, data, vi, ci, ws)
        key = tuple(active.tolist())
        assert f.stats()["symbolic_cached"] == (1 if key in seen else 0)
        seen[key] = True
        K = p.kkt_full()
        idx, val = p.rhs("project_nullspace", step)
        f.solve(idx, val, p.N)
        x = f.solution_dense(0, p.N)
        b = orc.vec_to_raw(idx, val, p.N)
        assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b), step
        assert np.abs(p.working_rows() @ x[: p.n]).max() <= 1e-10 * np.abs(val).max()
