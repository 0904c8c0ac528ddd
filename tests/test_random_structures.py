"""Irregular working sets: random sparse Jacobians (dense rows, duplicated patterns, disconnected blocks, cliques,
tiny systems). The host analysis + plan emulation run on the CPU; the device path is checked against a sparse LU."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import sleqp_oracle as orc
from oracle.multifrontal_emul import Emulated
from sleqp_b200 import Symbolic, problems


def random_problem(seed):
    rng = np.random.default_rng(seed)
    kind = seed % 6
    if kind == 0:  # generic sparse rows
        n, m = int(rng.integers(20, 400)), 0
        m = int(rng.integers(1, max(2, n // 2)))
        J = sp.random(m, n, density=min(1.0, 4.0 / n), random_state=np.random.RandomState(seed), format="lil")
    elif kind == 1:  # a few dense rows couple everything (clique in S)
        n = int(rng.integers(30, 200))
        m = int(rng.integers(3, 12))
        J = sp.lil_matrix(rng.standard_normal((m, n)))
    elif kind == 2:  # block diagonal: disconnected components
        blocks = [sp.random(int(rng.integers(1, 8)), int(rng.integers(8, 30)), density=0.4, random_state=np.random.RandomState(seed + b)) for b in range(int(rng.integers(2, 9)))]
        J = sp.block_diag(blocks, format="lil")
        m, n = J.shape
    elif kind == 3:  # banded with long range couplings
        n = int(rng.integers(100, 600))
        m = n // 3
        J = sp.diags([rng.standard_normal(m), rng.standard_normal(m), rng.standard_normal(m)], [0, 1, n // 2], shape=(m, n), format="lil")
    elif kind == 4:  # tiny
        n, m = int(rng.integers(1, 6)), 1
        J = sp.lil_matrix(rng.standard_normal((m, n)))
    else:  # wide 2D-like stencil rows with random weights
        g = int(rng.integers(6, 18))
        p0 = problems.poisson_control(g, 2, seed=seed)
        J = p0.J.tolil()
        m, n = J.shape
    J = J.tocsr()
    # make every row non-empty and the rows independent: add a private column entry per row
    m, n = J.shape
    priv = rng.choice(n, size=m, replace=False) if m <= n else None
    if priv is None:
        J = J[:n]
        m = n
        priv = rng.permutation(n)
    J = (J + sp.csr_matrix((3.0 + rng.random(m), (np.arange(m), priv)), shape=(m, n))).tocsc()
    J.sort_indices()
    free = np.setdiff1d(np.arange(n), priv)
    na = int(rng.integers(0, max(1, len(free) // 3 + 1)))
    active_vars = np.sort(rng.choice(free, size=na, replace=False)) if na and len(free) else np.zeros(0, dtype=np.int64)
    return problems.KKTProblem(name=f"random_{seed}", n=n, m=m, J=J, H=sp.identity(n, format="csc"), active_vars=active_vars.astype(np.int64),
                               active_cons=np.arange(m, dtype=np.int64))


SEEDS = list(range(24))


@pytest.mark.parametrize("seed", SEEDS)
def test_analysis_and_plan_on_random_structures(seed):
    p = random_problem(seed)
    cp, ri, v = p.kkt_lower()
    s = Symbolic(p.N, cp, ri, v)
    perm, parent, cc, sf = s.structure()
    par_ref, cc_ref = orc.symbolic_reference(p.N, cp, ri, perm)
    assert np.array_equal(parent, par_ref) and np.array_equal(cc, cc_ref)
    if p.N <= 700:
        em = Emulated(s.plan(), v)
        K = p.kkt_full().tocsc()
        b = np.random.default_rng(seed).standard_normal(p.N)
        z = em.solve(b, refine=1)
        zr = spla.spsolve(K, b)
        assert np.linalg.norm(K @ z - b) <= 1e-9 * np.linalg.norm(b)
        assert np.linalg.norm(z - zr) <= 1e-7 * max(1.0, np.linalg.norm(zr))
        zs = em.solve(b, refine=1, flow=False)  # plain substitution with the factor as a cross-check of the task lists
        assert np.linalg.norm(zs - z) <= 1e-9 * max(1.0, np.linalg.norm(z))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", SEEDS)
def test_device_path_on_random_structures(seed):
    from sleqp_b200 import Fact

    p = random_problem(seed)
    f = Fact()
    f.set_matrix(p.N, *p.kkt_lower())
    K = p.kkt_full().tocsc()
    rng = np.random.default_rng(seed)
    for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
        idx, val = p.rhs(kind, seed)
        f.solve(idx, val, p.N)
        x = f.solution_dense(0, p.N)
        b = orc.vec_to_raw(idx, val, p.N)
        assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b), (kind, f.stats()["probe_residual"])
    f.release()


# ---- candidates for elimination that stay in the reduced system -------------------------------
def general_kkt(kind, n, m, seed):
    """tril(K) of K = [H A^T; A 0] outside what standard_aug_jac.c builds: `coupled` has a tridiagonal (1,1) block (any
    two neighbouring variables are coupled: they cannot both be eliminated in closed form), `dense_column` has the
    identity there but one variable that appears in every constraint (its elimination would make S dense)."""
    rng = np.random.default_rng(seed)
    if n <= 1000:  # random rows (fill of a random graph: small sizes only)
        A = sp.random(m, n, density=3.0 / n, random_state=np.random.RandomState(seed), format="lil")
        priv = rng.choice(n - 1, size=m, replace=False) + 1
        A = (A.tocsr() + sp.csr_matrix((3.0 + rng.random(m), (np.arange(m), priv)), shape=(m, n))).tolil()
    else:  # chain-like rows: row i couples the variables 2 i + 1 .. 2 i + 3
        assert 2 * m + 3 <= n
        ii = np.repeat(np.arange(m), 3)
        jj = (2 * np.arange(m)[:, None] + np.array([1, 2, 3])).ravel()
        A = sp.csr_matrix((rng.standard_normal(3 * m) + np.tile([3.0, 0.0, 0.0], m), (ii, jj)), shape=(m, n)).tolil()
    if kind == "coupled":
        H = sp.diags([-np.ones(n - 1), 2.5 * np.ones(n), -np.ones(n - 1)], [-1, 0, 1], format="csc")
    else:
        H = sp.identity(n, format="csc")
        A[:, 0] = rng.standard_normal((m, 1)) + 2.0
    A = A.tocsc()
    K = sp.bmat([[H, A.T], [A, None]], format="csc")
    L = sp.tril(K, format="csc")
    L.sort_indices()
    return n + m, L.indptr.astype(np.int32), L.indices.astype(np.int32), L.data.copy(), K


@pytest.mark.parametrize("kind", ["coupled", "dense_column"])
def test_candidates_kept_in_the_reduced_system(kind):
    N, cp, ri, v, K = general_kkt(kind, 300, 150, seed=5)
    s = Symbolic(N, cp, ri, v)
    plan = s.plan()
    e_of_k = plan["e_of_k"]
    if kind == "coupled":
        # an independent set of the path graph: no two neighbours eliminated, and a maximal one
        el = e_of_k[:300] >= 0
        assert not (el[1:] & el[:-1]).any() and el.sum() == 150
    else:
        assert e_of_k[0] < 0 and (e_of_k[1:300] >= 0).all()  # the dense column stays, everything else goes
        assert s.stats()["nnz_S"] < 20 * len(ri)
    assert (e_of_k[300:] < 0).all()
    em = Emulated(plan, v)
    b = np.random.default_rng(1).standard_normal(N)
    # `coupled`: working-set rows without an eliminated neighbour get the static pivot -sqrt(eps) |S|_max when they
    # come before their variables; refinement against the unperturbed K removes it (numeric.cu: k_set_tau)
    z = em.solve(b, refine=3)
    assert np.linalg.norm(K @ z - b) <= 1e-9 * np.linalg.norm(b)
    assert np.linalg.norm(z - spla.spsolve(K, b)) <= 1e-7 * np.linalg.norm(z)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,m", [("coupled", 300, 150), ("dense_column", 300, 150), ("coupled", 20000, 8000), ("dense_column", 10010, 5000)])
def test_device_path_with_candidates_kept_in_the_reduced_system(kind, n, m):
    """a J with one dense column at n = 1e4 does not densify S; a non-diagonal (1,1) block is accepted."""
    from sleqp_b200 import Fact

    N, cp, ri, v, K = general_kkt(kind, n, m, seed=7)
    f = Fact()
    f.set_matrix(N, cp, ri, v)
    st = f.stats()
    assert st["n_demoted"] == (n // 2 if kind == "coupled" else (1 if m > 128 else 0))
    if n > 1000:
        assert st["nnz_L"] < 20 * len(ri), st  # no densification (eliminating the dense column: nnz(S) = m^2 / 2 = 1.25e7)
    rng = np.random.default_rng(3)
    for _ in range(3):
        b = rng.standard_normal(N)
        idx = np.arange(N, dtype=np.int32)
        f.solve(idx, b, N)
        x = f.solution_dense(0, N)
        assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b), st["probe_residual"]
    f.release()
