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
