"""Host symbolic analysis (sleqp_b200/csrc/symbolic.cpp) through the C-ABI, no GPU needed:
structural outputs vs an independent symbolic factorization, and the numeric plan executed by a
task-by-task numpy emulation vs a sparse LU of the same K."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import sleqp_oracle as orc
from oracle.multifrontal_emul import Emulated
from sleqp_b200 import B200Error, Symbolic, problems

CASES = {
    "config1": lambda: problems.config(0),
    "poisson2d_g8": lambda: problems.poisson_control(8, 2),
    "poisson2d_g24": lambda: problems.poisson_control(24, 2, seed=2),
    "poisson3d_g6": lambda: problems.poisson_control(6, 3),
    "chain_n2000": lambda: problems.chain_rosenbrock(2000, 0.1),
    "chain_n33_noactive": lambda: problems.chain_rosenbrock(33, 0.0),
    "no_constraints": lambda: _no_cons(),
    "poisson2d_g48_wide_supernodes": lambda: problems.poisson_control(48, 2, seed=5),  # supernodes wider than one outer block
    "poisson2d_g100_three_outer_blocks": lambda: problems.poisson_control(100, 2, seed=1),  # k = 402: far updates, K = 128 prologues
}


def _no_cons():
    p = problems.chain_rosenbrock(20, 0.0)
    p.active_cons = p.active_cons[:0]
    return p


@pytest.mark.parametrize("name", list(CASES))
def test_structure_matches_independent_symbolic(name):
    p = CASES[name]()
    cp, ri, v = p.kkt_lower()
    s = Symbolic(p.N, cp, ri, v)
    perm, parent, cc, sf = s.structure()
    assert sorted(perm.tolist()) == list(range(p.N))
    par_ref, cc_ref = orc.symbolic_reference(p.N, cp, ri, perm)
    assert np.array_equal(parent, par_ref)  # bit-exact: functions of pattern and permutation only
    assert np.array_equal(cc, cc_ref)
    st = s.stats()
    assert st["nnz_L"] == int(cc[st["n_elim"]:].sum())
    # every constraint is ordered after every variable it touches (SURVEY.md hard part 1)
    pinv = np.empty(p.N, dtype=np.int64)
    pinv[perm] = np.arange(p.N)
    for j in range(p.n):
        rows = ri[cp[j] + 1: cp[j + 1]]
        assert np.all(pinv[rows] > pinv[j])
    # supernodes partition the order, the etree is a forest with parent > child
    assert sf[0] == 0 and sf[-1] == p.N and np.all(np.diff(sf) > 0)
    nz = parent >= 0
    assert np.all(parent[nz] > np.nonzero(nz)[0])


@pytest.mark.parametrize("name", ["config1", "poisson2d_g24", "poisson3d_g6", "chain_n2000", "poisson2d_g48_wide_supernodes", "poisson2d_g100_three_outer_blocks"])
def test_plan_emulation_solves_kkt(name):
    p = CASES[name]()
    cp, ri, v = p.kkt_lower()
    s = Symbolic(p.N, cp, ri, v)
    em = Emulated(s.plan(), v)
    assert em.n_perturbed == 0
    K = p.kkt_full()
    for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
        idx, val = p.rhs(kind, 3)
        b = orc.vec_to_raw(idx, val, p.N)
        # the dataflow sweep tasks (what the device runs): list order respects every dependency counter
        z = em.solve(b, refine=0)
        assert np.linalg.norm(K @ z - b) <= 1e-12 * np.linalg.norm(b)
        zr = spla.spsolve(K.tocsc(), b)
        assert np.linalg.norm(z - zr) <= 1e-10 * np.linalg.norm(zr)
        # plain supernodal substitution with the factor itself (no inverse panels) as a cross-check
        zs = em.solve(b, refine=0, flow=False)
        assert np.linalg.norm(zs - z) <= 1e-10 * np.linalg.norm(z)


DEEP_CASES = {
    "poisson2d_g48_wide_supernodes": CASES["poisson2d_g48_wide_supernodes"],
    "poisson3d_g10": lambda: problems.poisson_control(10, 3, seed=4),
}


def _depths(plan):
    f = np.asarray(plan["ffl_tasks"]).reshape(-1, 16)
    b = np.asarray(plan["bfl_tasks"]).reshape(-1, 16)
    return f[:, 9] - f[:, 8], b[:, 7] - b[:, 6]


def test_gpu_deep_task_cases_have_deep_tasks(monkeypatch):
    """The problems of tests/test_gpu_parity.py::test_deep_sweep_tasks really produce depth > 32 tasks in both sweeps
    (only supernodes with at least 128 columns get them), and never without the lowered threshold at these sizes."""
    for make in (lambda: problems.poisson_control(44, 2, seed=11), lambda: problems.poisson_control(12, 3, seed=12),
                 lambda: problems.poisson_control(90, 2, seed=13)):
        p = make()
        cp, ri, v = p.kkt_lower()
        monkeypatch.delenv("B200_FLOW_DEEP_TASKS", raising=False)
        s = Symbolic(p.N, cp, ri, v)
        df, db = _depths(s.plan())
        assert df.max() <= 32 and db.max() <= 32
        s.close()
        monkeypatch.setenv("B200_FLOW_DEEP_TASKS", "1")
        s = Symbolic(p.N, cp, ri, v)
        df, db = _depths(s.plan())
        assert 64 <= df.max() <= 128 and 64 <= db.max() <= 128, p.name
        # a deep chunk is at least 64 deep; everything else keeps the shapes of the other levels
        assert not np.any((df > 32) & (df < 64)) and not np.any((db > 32) & (db < 64))
        s.close()


@pytest.mark.parametrize("name", list(DEEP_CASES))
def test_deep_sweep_tasks_emulate(name, monkeypatch):
    """The bandwidth-bound levels of large fronts get tasks of depth 128 (symbolic.cpp); the threshold is lowered so
    that small problems produce them too. Same counters, same result."""
    monkeypatch.setenv("B200_FLOW_DEEP_TASKS", "1")
    p = DEEP_CASES[name]()
    cp, ri, v = p.kkt_lower()
    s = Symbolic(p.N, cp, ri, v)
    plan = s.plan()
    f = np.asarray(plan["ffl_tasks"]).reshape(-1, 16)
    b = np.asarray(plan["bfl_tasks"]).reshape(-1, 16)
    assert (f[:, 9] - f[:, 8]).max() > 32 and (b[:, 7] - b[:, 6]).max() > 32
    assert (f[:, 9] - f[:, 8]).max() <= 128 and (b[:, 7] - b[:, 6]).max() <= 128
    em = Emulated(plan, v)
    K = p.kkt_full()
    idx, val = p.rhs("solve_lsq", 3)
    rhs = orc.vec_to_raw(idx, val, p.N)
    z = em.solve(rhs, refine=0)
    assert np.linalg.norm(K @ z - rhs) <= 1e-12 * np.linalg.norm(rhs)


@pytest.mark.parametrize("make", [lambda: problems.chain_rosenbrock(6000, 0.1, seed=3), lambda: problems.config(0), lambda: problems.chain_rosenbrock(300, 0.3, seed=8)],
                         ids=["chain_n6000", "config1", "chain_n300"])
def test_sparse_subtrees_on_chains(make, monkeypatch):
    """Chains / banded systems (configs 1 and 3) keep their bottom subtrees as sparse supernodes (plan.hpp SST, sst.cu):
    the stored factor is the exact one (dense 32-column blocks stored ~6 x as much), the tree above is a few levels, and
    the emulated sparse factorization + level-scheduled substitution solves K like the all-dense plan."""
    p = make()
    cp, ri, v = p.kkt_lower()
    s = Symbolic(p.N, cp, ri, v)
    st, plan = s.stats(), s.plan()
    assert len(plan["sst"]) >= 1 and plan["sn_sparse"].sum() == len(plan["sst"])
    assert st["nnz_L_stored"] <= 1.05 * st["nnz_L"] + 2048
    monkeypatch.setenv("B200_SST", "0")
    sd = Symbolic(p.N, cp, ri, v)
    std_, pland = sd.stats(), sd.plan()
    assert len(pland["sst"]) == 0 and std_["nnz_L_stored"] > 2 * st["nnz_L_stored"] and std_["n_levels"] >= st["n_levels"]
    assert std_["nnz_L"] == st["nnz_L"]  # same fill: only the representation differs
    K = p.kkt_full()
    sols = []
    for pl in (plan, pland):
        em = Emulated(pl, v)
        assert em.n_perturbed == 0
        idx, val = p.rhs("solve_lsq", 5)
        b = orc.vec_to_raw(idx, val, p.N)
        z = em.solve(b, refine=0)
        assert np.linalg.norm(K @ z - b) <= 1e-12 * np.linalg.norm(b)
        assert np.linalg.norm(em.solve(b, refine=0, flow=False) - z) <= 1e-10 * np.linalg.norm(z)
        sols.append(z)
    assert np.abs(sols[0] - sols[1]).max() <= 1e-10 * np.abs(sols[1]).max()
    # every level of a subtree only depends on the levels before it (checked inside Emulated._sst_factor)


def test_two_dimensional_problems_have_no_sparse_subtrees():
    p = problems.poisson_control(24, 2, seed=2)
    s = Symbolic(p.N, *p.kkt_lower())
    assert len(s.plan()["sst"]) == 0


def test_plan_does_not_depend_on_the_number_of_host_threads(monkeypatch):
    """Nested dissection, product-term search and assembly map run on several host threads; ordering, task lists and
    the order of the product terms inside every entry of S (the summation order on the device) must not depend on how many."""
    p = problems.poisson_control(60, 2, seed=9)
    cp, ri, v = p.kkt_lower()
    plans = []
    for nth in ("1", "3", "8"):
        monkeypatch.setenv("B200_HOST_THREADS", nth)
        s = Symbolic(p.N, cp, ri, v)
        plans.append((s.stats()["perm_hash"], s.plan()))
        s.close()
    for h, plan in plans[1:]:
        assert h == plans[0][0]
        for k, a in plans[0][1].items():
            if k != "ms_symbolic":
                assert np.array_equal(np.asarray(a), np.asarray(plan[k])), k


def test_same_pattern_same_structure_different_values():
    a = problems.chain_rosenbrock(300, 0.2, seed=1)
    b = problems.chain_rosenbrock(300, 0.2, seed=1)
    b.J.data[:] = np.random.default_rng(9).uniform(0.5, 1.5, size=len(b.J.data))
    sa = Symbolic(a.N, *a.kkt_lower())
    sb = Symbolic(b.N, *b.kkt_lower())
    assert sa.stats()["pattern_hash"] == sb.stats()["pattern_hash"]
    assert sa.stats()["perm_hash"] == sb.stats()["perm_hash"]
    c = problems.chain_rosenbrock(300, 0.2, seed=2)  # other active set => other pattern
    assert Symbolic(c.N, *c.kkt_lower()).stats()["pattern_hash"] != sa.stats()["pattern_hash"]


def test_full_matrix_input_equals_lower_input():
    p = problems.poisson_control(6, 2)
    cp, ri, v = p.kkt_lower()
    K = p.kkt_full()
    K.sort_indices()
    a = Symbolic(p.N, cp, ri, v, lower_only=True).structure()
    b = Symbolic(p.N, K.indptr, K.indices, K.data, lower_only=False).structure()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_malformed_input_is_rejected():
    p = problems.config(0)
    cp, ri, v = p.kkt_lower()
    bad = ri.copy()
    bad[1], bad[2] = bad[2], bad[1]  # rows not increasing in a column
    with pytest.raises(B200Error):
        Symbolic(p.N, cp, bad, v)
    K = p.kkt_full()
    K.sort_indices()
    with pytest.raises(B200Error):  # upper entries in a matrix declared lower
        Symbolic(p.N, K.indptr, K.indices, K.data, lower_only=True)
    # off-diagonal coupling inside the (1,1) block: accepted since round 2 -- the later of the two coupled candidates stays
    # in the reduced system (numerics: tests/test_random_structures.py)
    import scipy.sparse as sp

    K2 = sp.tril(K).tolil()
    K2[1, 0] = 0.5
    K2 = K2.tocsc()
    K2.sort_indices()
    s2 = Symbolic(p.N, K2.indptr, K2.indices, K2.data)
    assert s2.stats()["n_demoted"] == 1 and s2.export("e_of_k")[1] < 0 <= s2.export("e_of_k")[0]


def test_malformed_colptr_is_an_error_not_a_crash():
    """the pattern hash dereferenced colptr before anything validated it (segfault on a negative entry).
    b200_symbolic_analyze and the plan lookup of set_matrix now check the header in an O(n) pass first."""
    ri = np.array([0, 1, 2, 3], dtype=np.int32)
    v = np.ones(4)
    for cp in ([0, 2, -1000000000, 3, 4], [0, 3, 2, 3, 4], [0, 1, 2, 9, 4], [1, 1, 2, 3, 4]):
        with pytest.raises(B200Error) as e:
            Symbolic(4, np.array(cp, dtype=np.int32), ri, v)
        assert e.value.code == 1


def test_pattern_key_has_two_independent_hashes():
    """The plan cache is keyed by (N, nnz, hash, second hash): both change with the pattern, neither with the values."""
    a = problems.chain_rosenbrock(300, 0.2, seed=1)
    c = problems.chain_rosenbrock(300, 0.2, seed=2)
    sa, sc = Symbolic(a.N, *a.kkt_lower()).stats(), Symbolic(c.N, *c.kkt_lower()).stats()
    assert sa["pattern_hash"] != sc["pattern_hash"] and sa["pattern_hash2"] != sc["pattern_hash2"]
    assert sa["pattern_hash"] != sa["pattern_hash2"]


def test_empty_matrix():
    s = Symbolic(0, np.zeros(1, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0))
    assert s.stats()["n"] == 0
