"""Parity of the CUDA path (through the C-ABI) with the oracle. Bars (BASELINE.json north_star):
||K x - b|| / ||b|| <= 1e-10 against the unperturbed K for all three aug_jac right-hand sides,
solution slices vs the oracle to 1e-8 relative, structure bit-identical to the host analysis,
SpMV / SpMV^T to 1e-12 relative (FP64, different summation order)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import sleqp_oracle as orc
from oracle.multifrontal_emul import Emulated
from sleqp_b200 import B200Error, Fact, Mat, Symbolic, problems

pytestmark = pytest.mark.gpu

KINDS = ("project_nullspace", "solve_min_norm", "solve_lsq")
RES_TOL = 1e-10   # north_star: relative residual
SOL_TOL = 1e-8    # north_star: iterate / solution agreement


def _check_problem(p, f=None, seeds=(1,)):
    cp, ri, v = p.kkt_lower()
    f = f or Fact()
    f.set_matrix(p.N, cp, ri, v)
    K = p.kkt_full()
    lu = orc.SparseLU()
    lu.set_matrix(p.N, cp, ri, v)
    for seed in seeds:
        for kind in KINDS:
            idx, val = p.rhs(kind, seed)
            begin, end = (p.n, p.N) if kind == "solve_lsq" else (0, p.n)
            f.solve(idx, val, p.N)
            x = f.solution_dense(0, p.N)
            b = orc.vec_to_raw(idx, val, p.N)
            res = np.linalg.norm(K @ x - b) / np.linalg.norm(b)
            assert res <= RES_TOL, (p.name, kind, res)
            lu.solve(idx, val)
            ref = lu.sol[begin:end]
            assert np.abs(x[begin:end] - ref).max() <= SOL_TOL * max(1.0, np.abs(ref).max()), (p.name, kind)
            si, sv = f.solution(begin, end, 1e-20)
            oi, ov = orc.vec_set_from_raw(x[begin:end], 1e-20)
            assert np.array_equal(si, oi) and np.array_equal(sv, ov)
            # a threshold that drops about half of the entries (the compaction path behind the device-side count)
            eps2 = float(np.median(np.abs(x[begin:end])))
            si, sv = f.solution(begin, end, eps2)
            oi, ov = orc.vec_set_from_raw(x[begin:end], eps2)
            assert np.array_equal(si, oi) and np.array_equal(sv, ov)
            if kind == "project_nullspace":
                A = p.working_rows()
                assert np.abs(A @ x[: p.n]).max() <= 1e-10 * np.abs(val).max()
    return f


def test_reference_lapack_fixtures(golden):
    g = golden("fact_reference_lapack.npz")
    f = Fact()
    for name in g["names"]:
        n, ws = g[f"{name}_n"]
        N = int(n + ws)
        f.set_matrix(N, g[f"{name}_colptr"], g[f"{name}_rows"], g[f"{name}_data"])
        for kind in KINDS:
            b, e = g[f"{name}_{kind}_range"]
            f.solve(g[f"{name}_{kind}_idx"], g[f"{name}_{kind}_val"], N)
            si, sv = f.solution(int(b), int(e), 1e-20)
            got = orc.vec_to_raw(si, sv, e - b)
            ref = orc.vec_to_raw(g[f"{name}_{kind}_si"], g[f"{name}_{kind}_sv"], e - b)
            assert np.abs(got - ref).max() <= SOL_TOL * max(1.0, np.abs(ref).max()), (name, kind)


def test_known_answers():
    # constrained_newton_test.c:204-275 (projection (2,4) -> (2,0)) and dual_estimation_test.c:84-85
    f = Fact()
    cp, ri, v = orc.fill_aug_jac(2, [0, 0, 1], [0], [1.0], np.array([-1, -1]), np.array([0]), 1)
    f.set_matrix(3, cp, ri, v)
    f.solve([0, 1], [2.0, 4.0], 3)
    si, sv = f.solution(0, 2)
    assert si.tolist() == [0] and abs(sv[0] - 2.0) <= 1e-8
    vi, ci, ws = orc.working_set_indices(2, 0, [0, 1], [])
    cp, ri, v = orc.fill_aug_jac(2, [0, 0, 0], [], [], vi, ci, ws)
    f.set_matrix(4, cp, ri, v)
    f.solve([0, 1], [-2.0, -4.0], 4)
    assert np.allclose(f.solution_dense(2, 4), [-2.0, -4.0], atol=1e-8)


@pytest.mark.parametrize("make", [
    lambda: problems.config(0),
    lambda: problems.poisson_control(8, 2),
    lambda: problems.poisson_control(40, 2, seed=3),
    lambda: problems.poisson_control(10, 3, seed=4),
    lambda: problems.chain_rosenbrock(5000, 0.15, seed=5),
    lambda: problems.chain_rosenbrock(33, 0.0),
    lambda: problems.poisson_control(100, 2, seed=1),  # supernodes of 400 columns: three outer blocks of the panel steps
], ids=["config1", "p2d_g8", "p2d_g40", "p3d_g10", "chain_5000", "chain_33", "p2d_g100"])
def test_small_and_medium_problems(make):
    _check_problem(make(), seeds=(1, 2))


@pytest.mark.parametrize("make", [
    lambda: problems.poisson_control(44, 2, seed=11),
    lambda: problems.poisson_control(12, 3, seed=12),
    lambda: problems.poisson_control(90, 2, seed=13),
], ids=["p2d_g44", "p3d_g12", "p2d_g90"])
def test_deep_sweep_tasks(make, monkeypatch):
    """Sweep tasks of depth 128 (rolling window of loads, solve.cu flow_task_deep): the threshold that reserves them
    for the bandwidth-bound levels of large 3D fronts is lowered so that these small fronts use them."""
    monkeypatch.setenv("B200_FLOW_DEEP_TASKS", "1")
    p = make()
    f = _check_problem(p, seeds=(1, 2))
    assert f.stats()["refine_steps"] == 0


@pytest.mark.parametrize("make", [
    lambda: problems.poisson_control(40, 2, seed=3),
    lambda: problems.poisson_control(10, 3, seed=4),
    lambda: problems.poisson_control(100, 2, seed=1),
    lambda: problems.chain_rosenbrock(5000, 0.15, seed=5),
], ids=["p2d_g40", "p3d_g10", "p2d_g100", "chain_5000"])
def test_update_tiles_through_tma(make, monkeypatch):
    """The DMMA update tiles of large fronts fetch their operands with cp.async.bulk.tensor through per-supernode
    tensor maps (numeric.cu tile_update_tma: 128-byte swizzle, permuted fragment rows). The threshold is lowered so
    that every front of these small problems takes that path: same pivots as the cp.async path, same solutions."""
    p = make()
    f_ref = Fact()
    f_ref.set_matrix(p.N, *p.kkt_lower())
    d_ref = f_ref.pivots()
    f_ref.release()
    monkeypatch.setenv("B200_TMA_MIN_FRONT", "1")
    f = _check_problem(p, seeds=(1, 2))
    d = f.pivots()
    assert np.abs(d - d_ref).max() <= 1e-11 * np.abs(d_ref).max()
    f.release()


@pytest.mark.parametrize("make", [
    lambda: problems.chain_rosenbrock(50_000, 0.1, seed=6),
    lambda: problems.config(0),
    lambda: problems.chain_rosenbrock(2500, 0.3, seed=7),
], ids=["chain_5e4", "config1", "chain_2500"])
def test_sparse_subtrees_equal_the_dense_path(make, monkeypatch):
    """sst.cu (sparse LDL^T and level-scheduled substitution of the chain-like bottom of the tree, one CTA per subtree)
    against the all-dense supernodal path on the same matrices: same pivots, same solutions, all parity gates."""
    p = make()
    f = _check_problem(p, seeds=(1, 2))
    st = f.stats()
    assert st["nnz_L_stored"] <= 1.05 * st["nnz_L"] + 2048  # the sparse path really is in use
    d = f.pivots()
    monkeypatch.setenv("B200_SST", "0")
    g = _check_problem(p, seeds=(1,))
    assert g.stats()["nnz_L_stored"] > 2 * st["nnz_L_stored"]
    dg = g.pivots()
    # same elimination order up to the labelling inside the subtrees: compare the pivots as multisets per magnitude
    assert np.abs(np.sort(d) - np.sort(dg)).max() <= 1e-10 * np.abs(dg).max()
    idx, val = p.rhs("solve_lsq", 9)
    f.solve(idx, val, p.N)
    g.solve(idx, val, p.N)
    xa, xb = f.solution_dense(0, p.N), g.solution_dense(0, p.N)
    assert np.abs(xa - xb).max() <= 1e-9 * max(1.0, np.abs(xb).max())
    f.release()
    g.release()


def test_structure_and_pivots_match_host_analysis_and_emulation():
    p = problems.poisson_control(20, 2, seed=7)
    cp, ri, v = p.kkt_lower()
    f = Fact()
    f.set_matrix(p.N, cp, ri, v)
    s = Symbolic(p.N, cp, ri, v)
    for a, b in zip(f.structure(), s.structure()):
        assert np.array_equal(a, b)  # bit-identical structure
    assert f.stats()["perm_hash"] == s.stats()["perm_hash"]
    em = Emulated(s.plan(), v)
    d = f.pivots()
    ref = em.pivots_full()
    assert np.abs(d - ref).max() <= 1e-11 * np.abs(ref).max()
    assert abs(f.cond() - 1.0 / orc.rcond_from_pivots(ref)) <= 1e-8 * f.cond()


def test_symbolic_cache_and_pattern_change():
    f = Fact()
    a = problems.chain_rosenbrock(3000, 0.2, seed=1)
    _check_problem(a, f)
    b = problems.chain_rosenbrock(3000, 0.2, seed=1)
    b.J.data[:] = np.random.default_rng(3).uniform(0.5, 1.5, size=len(b.J.data))
    _check_problem(b, f)
    assert f.stats()["symbolic_cached"] == 1  # same pattern, new values: analysis reused
    c = problems.chain_rosenbrock(3000, 0.3, seed=2)  # working set changed: new pattern
    _check_problem(c, f)
    d = problems.poisson_control(12, 2)  # different size on the same handle
    _check_problem(d, f)
    _check_problem(a, f)
    assert f.stats()["symbolic_cached"] == 1


def test_edge_cases():
    f = Fact()
    # no working set at all: K = I
    p = problems.chain_rosenbrock(50, 0.0)
    p.active_cons = p.active_cons[:0]
    cp, ri, v = p.kkt_lower()
    f.set_matrix(p.N, cp, ri, v)
    idx, val = p.rhs("project_nullspace", 1)
    f.solve(idx, val, p.N)
    assert np.array_equal(f.solution_dense(0, p.N), val)
    # empty right-hand side -> zero solution, empty sparse slice
    q = problems.config(0)
    f.set_matrix(q.N, *q.kkt_lower())
    f.solve([], [], q.N)
    si, sv = f.solution(0, q.n)
    assert len(si) == 0
    # bounds only (no general constraints active)
    r = problems.chain_rosenbrock(64, 0.5, seed=3)
    r.active_cons = r.active_cons[:0]
    _check_problem(r, f)
    # protocol and argument errors
    g = Fact()
    with pytest.raises(B200Error):
        g.solve([0], [1.0], 3)
    with pytest.raises(B200Error):
        f.solve([0], [1.0], q.N + 1)
    with pytest.raises(B200Error):
        f.solution_dense(0, 10 * q.N)


def test_singular_working_set_raises():
    # two identical working-set rows: K is singular -> error, like Umfpack (fact_umfpack.c:66-82)
    n = 6
    A = sp.csr_matrix(np.array([[1.0, 2.0, 0, 0, 0, 0], [1.0, 2.0, 0, 0, 0, 0], [0, 0, 1.0, 0, 1.0, 0]]))
    K = sp.tril(sp.bmat([[sp.identity(n), A.T], [A, None]])).tocsc()
    K.sort_indices()
    # keep empty trailing columns like the reference layout
    f = Fact()
    with pytest.raises(B200Error) as e:
        f.set_matrix(n + 3, K.indptr, K.indices, K.data)
    assert e.value.code == 3


def test_scaled_diagonal_and_full_input():
    # D != I and the full symmetric matrix as input (flags NONE layout)
    p = problems.poisson_control(9, 2, seed=5)
    K = p.kkt_full().tolil()
    rng = np.random.default_rng(0)
    dscale = rng.uniform(0.5, 4.0, size=p.n)
    for j in range(p.n):
        K[j, j] = dscale[j]
    K = K.tocsc()
    K.sort_indices()
    f = Fact()
    f.set_matrix(p.N, K.indptr, K.indices, K.data, lower_only=False)
    b = rng.standard_normal(p.N)
    f.solve(np.arange(p.N), b, p.N)
    x = f.solution_dense(0, p.N)
    assert np.linalg.norm(K @ x - b) <= RES_TOL * np.linalg.norm(b)


def test_spmv_reference_fixtures(golden):
    g = golden("spmv_reference.npz")
    m_ = Mat()
    for name in g["names"]:
        m, n = (int(x) for x in g[f"{name}_shape"])
        m_.set(m, n, g[f"{name}_colptr"], g[f"{name}_rows"], g[f"{name}_data"])
        y = m_.mult_vec(g[f"{name}_xi"], g[f"{name}_xv"])
        ref = g[f"{name}_y"]
        assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), name
        ti, tv = m_.mult_vec_trans(g[f"{name}_vi"], g[f"{name}_vv"], eps=1e-12)
        got = orc.vec_to_raw(ti, tv, n)
        want = orc.vec_to_raw(g[f"{name}_ti"], g[f"{name}_tv"], n)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), name


def test_spmv_on_workloads():
    rng = np.random.default_rng(2)
    for p in (problems.poisson_control(64, 2), problems.chain_rosenbrock(100_000), problems.poisson_control(12, 3)):
        for A in (p.J, p.H):
            A = A.tocsc()
            A.sort_indices()
            m_ = Mat()
            m_.set(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
            xi = np.sort(rng.choice(A.shape[1], size=A.shape[1] // 2, replace=False)).astype(np.int32)
            xv = rng.standard_normal(len(xi))
            y = m_.mult_vec(xi, xv)
            ref = orc.mat_mult_vec(A.shape[0], A.indptr, A.indices, A.data, xi, xv)
            assert np.abs(y - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
            vi = np.arange(A.shape[0], dtype=np.int32)
            vv = rng.standard_normal(len(vi))
            ti, tv = m_.mult_vec_trans(vi, vv, eps=0.0)
            oi, ov = orc.mat_mult_vec_trans(A.shape[1], A.indptr, A.indices, A.data, vi, vv, A.shape[0], 0.0)
            assert np.abs(orc.vec_to_raw(ti, tv, A.shape[1]) - orc.vec_to_raw(oi, ov, A.shape[1])).max() <= 1e-12 * max(1.0, np.abs(ov).max())
            # the same product sparsified on the device (what the SpMV glue hands back as a SleqpVec)
            si, sv = m_.mult_vec_trans_sparse(vi[::50], vv[::50], eps=1e-14)
            oi2, ov2 = orc.mat_mult_vec_trans(A.shape[1], A.indptr, A.indices, A.data, vi[::50], vv[::50], A.shape[0], 1e-14)
            assert np.array_equal(si, oi2) and np.abs(sv - ov2).max() <= 1e-12 * max(1.0, np.abs(ov2).max())
            # linearity: A(2x) = 2 A x exactly in binary floating point
            assert np.array_equal(m_.mult_vec(xi, 2.0 * xv), 2.0 * y)


LARGE = {
    "config3_chain_n1e6": lambda: problems.config(2),
    "config2_poisson2d_g354": lambda: problems.config(1),
    "config4_poisson3d_g32": lambda: problems.poisson_control(32, 3, name="config4_poisson3d_g32"),
    # the sizes config 4 is benchmarked at: fronts of 7-13 k rows, depth-128 sweep tasks
    "config4_poisson3d_g48": lambda: problems.poisson_control(48, 3, name="config4_poisson3d_g48"),
    "config4_poisson3d_g64": lambda: problems.poisson_control(64, 3, name="config4_poisson3d_g64"),
}


@pytest.mark.parametrize("name", list(LARGE))
def test_large_configs_by_residual(name):
    """BASELINE.json sizes, checked through size-independent properties: residual of the
    unperturbed K, feasibility of the projection (A_W P r = 0), idempotence P(P r) = P r."""
    p = LARGE[name]()
    cp, ri, v = p.kkt_lower()
    f = Fact()
    f.set_matrix(p.N, cp, ri, v)
    K = p.kkt_full()
    A = p.working_rows()
    for kind in KINDS:
        idx, val = p.rhs(kind, 4)
        f.solve(idx, val, p.N)
        x = f.solution_dense(0, p.N)
        b = orc.vec_to_raw(idx, val, p.N)
        assert np.linalg.norm(K @ x - b) <= RES_TOL * np.linalg.norm(b), (p.name, kind)
    idx, val = p.rhs("project_nullspace", 5)
    f.solve(idx, val, p.N)
    pr = f.solution_dense(0, p.n)
    assert np.abs(A @ pr).max() <= 1e-10 * np.abs(val).max()
    f.solve(idx, pr, p.N)
    ppr = f.solution_dense(0, p.n)
    assert np.abs(ppr - pr).max() <= 1e-10 * np.abs(pr).max()
    st = f.stats()
    if "g48" in name or "g64" in name:
        assert st["max_front"] >= 7000 and st["refine_steps"] == 0
        plan_depth = st["n_levels"]
        assert plan_depth >= 8
    print(p.name, {k: st[k] for k in ("n", "nnz_L", "nnz_L_stored", "n_supernodes", "n_levels", "n_stages", "max_front", "ms_symbolic", "ms_numeric", "ms_solve", "refine_steps", "probe_residual")})
    f.release()


def test_config5_batch_of_64_concurrent_handles():
    """Config 5 as bench.py --batch 64 drives it: 64 independent g=128 instances on one device,
    one handle and CUDA stream each, refactorizations and solves of all instances interleaved from one host thread so
    that the sweeps of different handles overlap on the GPU. Every instance is checked by its own residual."""
    import torch

    from bench import make_workload, step_rhs

    dev = torch.device("cuda", 0)
    inst = []
    for i in range(64):
        w = make_workload(4, seed=i)
        p = w["p"]
        f = Fact(device=0)
        f.set_matrix(p.N, w["cp"], w["ri"], w["v"])
        rhs = []
        for kind, idx, val, b, e in step_rhs(w, 2)[:4]:
            rhs.append(torch.from_numpy(orc.vec_to_raw(idx, val, p.N)).to(dev))
        inst.append(dict(f=f, p=p, d_val=torch.from_numpy(w["v"]).to(dev), rhs=rhs, sol=[torch.empty(p.N, dtype=torch.float64, device=dev) for _ in rhs]))
    assert len({it["f"].stats()["pattern_hash"] for it in inst}) > 1  # the seeds move the active bounds: several patterns
    torch.cuda.synchronize()
    for rep in range(2):
        for it in inst:
            it["f"].refactor_device(it["d_val"].data_ptr())
        for s_ in range(4):
            for it in inst:
                it["f"].solve_device(it["rhs"][s_].data_ptr(), it["sol"][s_].data_ptr())
    torch.cuda.synchronize()
    worst = 0.0
    for it in inst:
        K = it["p"].kkt_full()
        for b, x in zip(it["rhs"], it["sol"]):
            b, x = b.cpu().numpy(), x.cpu().numpy()
            res = np.linalg.norm(K @ x - b) / np.linalg.norm(b)
            worst = max(worst, res)
            assert res <= RES_TOL, (it["p"].name, res)
    print("config5 x64: worst relative residual", worst)
    for it in inst:
        it["f"].release()


def test_two_handles_on_two_devices_in_one_process():
    """the opt-in to > 48 KB of dynamic shared memory and the SM count were taken from the first device
    only. Needs two visible GPUs (skipped on the single-GPU test box; the multi-GPU bench exercises one device per
    process)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two devices")
    p = problems.poisson_control(40, 2, seed=3)
    for dev in (0, 1):
        f = Fact(device=dev)
        _check_problem(p, f)
        f.release()


def test_drop_in_through_reference_code():
    """The reference's own SleqpFact wrapper (fact.c) driving fact_b200.c: same results as the
    reference LAPACK backend on the same calls."""
    from oracle import ref_lib

    if not (ref_lib.available("b200") and ref_lib.available("lapack")):
        pytest.skip("oracle/_ref libraries not shipped")
    ours, theirs = ref_lib.RefLib("b200").fact(), ref_lib.RefLib("lapack").fact()
    assert ours.name() == "B200" and ours.flags() == 2 and theirs.name() == "LAPACK"
    for p in (problems.config(0), problems.poisson_control(10, 2, seed=8), problems.chain_rosenbrock(600, 0.2, seed=4)):
        cp, ri, v = p.kkt_lower()
        ours.set_matrix(p.N, cp, ri, v)
        theirs.set_matrix(p.N, cp, ri, v)
        assert np.isfinite(ours.cond()) and ours.cond() >= 1.0
        for kind in KINDS:
            idx, val = p.rhs(kind, 6)
            begin, end = (p.n, p.N) if kind == "solve_lsq" else (0, p.n)
            ours.solve(idx, val)
            theirs.solve(idx, val)
            a = orc.vec_to_raw(*ours.solution(begin, end), end - begin)
            b = orc.vec_to_raw(*theirs.solution(begin, end), end - begin)
            assert np.abs(a - b).max() <= SOL_TOL * max(1.0, np.abs(b).max()), (p.name, kind)
    ours.release()
    theirs.release()


@pytest.mark.parametrize("eps", [1e-3, 1e-5, 1e-6, 1e-8])
def test_nearly_dependent_rows_use_refinement(eps):
    """A working-set row that is a 1 + eps copy of another one: the Schur pivots lose ~eps^2 of relative accuracy,
    iterative refinement against the unperturbed K restores the 1e-10 residual."""
    base = problems.chain_rosenbrock(400, 0.1, seed=3)
    rng = np.random.default_rng(1)
    r10 = base.J.tocsr()[10].toarray().ravel()
    new = r10 + eps * rng.standard_normal(base.n) * (r10 != 0)
    J2 = sp.vstack([base.J.tocsr(), sp.csr_matrix(new)]).tocsc()
    J2.sort_indices()
    p = problems.KKTProblem(name="ill", n=base.n, m=base.m + 1, J=J2, H=base.H, active_vars=base.active_vars, active_cons=np.arange(base.m + 1))
    f = Fact()
    if eps < 1e-6:
        # the limit of the reduced form: S = A A^T squares the condition number, a pivot of relative size eps^2 = 1e-16
        # cannot be told from an exactly dependent row in double precision -- reported loudly as singular (what
        # fact_umfpack.c:66-82 does with UMFPACK_WARNING_singular_matrix for dependent rows), never a wrong solution
        with pytest.raises(B200Error, match="singular"):
            f.set_matrix(p.N, *p.kkt_lower())
        return
    f.set_matrix(p.N, *p.kkt_lower())
    assert f.stats()["refine_steps"] >= 1
    assert f.cond() > 1e6
    K = p.kkt_full()
    for kind in KINDS:
        idx, val = p.rhs(kind, 2)
        f.solve(idx, val, p.N)
        x = f.solution_dense(0, p.N)
        b = orc.vec_to_raw(idx, val, p.N)
        # ||x|| reaches 2e9 ||b|| for the right-hand sides that excite the nearly dependent pair, so the residual of ANY
        # backward-stable solver is bounded by eps_mach ||K|| ||x|| there, not by 1e-10 ||b|| (SuperLU with partial
        # pivoting: 3e-7 ||b|| on solve_min_norm): the gate adds 64 ulps of normwise backward error
        slack = 64 * np.finfo(float).eps * spla.norm(K, np.inf) * np.linalg.norm(x)
        assert np.linalg.norm(K @ x - b) <= RES_TOL * np.linalg.norm(b) + slack, kind
