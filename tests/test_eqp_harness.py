"""SLEQP-iterate-sequence proxy: the reference's own aug_jac + Steihaug projected CG
(oracle/eqp_harness.c, unmodified reference code) over the B200 backend must reproduce what it computes over
the reference LAPACK backend -- projections, min-norm step, LSQ multipliers, the converged Newton step and
samples of the CG path -- to 1e-8 (north_star: "the same SLEQP iterate sequence to 1e-8")."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
N, RADII = 400, 12
# (harness arguments, committed fixture of the reference LAPACK backend, problem builder in sleqp_b200.problems)
CASES = {
    "chain_n400": ((str(N), str(RADII), "chain"), f"eqp_harness_lapack_n{N}.npz", "eqp_harness_problem", N),
    # second case: a Poisson-control problem, working set by hand like constrained_newton_test.c:196-201
    "poisson_g12": (("12", str(RADII), "poisson"), "eqp_harness_lapack_poisson_g12.npz", "eqp_harness_poisson", 12),
}


def _run(exe, case="chain_n400"):
    out = subprocess.run([os.path.join(REF, exe), *CASES[case][0]], check=True, capture_output=True, text=True, timeout=600)
    data = {}
    for line in out.stdout.splitlines():
        parts = line.split()
        data[parts[0]] = np.array(parts[2:], dtype=np.float64)
    return data, out.stderr


def _compare(got, want, tol):
    assert set(want.keys()) <= set(got.keys())
    for k in want:
        scale = max(1.0, float(np.abs(want[k]).max()))
        assert got[k].shape == want[k].shape, k
        assert np.abs(got[k] - want[k]).max() <= tol * scale, (k, float(np.abs(got[k] - want[k]).max()), scale)


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "eqp_harness_lapack")), reason="oracle/_ref not built")
def test_reference_harness_matches_committed_fixture(golden, case):
    want = golden(CASES[case][1])
    got, err = _run("eqp_harness_lapack", case)
    assert "LAPACK" in err
    _compare(got, {k: want[k] for k in want.files}, 1e-9)
    # the samples really cover several CG segments (otherwise the comparison would be weak)
    a, b = got["cg_path_sample_0"], got[f"cg_path_sample_{RADII - 1}"]
    assert a @ b / np.linalg.norm(a) / np.linalg.norm(b) < 0.95


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_b200_backend_reproduces_reference_eqp_iterates(golden, case):
    if not os.path.exists(os.path.join(REF, "eqp_harness_b200")):
        pytest.skip("oracle/_ref/eqp_harness_b200 not shipped")
    want = golden(CASES[case][1])
    got, err = _run("eqp_harness_b200", case)
    assert "B200" in err
    _compare(got, {k: want[k] for k in want.files}, 1e-8)
    assert got["backend"][0] == 2  # SLEQP_FACT_FLAGS_LOWER


@pytest.mark.gpu
@pytest.mark.parametrize("exe", ["eqp_harness_b200tr", "eqp_harness_b200aj"])
@pytest.mark.parametrize("hess", ["callback", "matrix"])
@pytest.mark.parametrize("case", list(CASES))
def test_b200_tr_solver_plugin_reproduces_reference_steihaug(golden, case, hess, exe):
    """host/tr/tr_b200.c registered through the reference's own tr_solver.c (sleqp_tr_solver_create /
    _solve / _current_rayleigh) over fact_b200.c, against what the reference's Steihaug solver computed over the
    reference LAPACK backend: steps, the dual of the trust region and the Rayleigh bounds, to 1e-8. Both Hessian modes:
    the reference's matrix-free callback (sleqp_problem_hess_prod on the host) and a device matrix. Second executable:
    the same with host/aug_jac/b200_aug_jac.c (KKT assembled on the device from the Jacobian and the working set, SURVEY
    section 8f rank 2) in place of the reference's standard_aug_jac.c."""
    if not os.path.exists(os.path.join(REF, exe)):
        pytest.skip(f"oracle/_ref/{exe} not shipped")
    want = golden(CASES[case][1])
    out = subprocess.run([os.path.join(REF, exe), *CASES[case][0], hess], check=True, capture_output=True, text=True, timeout=600)
    got = {}
    for line in out.stdout.splitlines():
        parts = line.split()
        got[parts[0]] = np.array(parts[2:], dtype=np.float64)
    assert "B200" in out.stderr
    _compare(got, {k: want[k] for k in want.files}, 1e-8)
    assert any(k.startswith("tr_info_") for k in want.files)


def _check_tr_info(info, want, tol=1e-8):
    """(tr_dual, min_rayleigh, max_rayleigh) as sleqp_tr_solver_solve / _current_rayleigh returned them."""
    got = np.array([info["tr_dual"], info["min_rayleigh"], info["max_rayleigh"]])
    assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), (got, want)


def _harness_setup(case="chain_n400"):
    from sleqp_b200 import problems

    p, grad = getattr(problems, CASES[case][2])(CASES[case][3])
    return p, grad


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_cg_restatement_matches_reference_fixture(golden, case):
    """oracle.steihaug_projected_cg (numpy restatement of steihaug_solver.c) on the harness problem, with the
    projection done by a sparse LU of K, reproduces what the reference's own Steihaug solver printed."""
    from oracle import sleqp_oracle as orc

    want = golden(CASES[case][1])
    p, grad = _harness_setup(case)
    cp, ri, v = p.kkt_lower()
    lu = orc.SparseLU()
    lu.set_matrix(p.N, cp, ri, v)
    H = p.H.tocsr()

    def project(r):
        return lu.solve(np.arange(p.n), r)[: p.n].copy()

    # the projection and the objective gradient themselves
    assert np.abs(project(grad) - want["project_nullspace"]).max() <= 1e-9 * np.abs(want["project_nullspace"]).max()
    info = {}
    step, it, how = orc.steihaug_projected_cg(project, lambda d: H @ d, grad, 1e8, 1e-4, 4 * N, info)
    assert how == "interior" and it > 10
    assert np.abs(step - want["cg_converged_step"]).max() <= 1e-8 * max(1.0, np.abs(want["cg_converged_step"]).max())
    _check_tr_info(info, want["tr_info_converged"])
    assert want["tr_info_converged"][0] == -1.0  # SLEQP_NONE: no dual of the trust region on an interior exit
    full = np.linalg.norm(want["cg_converged_step"])
    for i in range(RADII):
        radius = full * (0.6 + 0.4 * (i + 0.5) / RADII)
        s, _, how = orc.steihaug_projected_cg(project, lambda d: H @ d, grad, radius, 1e-4, 4 * N, info)
        assert how == "boundary"
        assert np.abs(s - want[f"cg_path_sample_{i}"]).max() <= 1e-8 * max(1.0, np.abs(want[f"cg_path_sample_{i}"]).max())
        _check_tr_info(info, want[f"tr_info_{i}"])
        assert want[f"tr_info_{i}"][0] >= 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_device_resident_projected_cg_matches_reference_iterates(golden, case):
    """the whole projected-CG loop on the device (b200_cg_solve) against the iterates
    of the reference's Steihaug solver (fixture from oracle/eqp_harness.c over the reference LAPACK backend)."""
    from sleqp_b200 import Fact, Mat, ProjectedCG

    want = golden(CASES[case][1])
    p, grad = _harness_setup(case)
    f = Fact()
    f.set_matrix(p.N, *p.kkt_lower())
    H = p.H.tocsc()
    H.sort_indices()
    mh = Mat()
    mh.set(p.n, p.n, H.indptr, H.indices, H.data)
    cg = ProjectedCG(f, mh)
    gi = np.arange(p.n, dtype=np.int32)
    step, it, how, dual, rmin, rmax = cg.solve_ex(p.n, gi, grad, 1e8, 1e-4, 4 * N)
    assert how == ProjectedCG.INTERIOR and it > 10
    ref = want["cg_converged_step"]
    assert np.abs(step - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())
    assert np.isnan(dual)  # not computed on an interior exit (the glue leaves SLEQP_NONE)
    _check_tr_info(dict(tr_dual=-1.0, min_rayleigh=rmin, max_rayleigh=rmax), want["tr_info_converged"])
    full = np.linalg.norm(ref)
    for i in range(RADII):
        radius = full * (0.6 + 0.4 * (i + 0.5) / RADII)
        s, _, how, dual, rmin, rmax = cg.solve_ex(p.n, gi, grad, radius, 1e-4, 4 * N)
        assert how == ProjectedCG.BOUNDARY
        ref = want[f"cg_path_sample_{i}"]
        assert np.abs(s - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max()), i
        _check_tr_info(dict(tr_dual=dual, min_rayleigh=rmin, max_rayleigh=rmax), want[f"tr_info_{i}"])
    # sparse result (what tr_b200.c hands to the caller's SleqpVec): same step, sparsified with the rule of
    # sleqp_vec_set_from_raw, through pageable arrays (host pass) and through page-locked ones (device compaction)
    eps = float(np.median(np.abs(step)))
    for pinned in (False, True):
        si, sv, it_s, how_s, _, _, _ = cg.solve_sparse(p.n, gi, grad, 1e8, 1e-4, 4 * N, zero_eps=eps, pinned=pinned)
        keep = np.nonzero(np.abs(step) > eps)[0]
        assert it_s == it and how_s == ProjectedCG.INTERIOR and 0 < len(keep) < p.n
        assert np.array_equal(si, keep) and np.abs(sv - step[keep]).max() <= 1e-8 * np.abs(step).max()
    # a sparse gradient (indices not contiguous: they are uploaded; a contiguous run keeps them on the host) equals the same
    # vector passed densely with explicit zeros
    g2 = grad.copy()
    g2[::3] = 0.0
    nz = np.nonzero(g2)[0].astype(np.int32)
    sa, ita, howa = cg.solve(p.n, nz, g2[nz], 1e8, 1e-4, 4 * N)
    sb, itb, howb = cg.solve(p.n, gi, g2, 1e8, 1e-4, 4 * N)
    assert ita == itb and howa == howb and np.abs(sa - sb).max() <= 1e-8 * max(1.0, np.abs(sb).max())
    # matrix-free Hessian (the reference's callback): same iterates with the products done on the host
    Hs = p.H.tocsr()
    cgf = ProjectedCG(f, None, hess_prod=lambda d: Hs @ d)
    s2, it2, how2, _, rmin2, rmax2 = cgf.solve_ex(p.n, gi, grad, 1e8, 1e-4, 4 * N)
    assert how2 == ProjectedCG.INTERIOR and it2 == it
    assert np.abs(s2 - step).max() <= 1e-8 * max(1.0, np.abs(step).max())
    cgf.release()
    # iteration cap: zero step, like the reference
    s, it, how = cg.solve(p.n, gi, grad, 1e8, 1e-4, 3)
    assert how == ProjectedCG.MAX_ITER and it == 3 and not s.any()
    # feasibility of the step: A_W p = 0
    assert np.abs(p.working_rows() @ step).max() <= 1e-9
    cg.release()
