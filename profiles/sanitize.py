"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel class on problems with
ragged tiles, several panel steps, wide supernodes (two-level blocking) and all three right-hand-side kinds."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Fact, Mat, problems  # noqa: E402

from sleqp_b200 import ProjectedCG  # noqa: E402

# chain_rosenbrock(6000): several sparse subtrees of two generations (sst.cu: ticketed sweeps, child assembly);
# poisson 2D / 3D: dense supernodes, TMA off (fronts below the threshold) -- B200_TMA_MIN_FRONT=1 in the environment turns it on
for p in (problems.config(0), problems.poisson_control(30, 2, seed=1), problems.poisson_control(7, 3, seed=2), problems.chain_rosenbrock(700, 0.3, seed=3),
          problems.chain_rosenbrock(6000, 0.1, seed=4)):
    f = Fact(device=0)
    f.set_matrix(p.N, *p.kkt_lower())
    K = p.kkt_full()
    for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
        idx, val = p.rhs(kind, 1)
        f.solve(idx[::2], val[::2], p.N)
        x = f.solution_dense(0, p.N)
        si, sv = f.solution(0, p.N, float(np.median(np.abs(x))))  # device-side sparsification (page-locked buffers)
        assert len(si) == int((np.abs(x) > float(np.median(np.abs(x)))).sum())
        b = np.zeros(p.N)
        b[idx[::2]] = val[::2]
        assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b)
    # device KKT assembly on the same handle (set_kkt twice: the second one takes the speculative path)
    Jc = p.J.tocsc()
    Jc.sort_indices()
    vi = np.full(p.n, -1, dtype=np.int32)
    vi[p.active_vars] = np.arange(len(p.active_vars), dtype=np.int32)
    ci = np.full(p.m, -1, dtype=np.int32)
    ci[p.active_cons] = len(p.active_vars) + np.arange(len(p.active_cons), dtype=np.int32)
    for _ in range(2):
        f.set_kkt(p.n, p.m, Jc.indptr, Jc.indices, Jc.data, vi, ci, len(p.active_vars) + len(p.active_cons))
    idx, val = p.rhs("project_nullspace", 2)
    f.solve(idx, val, p.N)
    x = f.solution_dense(0, p.N)
    b = np.zeros(p.N)
    b[idx] = val
    assert np.linalg.norm(K @ x - b) <= 1e-10 * np.linalg.norm(b)
    # device-controlled projected CG (fused reductions, last-block sums)
    H = p.H.tocsc()
    H.sort_indices()
    mh = Mat(device=0)
    mh.set(p.n, p.n, H.indptr, H.indices, H.data)
    cg = ProjectedCG(f, mh)
    g = np.random.default_rng(5).standard_normal(p.n)
    cg.solve_sparse(p.n, np.arange(p.n, dtype=np.int32), g, 1e8, 1e-6, 12, zero_eps=1e-12, pinned=True)
    cg.solve_ex(p.n, np.arange(p.n, dtype=np.int32), g, 0.5, 1e-6, 12)
    cg.release()
    mh.release()
    m = Mat(device=0)
    J = p.J
    m.set(J.shape[0], J.shape[1], J.indptr, J.indices, J.data)
    m.mult_vec(np.arange(0, p.n, 3, dtype=np.int32), np.ones(len(range(0, p.n, 3))))
    m.mult_vec_trans(np.arange(p.m, dtype=np.int32), np.ones(p.m))
    print(p.name, "ok", f.stats()["n_stages"])
    m.release()
    f.release()
