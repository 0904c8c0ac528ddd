"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel class on problems with
ragged tiles, several panel steps, wide supernodes (two-level blocking) and all three right-hand-side kinds."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Fact, Mat, problems  # noqa: E402

for p in (problems.config(0), problems.poisson_control(30, 2, seed=1), problems.poisson_control(7, 3, seed=2), problems.chain_rosenbrock(700, 0.3, seed=3)):
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
    m = Mat(device=0)
    J = p.J
    m.set(J.shape[0], J.shape[1], J.indptr, J.indices, J.data)
    m.mult_vec(np.arange(0, p.n, 3, dtype=np.int32), np.ones(len(range(0, p.n, 3))))
    m.mult_vec_trans(np.arange(p.m, dtype=np.int32), np.ones(p.m))
    print(p.name, "ok", f.stats()["n_stages"])
    m.release()
    f.release()
