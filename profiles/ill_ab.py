import sys; sys.path.insert(0,'/root/repo')
import numpy as np, scipy.sparse as sp
from sleqp_b200 import problems, Fact
from oracle import sleqp_oracle as orc
for eps in (1e-3,1e-5):
    base = problems.chain_rosenbrock(400, 0.1, seed=3)
    rng = np.random.default_rng(1)
    r10 = base.J.tocsr()[10].toarray().ravel()
    new = r10 + eps * rng.standard_normal(base.n) * (r10 != 0)
    J2 = sp.vstack([base.J.tocsr(), sp.csr_matrix(new)]).tocsc(); J2.sort_indices()
    p = problems.KKTProblem(name="ill", n=base.n, m=base.m + 1, J=J2, H=base.H, active_vars=base.active_vars, active_cons=np.arange(base.m + 1))
    f = Fact(); f.set_matrix(p.N, *p.kkt_lower())
    st=f.stats(); print(eps, 'refine', st['refine_steps'], 'probe', st['probe_residual'], 'cond', f.cond())
    K = p.kkt_full()
    for rep in range(3):
      for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
        idx, val = p.rhs(kind, 2); f.solve(idx, val, p.N); x = f.solution_dense(0, p.N); b = orc.vec_to_raw(idx, val, p.N)
        print('  ', kind, np.linalg.norm(K @ x - b)/np.linalg.norm(b))
