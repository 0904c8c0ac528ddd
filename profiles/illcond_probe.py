import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo')
from sleqp_b200 import Fact, problems, B200Error
base = problems.chain_rosenbrock(400, 0.1, seed=3)
for eps in (1e-3, 1e-5, 1e-7, 1e-9, 1e-11, 0.0):
    J = base.J.tolil()
    rng = np.random.default_rng(1)
    # make constraint row 11 nearly equal to row 10 (same pattern positions shifted): add a new near-duplicate row
    r10 = base.J.tocsr()[10].toarray().ravel()
    new = r10 + eps * rng.standard_normal(base.n) * (r10 != 0)
    J2 = sp.vstack([base.J.tocsr(), sp.csr_matrix(new)]).tocsc(); J2.sort_indices()
    p = problems.KKTProblem(name='ill', n=base.n, m=base.m + 1, J=J2, H=base.H, active_vars=base.active_vars, active_cons=np.arange(base.m + 1))
    f = Fact()
    try:
        f.set_matrix(p.N, *p.kkt_lower())
        K = p.kkt_full()
        idx, val = p.rhs('project_nullspace', 2)
        f.solve(idx, val, p.N)
        x = f.solution_dense(0, p.N)
        b = np.zeros(p.N); b[idx] = val
        st = f.stats()
        print(eps, 'res', np.linalg.norm(K @ x - b) / np.linalg.norm(b), 'refine', st['refine_steps'], 'probe', st['probe_residual'], 'nper', st['n_perturbed'], 'cond', f.cond())
    except B200Error as e:
        print(eps, 'ERROR', e)
    f.release()
