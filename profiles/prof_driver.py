"""Minimal driver for ncu captures: one cold set_matrix (analysis + numeric graph + probe solve) on config 2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Fact, problems  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
p = problems.config(cfg)
f = Fact(device=0)
f.set_matrix(p.N, *p.kkt_lower())
idx, val = p.rhs("project_nullspace", 1)
f.solve(idx, val, p.N)
f.solution_dense(0, p.n)
print(f.stats())
print('numeric by class (ms):', f.profile_numeric())
print('solve phases (ms):', f.profile_solve(10))
f.release()
