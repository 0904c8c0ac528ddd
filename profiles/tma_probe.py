"""Small reproducer for the TMA update path (run under compute-sanitizer when it misbehaves)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200_TMA_MIN_FRONT", "1")
from sleqp_b200 import Fact, problems  # noqa: E402

p = problems.poisson_control(int(sys.argv[1]) if len(sys.argv) > 1 else 12, 2, seed=3)
f = Fact(device=0)
f.set_matrix(p.N, *p.kkt_lower())
K = p.kkt_full()
idx, val = p.rhs("project_nullspace", 1)
f.solve(idx, val, p.N)
x = f.solution_dense(0, p.N)
b = np.zeros(p.N)
b[idx] = val
print("residual", np.linalg.norm(K @ x - b) / np.linalg.norm(b), f.stats()["max_front"])
