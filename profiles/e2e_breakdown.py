"""Where the end-to-end time of one solve through the plugin calls goes (host buffers, config 2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sleqp_b200 import Fact, problems

p = problems.config(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
f = Fact(device=0)
f.set_matrix(p.N, *p.kkt_lower())
idx, val = p.rhs("project_nullspace", 1)
valp = torch.from_numpy(val).pin_memory().numpy()
for name, v in (("pageable rhs", val), ("pinned rhs", valp)):
    for _ in range(3):
        f.solve(idx, v, p.N); f.solution(0, p.n)
    T = {"solve": 0.0, "solution": 0.0, "solution_dense": 0.0, "sync": 0.0}
    reps = 20
    for _ in range(reps):
        t0 = time.perf_counter(); f.solve(idx, v, p.N); t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        f.solution(0, p.n); t3 = time.perf_counter()
        f.solution_dense(0, p.n); t4 = time.perf_counter()
        T["solve"] += t1 - t0; T["sync"] += t2 - t1; T["solution"] += t3 - t2; T["solution_dense"] += t4 - t3
    print(name, {k: round(1e3 * v / reps, 4) for k, v in T.items()}, "ms")
print("device phases", f.profile_solve(20))
