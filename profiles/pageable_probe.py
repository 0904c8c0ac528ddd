import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from sleqp_b200 import Fact, problems
p = problems.config(1)
f = Fact(device=0)
f.set_matrix(p.N, *p.kkt_lower())
idx, val = p.rhs("project_nullspace", 1)
for rep in range(8):
    t0 = time.perf_counter(); f.solve(idx, val, p.N); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    x = f.solution_dense(0, p.n); t3 = time.perf_counter()
    print(rep, 'solve %.3f sync %.3f solution_dense %.3f ms' % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)))
for rep in range(4):
    t0 = time.perf_counter(); f.solve(idx, val, p.N); t1 = time.perf_counter()
    x = f.solution_dense(0, p.n); t3 = time.perf_counter()
    print(rep, 'nosync: solve %.3f solution_dense %.3f ms' % (1e3*(t1-t0), 1e3*(t3-t1)))
