"""Timing of the result read-back variants (config 2 slice [0, n))."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sleqp_b200 import Fact, problems
from sleqp_b200._lib import lib

p = problems.config(1)
f = Fact(device=0)
f.set_matrix(p.N, *p.kkt_lower())
idx, val = p.rhs("project_nullspace", 1)
f.solve(idx, val, p.N)
torch.cuda.synchronize()
L = lib()
n = p.n
ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

def bench(name, vi, vv):
    nnz = C.c_int()
    a, b = vi.ctypes.data_as(ip), vv.ctypes.data_as(dp)
    for _ in range(3):
        L.b200_fact_solution_sparse(f._h, 0, n, 1e-20, a, b, C.byref(nnz))
    t0 = time.perf_counter()
    for _ in range(20):
        L.b200_fact_solution_sparse(f._h, 0, n, 1e-20, a, b, C.byref(nnz))
    print(name, "sparse %.4f ms" % (1e3 * (time.perf_counter() - t0) / 20), nnz.value)
    for _ in range(3):
        L.b200_fact_solution(f._h, 0, n, b)
    t0 = time.perf_counter()
    for _ in range(20):
        L.b200_fact_solution(f._h, 0, n, b)
    print(name, "dense  %.4f ms" % (1e3 * (time.perf_counter() - t0) / 20))

bench("pageable        ", np.empty(n, np.int32), np.empty(n, np.float64))
vi, vv = np.empty(n, np.int32), np.empty(n, np.float64)
print("pin rc", L.b200_host_pin(vv.ctypes.data_as(C.c_void_p), vv.nbytes), L.b200_host_pin(vi.ctypes.data_as(C.c_void_p), vi.nbytes))
bench("registered      ", vi, vv)
ti, tv = torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
bench("cudaHostAlloc'd ", ti.numpy(), tv.numpy())
