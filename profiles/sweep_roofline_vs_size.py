"""Sweeps against the HBM roofline as the 2D Poisson-control problem grows (config 2 is g = 354): the kernels are the
same, the share of the level-to-level handoff in a sweep shrinks. Usage: python profiles/sweep_roofline_vs_size.py [g ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sleqp_b200 import Fact, problems

for g in [int(a) for a in sys.argv[1:]] or [177, 354, 708, 1000]:
    p = problems.poisson_control(g, 2)
    f = Fact(device=0)
    f.set_matrix(p.N, *p.kkt_lower())
    idx, val = p.rhs("project_nullspace", 1)
    f.solve(idx, val, p.N)
    x = f.solution_dense(0, p.N)
    b = np.zeros(p.N); b[idx] = val
    res = float(np.linalg.norm(p.kkt_full() @ x - b) / np.linalg.norm(b))
    st = f.stats()
    ph = f.profile_solve(10)
    byts = 8 * st["nnz_L"] + 4 * st["n_row_idx"] + 16 * st["n_reduced"]
    print(f"g={g} N={p.N} nnz_L={st['nnz_L']} levels={st['n_levels']} numeric={st['ms_numeric']:.2f} ms residual={res:.1e} | forward {ph[1]*1e3:.0f} us = {100*byts/ph[1]/1e6/6545.6:.1f} % | backward {ph[2]*1e3:.0f} us = {100*byts/ph[2]/1e6/6545.6:.1f} % of 6545.6 GB/s")
    f.release()
