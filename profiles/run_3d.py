"""Config 4 (3D Poisson control) sizing run: factor + solve + residual + per-class timings."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Fact, problems  # noqa: E402

for g in [int(a) for a in sys.argv[1:]] or [48]:
    p = problems.poisson_control(g, 3)
    cp, ri, v = p.kkt_lower()
    f = Fact(device=0)
    t0 = time.perf_counter()
    f.set_matrix(p.N, cp, ri, v)
    t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    f.set_matrix(p.N, cp, ri, v)
    t_warm = time.perf_counter() - t0
    K = p.kkt_full()
    res = []
    for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
        idx, val = p.rhs(kind, 3)
        f.solve(idx, val, p.N)
        x = f.solution_dense(0, p.N)
        b = np.zeros(p.N)
        b[idx] = val
        res.append(float(np.linalg.norm(K @ x - b) / np.linalg.norm(b)))
    st = f.stats()
    prof = f.profile_numeric()
    ph = f.profile_solve(5)
    gf = st["flops_factor_stored"] / 1e9
    print(f"g={g} N={p.N} nnz_L={st['nnz_L']} max_front={st['max_front']} stages={st['n_stages']} cold={t_cold:.2f}s warm={t_warm*1e3:.1f}ms "
          f"numeric={st['ms_numeric']:.1f}ms ({gf / st['ms_numeric']:.1f} TF/s on {gf:.0f} GF) solve={st['ms_solve']:.2f}ms residuals={res} refine={st['refine_steps']}")
    print("  numeric by class (ms):", {k: round(x, 2) for k, x in prof.items()}, " update DMMA rate:", round(gf / max(prof['update'], 1e-9), 2), "TF/s")
    print("  solve phases (ms):", ph.round(3).tolist())
    # sweeps against the HBM roofline (algorithmic bytes as in bench.py: exact nnz(L), row indices, vectors)
    sweep_bytes = 8 * st["nnz_L"] + 4 * st["n_row_idx"] + 16 * st["n_reduced"]
    for name, ms in (("forward", ph[1]), ("backward", ph[2])):
        print(f"  {name} sweep: {sweep_bytes / 1e6:.0f} MB in {ms * 1e3:.0f} us = {sweep_bytes / ms / 1e6:.0f} GB/s = {100 * sweep_bytes / ms / 1e6 / 6545.6:.1f} % of the measured HBM peak (6545.6 GB/s)")
    f.release()
