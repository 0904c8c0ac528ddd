"""Cost of a working-set change: set_matrix with a new sparsity pattern on a warm handle (symbolic analysis on the host,
plan upload, graph capture, numeric) against a repeat of a known pattern. Config 2 geometry, two different active sets.
B200_TIMING=1 prints the host-side phases of every call on stderr."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sleqp_b200 import Fact, problems  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 354
mats = []
for seed in (0, 1, 2):
    p = problems.poisson_control(g, 2, seed=seed)
    cp, ri, v = p.kkt_lower()
    mats.append((p.N, cp, ri, torch.from_numpy(v).pin_memory().numpy()))
f = Fact(device=0)
for label, m in (("A cold (first call of the process)", 0), ("A again", 0), ("B new pattern", 1), ("B again", 1), ("A (known pattern)", 0),
                 ("C new pattern", 2), ("C again", 2)):
    N, cp, ri, v = mats[m]
    print(f"--- {label}", file=sys.stderr)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    f.set_matrix(N, cp, ri, v)
    ms = 1e3 * (time.perf_counter() - t0)
    st = f.stats()
    print(f"{label:38s} N={N} set_matrix {ms:8.2f} ms  (symbolic {st['ms_symbolic']:.1f} ms, numeric {st['ms_numeric']:.2f} ms, cached={st['symbolic_cached']})")
