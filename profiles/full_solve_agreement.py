"""How long the accepted iterates of a full sleqp_solver_solve agree between the B200 backend and the reference LAPACK
backend (oracle/_ref/full_solve_*): prints the number of leading iterates equal to 1e-8 and the final objectives."""
import os
import subprocess
import sys

import numpy as np

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def run(exe, args, env=None):
    out = subprocess.run([os.path.join(REF, exe), *args], check=True, capture_output=True, text=True, env=dict(os.environ, **(env or {})))
    d = {}
    for line in out.stdout.splitlines():
        p = line.split()
        d[p[0]] = np.array(p[2:], dtype=float)
    return d


for args in (["hs71"], ["chain", "100", "200"], ["chain", "40", "200"]):
    a, b = run("full_solve_b200", args), run("full_solve_lapack", args)
    k = 0
    while f"iterate_{k}" in a and f"iterate_{k}" in b and np.abs(a[f"iterate_{k}"] - b[f"iterate_{k}"]).max() <= 1e-8 * max(1.0, np.abs(b[f"iterate_{k}"]).max()):
        k += 1
    print(args, "leading iterates equal to 1e-8:", k, "| iterations", int(a["iterations"][0]), int(b["iterations"][0]), "| status", int(a["status"][0]), int(b["status"][0]),
          "| objective", a["objective"][0], b["objective"][0], "| ms", a["elapsed_ms"][0], b["elapsed_ms"][0],
          "| |x_b200 - x_lapack|_inf", float(np.abs(a["solution"] - b["solution"]).max()))
