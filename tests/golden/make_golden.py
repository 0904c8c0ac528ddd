"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/
libsleqp_ref_lapack.so, built from /root/reference by oracle/build_ref.sh) on seeded inputs.

Run in the build container (the reference does not travel to the GPU box):
    bash oracle/build_ref.sh && python tests/golden/make_golden.py
The fixtures pin (a) oracle/sleqp_oracle.py and (b) the CUDA path against reference outputs.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_lib import RefLib  # noqa: E402
from sleqp_b200 import problems  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def spmv_cases(ref):
    out = {}
    rng = np.random.default_rng(1234)
    cases = []
    # the reference's own known-answer test (src/test/sparse/sleqp_sparse_matrix_test.c:12-56)
    A = sp.csc_matrix(np.array([[1.0, 0.0, 2.0], [0.0, 2.0, 3.0]]))
    cases.append(("known", A, np.array([0, 1, 2]), np.array([2.0, 4.0, 3.0]), np.array([0, 1]), np.array([1.0, -2.0])))
    for name, (m, n, dens) in dict(small=(7, 5, 0.4), ragged=(40, 60, 0.08), tall=(300, 20, 0.1), wide=(20, 300, 0.1), emptycols=(30, 30, 0.02)).items():
        A = sp.random(m, n, density=dens, random_state=np.random.RandomState(rng.integers(1 << 30)), format="csc")
        A.sort_indices()
        xs = np.sort(rng.choice(n, size=max(1, n // 2), replace=False))
        vs = np.sort(rng.choice(m, size=max(1, m // 3), replace=False))
        cases.append((name, A, xs, rng.standard_normal(len(xs)), vs, rng.standard_normal(len(vs))))
    # empty sparse vectors
    A = cases[1][1]
    cases.append(("emptyvec", A, np.zeros(0, dtype=int), np.zeros(0), np.zeros(0, dtype=int), np.zeros(0)))
    for name, A, xi, xv, vi, vv in cases:
        m, n = A.shape
        y = ref.mat_mult_vec(m, n, A.indptr, A.indices, A.data, xi, xv)
        ti, tv = ref.mat_mult_vec_trans(m, n, A.indptr, A.indices, A.data, vi, vv, 1e-12)
        out.update({
            f"{name}_shape": np.array([m, n]), f"{name}_colptr": A.indptr.astype(np.int32), f"{name}_rows": A.indices.astype(np.int32),
            f"{name}_data": A.data, f"{name}_xi": xi.astype(np.int32), f"{name}_xv": xv, f"{name}_vi": vi.astype(np.int32), f"{name}_vv": vv,
            f"{name}_y": y, f"{name}_ti": ti, f"{name}_tv": tv,
        })
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "spmv_reference.npz"), **out)
    print("spmv cases:", [c[0] for c in cases])


def fact_cases(ref):
    out = {}
    probs = {
        "config1": problems.config(0),
        "poisson2d_g8": problems.poisson_control(8, 2, seed=3),
        "poisson3d_g5": problems.poisson_control(5, 3, seed=4),
        "chain_n400": problems.chain_rosenbrock(400, active_fraction=0.25, seed=5),
        "chain_n40_noactive": problems.chain_rosenbrock(40, active_fraction=0.0, seed=6),
    }
    names = []
    for name, p in probs.items():
        cp, ri, v = p.kkt_lower()
        f = ref.fact()
        assert f.name() == "LAPACK" and f.flags() == 2
        f.set_matrix(p.N, cp, ri, v)
        out.update({f"{name}_n": np.array([p.n, p.ws_size]), f"{name}_colptr": cp, f"{name}_rows": ri, f"{name}_data": v})
        for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
            idx, val = p.rhs(kind, seed=11)
            # drop a few entries so the rhs is genuinely sparse
            keep = np.ones(len(idx), dtype=bool)
            keep[::7] = False
            idx, val = idx[keep], val[keep]
            begin, end = (p.n, p.N) if kind == "solve_lsq" else (0, p.n)
            f.solve(idx, val)
            si, sv = f.solution(begin, end, 1e-20)
            out.update({f"{name}_{kind}_idx": idx, f"{name}_{kind}_val": val, f"{name}_{kind}_range": np.array([begin, end]),
                        f"{name}_{kind}_si": si, f"{name}_{kind}_sv": sv})
        f.release()
        names.append(name)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "fact_reference_lapack.npz"), **out)
    print("fact cases:", names)


def vec_cases(ref):
    rng = np.random.default_rng(7)
    v = rng.standard_normal(50)
    v[::3] = 0.0
    v[5] = 1e-21
    v[6] = -1e-20
    v[7] = 2e-20
    out = {"values": v}
    for eps in (0.0, 1e-20, 1e-3):
        i, x = ref.vec_set_from_raw(v, eps)
        out[f"idx_{eps}"] = i
        out[f"val_{eps}"] = x
    np.savez_compressed(os.path.join(OUT, "vec_reference.npz"), **out)


if __name__ == "__main__":
    ref = RefLib("lapack")
    spmv_cases(ref)
    fact_cases(ref)
    vec_cases(ref)


def eqp_harness_case(n=400, radii=12, problem="chain"):
    """Output of the reference EQP harness (oracle/eqp_harness.c over the reference LAPACK backend). problem = "chain"
    (n variables) or "poisson" (2D Poisson control, n = grid size)."""
    import subprocess

    exe = os.path.join(ROOT, "oracle", "_ref", "eqp_harness_lapack")
    out = subprocess.run([exe, str(n), str(radii), problem], check=True, capture_output=True, text=True).stdout
    data = {}
    for line in out.splitlines():
        parts = line.split()
        data[parts[0]] = np.array(parts[2:], dtype=np.float64)
    name = f"eqp_harness_lapack_n{n}.npz" if problem == "chain" else f"eqp_harness_lapack_{problem}_g{n}.npz"
    np.savez_compressed(os.path.join(OUT, name), **data)
    print("eqp harness:", name, list(data))


if __name__ == "__main__":
    eqp_harness_case()
    eqp_harness_case(12, 12, "poisson")
