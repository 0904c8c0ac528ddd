"""SLEQP-iterate-sequence proxy: the reference's own aug_jac + Steihaug projected CG
(oracle/eqp_harness.c, unmodified reference code) over the B200 backend must reproduce what it computes over
the reference LAPACK backend -- projections, min-norm step, LSQ multipliers, the converged Newton step and
samples of the CG path -- to 1e-8 (north_star: "the same SLEQP iterate sequence to 1e-8")."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
N, RADII = 400, 12


def _run(exe):
    out = subprocess.run([os.path.join(REF, exe), str(N), str(RADII)], check=True, capture_output=True, text=True, timeout=600)
    data = {}
    for line in out.stdout.splitlines():
        parts = line.split()
        data[parts[0]] = np.array(parts[2:], dtype=np.float64)
    return data, out.stderr


def _compare(got, want, tol):
    assert set(want.keys()) <= set(got.keys())
    for k in want:
        scale = max(1.0, float(np.abs(want[k]).max()))
        assert got[k].shape == want[k].shape, k
        assert np.abs(got[k] - want[k]).max() <= tol * scale, (k, float(np.abs(got[k] - want[k]).max()), scale)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "eqp_harness_lapack")), reason="oracle/_ref not built")
def test_reference_harness_matches_committed_fixture(golden):
    want = golden(f"eqp_harness_lapack_n{N}.npz")
    got, err = _run("eqp_harness_lapack")
    assert "LAPACK" in err
    _compare(got, {k: want[k] for k in want.files}, 1e-9)
    # the samples really cover several CG segments (otherwise the comparison would be weak)
    a, b = got["cg_path_sample_0"], got[f"cg_path_sample_{RADII - 1}"]
    assert a @ b / np.linalg.norm(a) / np.linalg.norm(b) < 0.95


@pytest.mark.gpu
def test_b200_backend_reproduces_reference_eqp_iterates(golden):
    if not os.path.exists(os.path.join(REF, "eqp_harness_b200")):
        pytest.skip("oracle/_ref/eqp_harness_b200 not shipped")
    want = golden(f"eqp_harness_lapack_n{N}.npz")
    got, err = _run("eqp_harness_b200")
    assert "B200" in err
    _compare(got, {k: want[k] for k in want.files}, 1e-8)
    assert got["backend"][0] == 2  # SLEQP_FACT_FLAGS_LOWER
