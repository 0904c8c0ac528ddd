"""bench.py on the CPU: helpers behind the roofline numbers, the committed bench lines against the JSON contract, and
the requirement that the product arm refuses to run without a CUDA device (no CPU fallback)."""
import glob
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "e2e", "gpu_launches", "clocks", "roofline")


def test_committed_bench_lines_follow_the_contract():
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_final_bench_config[234]*.json"))
                   + glob.glob(os.path.join(ROOT, "profiles", "r02_bench_default_config3*.json")))
    assert len(files) >= 4
    for path in files:
        d = json.loads(open(path).read().strip().splitlines()[-1])
        for k in REQUIRED:
            assert k in d, (path, k)
        assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True
        assert d["vs_baseline"] is None  # BASELINE.md holds no published number for this metric
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["gpu_launches"] > 0
        assert abs(d["value"] - d["n_gpus"] * 1e3 / d["ms_per_step"]) <= 1e-6 * d["value"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-12
        # achieved = algorithmic bytes of one launch / its measured duration
        assert abs(r["achieved"] - r["bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert base["metric"]


def test_traffic_comes_from_the_committed_ncu_capture():
    for kernel in ("k_flow_fwd", "k_flow_bwd"):
        t = bench.ncu_traffic_per_launch(kernel, "config2_poisson2d_g354")
        assert t is not None and 1.2e8 < t < 2.5e8  # padded panels (136 MB) + records and indices
    # config 3 is swept by the sparse-subtree kernels: traffic ~ the exact factor (12 MB) + indices + vectors
    for kernel in ("k_sst_forward", "k_sst_backward"):
        t = bench.ncu_traffic_per_launch(kernel, "config3_chain_n1e6")
        assert t is not None and 1.5e7 < t < 3.5e7
    assert bench.ncu_traffic_per_launch("k_flow_fwd", "config3_chain_n1e6") is None  # no capture for that workload
    assert bench.ncu_traffic_per_launch("no_such_kernel", "config2_poisson2d_g354") is None


def test_peaks_have_a_stated_source():
    peaks, src = bench.load_peaks()
    assert peaks["hbm_gbs"] > 1000 and isinstance(src, str) and src


def test_step_workload_shapes():
    w = bench.make_workload(0)
    p = w["p"]
    rhs = bench.step_rhs(w, 5)
    assert len(rhs) == 7  # min-norm + least-squares solves, then one projection per CG iteration
    for kind, idx, val, b, e in rhs:
        assert len(idx) == len(val) and np.all(np.diff(idx) > 0) and 0 <= b < e <= p.N
    assert w["J"].shape == (p.m, p.n) and w["H"].shape == (p.n, p.n)


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--config", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CPU fallback" in (out.stdout + out.stderr)
