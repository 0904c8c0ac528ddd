#!/usr/bin/env python
"""Benchmark of the KKT hot path (BASELINE.json): one "step" = the EQP inner loop of one SQP
iteration on a fixed synthetic KKT system (unit iii):

    1 numeric factorization of [I A_W^T; A_W 0] (symbolic analysis cached)
  + (2 + k) solves   (1 min-norm, 1 LSQ multipliers, k null-space projections of the CG loop)
  + k Hessian SpMV, 1 Jacobian SpMV^T (sparse multipliers, as newton.c:377 passes them), 1 Jacobian SpMV

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4] [--cg-iters k]

Default workload: config 3 (chained Rosenbrock, n = 1e6), the configuration BASELINE.json's >=10x target is quoted
on. At N = 1 the same JSON line carries sub-lines for config 2 (2D Poisson control) and config 4 (3D Poisson control
at the largest grid that fits, DMMA-bound) under "configs"; under torchrun (N > 1) it carries a config-5 leg (64
independent instances sharded i mod N) instead.

`value` (SLEQP EQP iterations per second) is timed with CUDA events on the handle's stream with every input resident
in HBM; `e2e` is the same step through the reference-facing calls with HOST buffers, copies inside the timed region
(e2e.via says which boundary: the C-ABI mirror, and -- when oracle/_ref holds the reference-driven harness -- the
reference's own aug_jac / TR-solver code over fact_b200.c and tr_b200.c).
`--impl reference` times the reference's CPU path for the same step on the host cores: the reference's own SpMV
(oracle/_ref, compiled from the unmodified sources) and, for the factorization arithmetic that lives in absent
SuiteSparse, SciPy SuperLU as the stand-in (two variants: unsymmetric LU of K like Umfpack, and an LU without pivoting
of the SPD reduced matrix A_W A_W^T like host CHOLMOD on the sparse reduced form).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sleqp_eqp_iterations_per_s"
UNIT = "iter/s"
DEFAULT_CONFIG = 2  # index into BASELINE.json "configs": chained Rosenbrock n = 1e6 (the headline config)
ROUND = "r02"


def load_peaks():
    """HBM: driver-measured copy bandwidth. FP64: cuBLAS DGEMM 8192^3 measured on this pool (profiles/fp64_peak.py);
    MEASURED_PEAKS.json has no FP64 entry."""
    peaks, src = {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks, src = json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_dgemm_peak.json")) as f:
            peaks["fp64_tflops"] = json.load(f)["fp64_tflops"]
    except Exception:
        peaks["fp64_tflops"] = 35.5
    return peaks, src


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line), read through NVML every
    5 ms by a thread (what `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints; the CLI
    block-buffers its output into a pipe and takes longer to start than a short timed region lasts). mark() opens /
    closes the window of the timed region; only samples inside it count."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, device):
        self.device = device
        self.rows = []  # (time, sm_mhz, reasons bitmask)
        self.window = [None, None]
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is a plain list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.device
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.device])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def run():
                while not self.stop_flag:
                    try:
                        try:
                            mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(mask)))
                    except Exception:
                        pass
                    time.sleep(0.005)

            self.thread = threading.Thread(target=run, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def mark(self, which):
        self.window[which] = time.perf_counter()

    def samples_in_window(self):
        t0, t1 = self.window[0], self.window[1] or time.perf_counter()
        return sum(1 for t, _, _ in self.rows if t0 is None or t0 <= t <= t1)

    def stop(self):
        if not self.thread:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        self.stop_flag = True
        self.thread.join(timeout=2)
        t0, t1 = self.window
        rows = [r for r in self.rows if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)]
        reasons = sorted(name for name, bit in self.REASONS.items() if any(r[2] & bit for r in rows))
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


def ncu_traffic_per_launch(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full`
    capture of the same workload (profiles/<round>_ncu_full_<kernel>_<config>.csv); None if there is no capture."""
    import csv

    tag = next((t for t in ("config2", "config3", "config4") if t in workload), None)
    if tag is None:
        return None
    for rnd in (ROUND, "r01"):
        path = os.path.join(ROOT, "profiles", f"{rnd}_ncu_full_{kernel}_{tag}.csv")
        try:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = 0.0
            for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(name)
                tot += sum(float(r[i]) for r in rows[2:]) * scale[units[i]]
            return tot / len(rows[2:])
        except Exception:
            continue
    return None


def pick_3d_grid(free_bytes):
    """Config 4: the nominal g = 159 needs ~0.8 TB of factor, so the grid is the largest one whose
    three panel copies (L, Mt, Mr), update workspace, inversion scratch and vectors fit the free HBM with headroom.
    Panel entries measured: 1.76e9 at g = 96, 2.59e9 at g = 104 (profiles/r01_run_3d_g48_to_g96.txt, r02_run_3d_g104.txt),
    growing like g^4.1..4.8; about 3.6 panel copies' worth of memory in total."""
    for g, entries in ((104, 2.59e9), (96, 1.76e9), (88, 1.25e9), (80, 0.84e9), (64, 0.352e9), (48, 0.104e9)):
        if 1.05 * 8 * 3.6 * entries <= 0.55 * free_bytes:
            return g
    return 32


def make_workload(cfg_idx, seed=0, **kw):
    from sleqp_b200 import problems

    p = problems.config(cfg_idx, seed=seed, **kw)
    cp, ri, v = p.kkt_lower()
    J = p.J.tocsc()
    J.sort_indices()
    H = p.H.tocsc()
    H.sort_indices()
    rng = np.random.default_rng(100 + seed)
    # the multipliers of the violated constraints (newton.c:377 hands a SPARSE vector to sleqp_mat_mult_vec_trans):
    # one constraint in a hundred
    vi = np.arange(0, p.m, 100, dtype=np.int32)
    return dict(p=p, cp=cp, ri=ri, v=v, J=J, H=H, rng=rng, viol_idx=vi, viol_val=rng.standard_normal(len(vi)))


def step_rhs(w, k):
    """Right-hand sides of one step: (kind, idx, val, begin, end)."""
    p = w["p"]
    out = []
    kinds = ["solve_min_norm", "solve_lsq"] + ["project_nullspace"] * k
    for i, kind in enumerate(kinds):
        idx, val = p.rhs(kind, seed=1000 + i)
        begin, end = (p.n, p.N) if kind == "solve_lsq" else (0, p.n)
        out.append((kind, idx, val, begin, end))
    return out


# ------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_step_sample(w, rhs, k, n_solve_sample=4, with_spd=True):
    """One bounded CPU sample of the step (oracle / reference code; never the thing shipped). Returns a dict of
    seconds per part. Sequential code: SuperLU and the reference's SpMV use one core whatever the box has."""
    from oracle import ref_lib
    from oracle import sleqp_oracle as orc

    p = w["p"]
    ref = ref_lib.RefLib("lapack") if ref_lib.available("lapack") else None
    J, H = w["J"], w["H"]
    out = {}
    t0 = time.perf_counter()
    lu = orc.SparseLU()
    lu.set_matrix(p.N, w["cp"], w["ri"], w["v"])
    out["factor_s"] = time.perf_counter() - t0
    ns = min(n_solve_sample, len(rhs))
    t0 = time.perf_counter()
    for kind, idx, val, b, e in rhs[:ns]:
        lu.solve(idx, val)
        lu.solution(b, e, 1e-20)
    out["solve_s"] = (time.perf_counter() - t0) / ns
    xi = np.arange(p.n, dtype=np.int32)
    xv = np.ones(p.n)

    def timed_spmv(A, idx, val, trans):
        """Only the product is timed: the SleqpMat / SleqpVec containers are built before (SLEQP holds them already)."""
        if ref is None:
            t0 = time.perf_counter()
            if trans:
                orc.mat_mult_vec_trans(A.shape[1], A.indptr, A.indices, A.data, idx, val, A.shape[0], 0.0)
            else:
                orc.mat_mult_vec(A.shape[0], A.indptr, A.indices, A.data, idx, val)
            return time.perf_counter() - t0
        import ctypes as C

        m_ = ref.mat(A.shape[0], A.shape[1], A.indptr, A.indices, A.data)
        v_ = ref.vec(A.shape[0] if trans else A.shape[1], idx, val)
        if trans:
            r_ = C.POINTER(ref_lib.SleqpVec)()
            ref.call(ref.L.sleqp_vec_create_empty(C.byref(r_), int(A.shape[1])))
            t0 = time.perf_counter()
            ref.call(ref.L.sleqp_mat_mult_vec_trans(m_, v_, 0.0, r_))
            dt = time.perf_counter() - t0
            ref.L.sleqp_vec_free(C.byref(r_))
        else:
            out_ = np.empty(A.shape[0])
            t0 = time.perf_counter()
            ref.call(ref.L.sleqp_mat_mult_vec(m_, v_, out_.ctypes.data_as(C.POINTER(C.c_double))))
            dt = time.perf_counter() - t0
        ref.L.sleqp_vec_free(C.byref(v_))
        ref.L.sleqp_mat_release(C.byref(m_))
        return dt

    out["hess_spmv_s"] = timed_spmv(H, xi, xv, False)
    out["jt_spmv_s"] = timed_spmv(J, w["viol_idx"], w["viol_val"], True)
    out["j_spmv_s"] = timed_spmv(J, xi, xv, False)
    # The J^T product is reported but NOT part of the CPU step: the reference's two-pointer merge restarts at the head of
    # the multiplier vector for every column (mat.c:329-331), so its cost is ~ num_cols x nnz(multipliers) / 2 and
    # depends on how many constraints happen to be violated, not on the KKT path this benchmark measures.
    out["step_s"] = out["factor_s"] + out["solve_s"] * len(rhs) + out["hess_spmv_s"] * k + out["j_spmv_s"]
    out["spmv_code"] = "reference sleqp_mat_mult_vec(_trans) from oracle/_ref" if ref is not None else "numpy port"
    if with_spd:
        # the stand-in closest to "host CHOLMOD": the SPD reduced matrix S = A_W A_W^T (sparse, NOT the reference's
        # dense builder reduced_aug_jac.c:291) factored without pivoting under a minimum-degree ordering of S + S^T;
        # one reduced solve = two triangular sweeps
        import scipy.sparse.linalg as spla

        A = p.working_rows().tocsc()
        t0 = time.perf_counter()
        S = (A @ A.T).tocsc()
        out["spd_form_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        slu = spla.splu(S, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
        out["spd_factor_s"] = time.perf_counter() - t0
        b = np.ones(S.shape[0])
        t0 = time.perf_counter()
        for _ in range(ns):
            slu.solve(b)
        out["spd_solve_s"] = (time.perf_counter() - t0) / ns
    return out


CPU_NOTE = ("factor/solve = SciPy SuperLU (COLAMD, partial pivoting) on the full K: stand-in for the absent Umfpack, sequential; "
            "spd_* = SuperLU without pivoting (MMD on S+S^T, SymmetricMode) on the sparse SPD S = A_W A_W^T: stand-in for host CHOLMOD on a sparse "
            "reduced form (the reference's own reduced form is dense, reduced_aug_jac.c:291, and infeasible at this size)")


def run_reference(args, rank, world):
    """CPU arm: the same step on the host cores, bounded sample per step."""
    if rank != 0:
        return
    w = make_workload(args.config)
    p, k = w["p"], args.cg_iters
    rhs = step_rhs(w, k)
    budget_s = 150.0
    t_probe = time.perf_counter()
    first = cpu_step_sample(w, rhs, k, with_spd=True)
    t_probe = time.perf_counter() - t_probe
    # keep the whole run within a few minutes whatever --steps/--warmup are: one step holds a full sparse LU on the
    # host (seconds); the first step is timed to size the rest
    n_warm = max(0, min(args.warmup - 1, int(0.2 * budget_s / max(t_probe, 1e-3))))
    for _ in range(n_warm):
        cpu_step_sample(w, rhs, k, with_spd=False)
    n_meas = max(1, min(args.steps, int(0.8 * budget_s / max(t_probe, 1e-3))))
    samples = [first] if args.warmup == 0 else []
    while len(samples) < n_meas:
        samples.append(cpu_step_sample(w, rhs, k, with_spd=False))
    mean = {key: float(np.mean([s[key] for s in samples])) for key in ("factor_s", "solve_s", "hess_spmv_s", "jt_spmv_s", "j_spmv_s", "step_s")}
    ms = 1e3 * mean["step_s"]
    value = 1e3 / ms
    sample = (f"per step: 1 SuperLU factorization of K + 4 of {len(rhs)} solves (scaled x{len(rhs)}/4) + 1 of {k} Hessian SpMV (scaled x{k}) + 1 J SpMV (the J^T "
              f"product with sparse multipliers, newton.c:377, is timed and reported as jt_spmv_ms but excluded from the step: its cost in the reference is quadratic, mat.c:329-331); SpMV = {first['spmv_code']}; {CPU_NOTE}; {len(samples)} measured steps after {n_warm + (1 if args.warmup else 0)} "
              f"warm-up steps (self-limited to a {budget_s:.0f} s budget)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(samples), "warmup": (n_warm + 1 if args.warmup else 0),
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["p"].name, "N": p.N, "n": p.n, "ws_size": p.ws_size, "nnz_K": int(len(w["ri"])), "cg_iters": k, "solves_per_step": len(rhs)},
        "factor_ms": 1e3 * mean["factor_s"], "solve_ms": 1e3 * mean["solve_s"],
        "hess_spmv_ms": 1e3 * mean["hess_spmv_s"], "jt_spmv_ms": 1e3 * mean["jt_spmv_s"], "j_spmv_ms": 1e3 * mean["j_spmv_s"],
        "spd_stand_in": {"form_ms": 1e3 * first.get("spd_form_s", float("nan")), "factor_ms": 1e3 * first.get("spd_factor_s", float("nan")),
                         "solve_ms": 1e3 * first.get("spd_solve_s", float("nan")),
                         "step_ms": 1e3 * (first.get("spd_form_s", 0) + first.get("spd_factor_s", 0) + first.get("spd_solve_s", 0) * len(rhs) + mean["hess_spmv_s"] * k
                                           + mean["j_spmv_s"] * (1 + 2 * len(rhs))),
                         "note": "reduced solve = 2 triangular sweeps on S plus one product with A_W and one with A_W^T (counted as 2 J SpMV per solve)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "host_cores_available": host_cores(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def reference_driver_e2e(kind, size, k, steps, device):
    """End to end through 100 % reference-side code: oracle/_ref/eqp_step_b200 is the reference's own problem /
    iterate / working set / standard_aug_jac / TR-solver code (unmodified, compiled by oracle/build_ref.sh) linked with
    the shipped host glue fact_b200.c + tr_b200.c. It times set_iterate (host fill_aug_jac + set_matrix), the min-norm
    and LSQ solves, J^T / J products and the trust-region solve with host vectors. Returns its JSON or None."""
    exe = os.path.join(ROOT, "oracle", "_ref", "eqp_step_b200")
    if not os.path.exists(exe):
        return None
    env = dict(os.environ, B200_DEVICE=str(device))
    try:
        out = subprocess.run([exe, kind, str(size), str(k), str(steps)], capture_output=True, text=True, timeout=900, env=env)
        for line in out.stdout.splitlines():
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def measure_config(cfg_idx, args, torch, dev, local_rank, rank, world, dist, steps, main, cfg_kw=None):
    """Device-resident step, per-kernel-class rooflines and the end-to-end legs of one workload."""
    from sleqp_b200 import Fact, Mat, ProjectedCG, _lib, shard

    cfg_kw = cfg_kw or {}
    w = make_workload(cfg_idx, seed=rank if world > 1 else 0, **cfg_kw)
    p, k = w["p"], args.cg_iters
    rhs = step_rhs(w, k)
    J, H = w["J"], w["H"]
    lib = _lib.lib()
    peaks, peak_src = load_peaks()

    fact = Fact(device=local_rank)
    mJ, mH = Mat(device=local_rank), Mat(device=local_rank)
    t0 = time.perf_counter()
    fact.set_matrix(p.N, w["cp"], w["ri"], w["v"])  # cold: symbolic + numeric
    cold_ms = 1e3 * (time.perf_counter() - t0)
    st0 = fact.stats()
    mJ.set(J.shape[0], J.shape[1], J.indptr, J.indices, J.data)
    mH.set(H.shape[0], H.shape[1], H.indptr, H.indices, H.data)
    stream = torch.cuda.ExternalStream(fact.stream, device=dev)
    mJ.set_stream(fact.stream)
    mH.set_stream(fact.stream)

    # ---- device-resident inputs -------------------------------------------------------------------
    d_val = torch.from_numpy(w["v"]).to(dev)
    n_distinct = min(len(rhs), 8 if p.N > 2_000_000 else len(rhs))  # the largest 3D grids cycle 8 right-hand sides
    d_rhs = []
    for kind, idx, val, b, e in rhs[:n_distinct]:
        full = np.zeros(p.N)
        full[idx] = val
        d_rhs.append(torch.from_numpy(full).to(dev))
    d_sol = torch.empty(p.N, dtype=torch.float64, device=dev)
    d_x = torch.from_numpy(w["rng"].standard_normal(p.n)).to(dev)
    vfull = np.zeros(p.m)
    vfull[w["viol_idx"]] = w["viol_val"]
    d_v = torch.from_numpy(vfull).to(dev)
    d_hx = torch.empty(p.n, dtype=torch.float64, device=dev)
    d_jx = torch.empty(p.m, dtype=torch.float64, device=dev)
    d_jtv = torch.empty(p.n, dtype=torch.float64, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev)  # 512 MB > 126 MB L2
    torch.cuda.synchronize()

    def rhs_ptr(i):  # the two special solves, then the projections cycling over the distinct right-hand sides
        return d_rhs[i if i < 2 else 2 + (i - 2) % (n_distinct - 2)].data_ptr()

    def device_step():
        fact.refactor_device(d_val.data_ptr())
        fact.solve_device(rhs_ptr(0), d_sol.data_ptr())
        fact.solve_device(rhs_ptr(1), d_sol.data_ptr())
        mJ.mult_vec_trans_device(d_v.data_ptr(), d_jtv.data_ptr())
        for i in range(k):
            mH.mult_vec_device(d_x.data_ptr(), d_hx.data_ptr())
            fact.solve_device(rhs_ptr(2 + i), d_sol.data_ptr())
        mJ.mult_vec_device(d_x.data_ptr(), d_jx.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(warm):
        device_step()
    barrier()
    launches0 = lib.b200_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    if args.profile_step and main:
        torch.cuda.profiler.start()  # ncu --profile-from-start off: capture exactly the timed steps
    sampler.mark(0)
    for s in range(steps):
        with torch.cuda.stream(stream):
            flush.fill_(1.0)  # L2 flush between timed steps, outside the step's event pair
        ev[s][0].record(stream)
        device_step()
        ev[s][1].record(stream)
    barrier()
    sampler.mark(1)
    if args.profile_step and main:
        torch.cuda.profiler.stop()
    launches = lib.b200_launch_count() - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    extended = 0
    if sampler.thread and sampler.samples_in_window() < 3:
        # the timed region is shorter than three sampling periods: keep the same load running (untimed) until the
        # window holds three samples, and say so
        while sampler.samples_in_window() < 3 and extended < 2000:
            device_step()
            torch.cuda.synchronize()
            extended += 1
            sampler.mark(1)
    clocks = sampler.stop()
    if extended:
        clocks["window"] = f"timed region + {extended} untimed identical steps (the region is shorter than three sampling periods)"
    ms_dev = shard.max_over_ranks(float(step_ms.mean()), dist, dev)
    value = world * 1e3 / ms_dev

    # ---- break-down and rooflines per kernel class (live, CUDA events) -------------------------------------
    fact.refactor_device(d_val.data_ptr())
    st = fact.stats()
    factor_ms = st["ms_numeric"]
    phases = fact.profile_solve(20 if p.N < 2_000_000 else 5)  # pre, fwd, bwd, post
    solve_ms = float(phases.sum())
    classes = fact.profile_numeric()  # eager run with events between the kernel classes
    n_solves = len(rhs)
    hbm, fp64 = peaks["hbm_gbs"], peaks["fp64_tflops"]

    def hbm_roof(kernel, nbytes, ms, launches_, extra=None):
        a = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        extra = dict(extra or {})
        r = {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
             "traffic": ncu_traffic_per_launch(extra.pop("traffic_key", kernel), p.name),
             "peak_source": peak_src, "bytes_per_launch": int(nbytes), "ms_per_launch": ms, "launches_per_step": launches_}
        r.update(extra or {})
        return r

    def fp64_roof(kernel, flops, ms, note):
        a = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        return {"kernel": kernel, "bound": "tensor", "achieved": a, "peak": fp64, "unit": "TFLOP/s", "frac": a / fp64, "traffic": None,
                "peak_source": "cuBLAS DGEMM 8192^3 measured on this pool (profiles/r01_fp64_dgemm_peak.json); MEASURED_PEAKS.json has no FP64 entry",
                "flops_per_factorization": flops, "ms_per_factorization": ms, "note": note}

    # algorithmic bytes of one sweep: the factor once (exact nnz(L), 8 B), the row indices of every
    # supernode (4 B), the right-hand side in and out (8 B each), plus the pivots for the backward sweep
    sweep_bytes = 8 * st["nnz_L"] + 4 * st["n_row_idx"] + 16 * st["n_reduced"]
    # SpMV: 12 nnz + 4 (ncols + 1) + 8 ncols + 8 nrows; each launch timed alone after an L2 flush
    def time_spmv(fn, reps=10):
        tot = 0.0
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn()
            b.record(stream)
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps

    spmv_ms = time_spmv(lambda: mH.mult_vec_device(d_x.data_ptr(), d_hx.data_ptr()))
    spmvj_ms = time_spmv(lambda: mJ.mult_vec_device(d_x.data_ptr(), d_jx.data_ptr()))
    spmvt_ms = time_spmv(lambda: mJ.mult_vec_trans_device(d_v.data_ptr(), d_jtv.data_ptr()))
    spmv_bytes = lambda A: 12 * A.nnz + 4 * (A.shape[1] + 1) + 8 * A.shape[1] + 8 * A.shape[0]  # noqa: E731
    # which kernels a sweep consists of: sparse subtrees (sst.cu, one ticketed launch for all generations), the dataflow
    # kernel over the dense supernodes (solve.cu), or the first followed by the second
    from sleqp_b200.fact import Symbolic

    sym = Symbolic(p.N, w["cp"], w["ri"], w["v"])  # plan cache hit: the handle analysed the same pattern
    n_sst, n_dense_tasks = len(sym.export("sst")), len(sym.export("ffl_tasks"))
    sym.close()
    names = {d: " + ".join(([f"k_sst_{'forward' if d == 'fwd' else 'backward'}"] if n_sst else []) + ([f"k_flow<{d}>"] if n_dense_tasks else []))
             for d in ("fwd", "bwd")}
    if n_sst and n_dense_tasks:
        names["bwd"] = "k_flow<bwd> + k_sst_backward"
    sweep_extra = {"tree_levels": st["n_levels"], "stored_bytes": 8 * st["panel_doubles"], "sparse_subtrees": n_sst, "dense_sweep_tasks": n_dense_tasks}
    rooflines = {
        "k_flow_fwd": hbm_roof(names["fwd"], sweep_bytes, float(phases[1]), n_solves,
                               dict(sweep_extra, traffic_key="k_flow_fwd" if n_dense_tasks else "k_sst_forward")),
        "k_flow_bwd": hbm_roof(names["bwd"], sweep_bytes + 8 * st["n_reduced"], float(phases[2]), n_solves,
                               dict(sweep_extra, traffic_key="k_flow_bwd" if n_dense_tasks else "k_sst_backward")),
        "k_pre+k_post": hbm_roof("k_pre+k_post", 2 * (12 * len(w["ri"]) + 16 * p.N), float(phases[0] + phases[3]), n_solves),
        "k_update": fp64_roof("k_update", st["flops_update"], classes["update"], "all in-panel + Schur DMMA tiles of one factorization, eager launch-by-launch timing"),
        "k_inv_gemm": fp64_roof("k_inv_gemm", st["flops_inv"], classes["inv_gemm"], "selective inversion tiles"),
        "k_panel": {"kernel": "k_panel", "bound": "latency", "ms_per_factorization": classes["panel"], "stages": st["n_stages"],
                    "us_per_stage": 1e3 * classes["panel"] / max(1, st["n_stages"]), "note": "32-column panel steps on the critical path of the supernodal tree"},
        "factor_total": fp64_roof("numeric factorization (graph)", st["flops_factor"], factor_ms, "exact flops sum cc_j^2 over the whole captured graph"),
        "spmv_hess": hbm_roof("k_spmv_gather(H)", spmv_bytes(H), spmv_ms, k),
        "spmv_jac": hbm_roof("k_spmv_gather(J)", spmv_bytes(J), spmvj_ms, 1),
        "spmv_jac_trans": hbm_roof("k_spmv_gather(J^T)", spmv_bytes(J), spmvt_ms, 1),
    }
    share = {"numeric_factor(graph)": factor_ms, "k_flow_fwd": float(phases[1]) * n_solves, "k_flow_bwd": float(phases[2]) * n_solves,
             "k_pre+k_post": float(phases[0] + phases[3]) * n_solves, "spmv": spmv_ms * k + spmvj_ms + spmvt_ms}
    share.update({f"factor:{c}": v_ for c, v_ in classes.items()})
    # the dominant kernel of the step: the larger of the sweep kernels and the largest factorization class
    cand = {"k_flow_fwd": share["k_flow_fwd"], "k_flow_bwd": share["k_flow_bwd"], "k_update": classes["update"], "k_inv_gemm": classes["inv_gemm"]}
    dom = max(cand, key=cand.get)
    roofline = dict(rooflines[dom])
    roofline["step_share_ms"] = {k_: float(v_) for k_, v_ in share.items()}
    roofline["why"] = "largest kernel class of the timed step by device time"

    # ---- end to end through the plugin calls with host buffers ------------------------------------------
    # host inputs and outputs of the step live in page-locked memory (the contract's "pinned host memory"): the
    # library then DMAs from / to them directly instead of staging through its own pinned buffers
    _keep = []

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        _keep.append(t)
        return t.numpy()

    xh_idx = np.arange(p.n, dtype=np.int32)
    xh = pinned(w["rng"].standard_normal(p.n))
    vh_idx, vh = w["viol_idx"], pinned(w["viol_val"])
    kv = pinned(w["v"])
    rhs_h = [(kind, idx, pinned(val), b, e) for kind, idx, val, b, e in rhs[:n_distinct]]
    while len(rhs_h) < len(rhs):
        rhs_h.append(rhs_h[2 + (len(rhs_h) - 2) % max(1, n_distinct - 2)])
    buf_n, buf_m = pinned(np.empty(p.n)), pinned(np.empty(p.m))

    def host_step():
        fact.set_matrix(p.N, w["cp"], w["ri"], kv)
        out = None
        for i, (kind, idx, val, b, e) in enumerate(rhs_h):
            if i == 2:
                mJ.mult_vec_trans(vh_idx, vh, 0.0, out=buf_n)
            if i >= 2:
                mH.mult_vec(xh_idx, xh, out=buf_n)
            fact.solve(idx, val, p.N)
            out = fact.solution(b, e, 1e-20)
        mJ.mult_vec(xh_idx, xh, out=buf_m)
        return out

    h2d = 8 * len(w["v"]) + sum(8 * len(val) for _, _, val, _, _ in rhs_h) + 12 * len(vh_idx) + 8 * p.n * (k + 1)
    # a solution slice comes back sparsified: values (8 B) + indices (4 B), copied at full length (fact.cu)
    d2h = sum(12 * (e - b) for _, _, _, b, e in rhs_h) + 8 * p.n + 8 * p.n * k + 8 * p.m
    e2e_steps = max(3, min(steps, 10))
    mJ.set_stream(0)
    mH.set_stream(0)
    host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    torch.cuda.synchronize()
    e2e_ms = shard.max_over_ranks(1e3 * (time.perf_counter() - t0) / e2e_steps, dist, dev)
    # factor / solve through the boundary, separately (absolute times the north-star asks for)
    t0 = time.perf_counter()
    fact.set_matrix(p.N, w["cp"], w["ri"], kv)
    e2e_factor_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    for kind, idx, val, b, e in rhs_h[2:7]:
        fact.solve(idx, val, p.N)
        fact.solution(b, e, 1e-20)
    e2e_solve_ms = 1e3 * (time.perf_counter() - t0) / len(rhs_h[2:7])

    # device-resident projected CG (behind SleqpTRSolver via host/tr_b200.c): the same inner loop
    # without the per-iteration boundary crossing -- gradient in, step out
    cgs = ProjectedCG(fact, mH)
    g_idx = np.arange(p.n, dtype=np.int32)
    g_val = w["rng"].standard_normal(p.n)
    cgs.solve(p.n, g_idx, g_val, 1e8, 1e-6, k)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, cg_it, cg_how = cgs.solve(p.n, g_idx, g_val, 1e8, 1e-6, k)
    cg_ms = 1e3 * (time.perf_counter() - t0)

    def host_step_device_cg():
        fact.set_matrix(p.N, w["cp"], w["ri"], kv)
        for kind, idx, val, b, e in rhs_h[:2]:
            fact.solve(idx, val, p.N)
            fact.solution(b, e, 1e-20)
        mJ.mult_vec_trans(vh_idx, vh, 0.0, out=buf_n)
        out = cgs.solve(p.n, g_idx, g_val, 1e8, 1e-6, k)
        mJ.mult_vec(xh_idx, xh, out=buf_m)
        return out

    host_step_device_cg()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step_device_cg()
    torch.cuda.synchronize()
    e2e_cg_ms = shard.max_over_ranks(1e3 * (time.perf_counter() - t0) / e2e_steps, dist, dev)
    h2d_cg = 8 * len(w["v"]) + sum(8 * len(val) for _, _, val, _, _ in rhs_h[:2]) + 12 * len(vh_idx) + 12 * p.n + 8 * p.n
    d2h_cg = sum(12 * (e - b) for _, _, _, b, e in rhs_h[:2]) + 8 * p.n + 8 * p.n + 8 * p.m
    device_cg = {"e2e_step_ms": e2e_cg_ms, "e2e_value": world * 1e3 / e2e_cg_ms, "iterations": int(cg_it), "exit": int(cg_how), "ms": cg_ms, "ms_per_iteration": cg_ms / max(1, cg_it),
                 "h2d_bytes_per_step": int(h2d_cg), "d2h_bytes_per_step": int(d2h_cg),
                 "note": "b200_cg_solve (the SleqpTRSolver of host/tr_b200.c): gradient in, step out; 1 SpMV + 1 KKT solve + vector kernels per iteration on the device"}
    cgs.release()

    out = {
        "value": value, "ms_per_step": ms_dev, "launches": int(launches), "clocks": clocks,
        "config": {"workload": p.name, "N": p.N, "n": p.n, "ws_size": p.ws_size, "nnz_K": int(len(w["ri"])), "nnz_L": st["nnz_L"], "nnz_L_stored": st["nnz_L_stored"],
                   "supernodes": st["n_supernodes"], "tree_levels": st["n_levels"], "stages": st["n_stages"], "max_front": st["max_front"],
                   "cg_iters": k, "solves_per_step": n_solves, "l2": "512 MB buffer written between timed steps",
                   "per_gpu": "each rank runs its own independent instance (replicas, no collective on the data path)"},
        "factor_ms": factor_ms, "solve_ms": solve_ms, "factor_cold_ms_incl_symbolic": cold_ms, "symbolic_ms": st0["ms_symbolic"],
        "refine_steps": st["refine_steps"], "probe_residual": st["probe_residual"],
        "e2e": {"value": world * 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
                "factor_ms": e2e_factor_ms, "solve_ms": e2e_solve_ms,
                "via": "C-ABI (ctypes mirror of SleqpFactCallbacks / sleqp_mat_mult_vec): set_matrix, 32 x (solve + sparse solution), 32 SpMV with pinned host buffers"},
        "roofline": roofline, "rooflines": rooflines, "device_cg": device_cg, "w": w, "rhs": rhs,
    }
    torch.cuda.synchronize()
    for obj in (mJ, mH, fact):
        obj.release()
    del d_rhs, d_val, flush
    torch.cuda.empty_cache()
    return out


def pin_rank_to_cores(rank, world):
    """One process per GPU on one host: every rank gets a disjoint slice of the cores the job may use and its host-side
    analysis threads are capped to that slice (8 ranks x 4-8 analysis threads + samplers on one 32-core
    affinity set cost 20 % of the end-to-end scaling)."""
    if world <= 1:
        return None
    try:
        cores = sorted(os.sched_getaffinity(0))
        local = int(os.environ.get("LOCAL_RANK", rank))
        mine = None
        try:
            # the cores next to this rank's GPU (same NUMA node / PCIe root): page-locked host buffers are first touched
            # there, so the copy engines do not cross the socket interconnect. Ranks whose GPUs share a node split its cores.
            import pynvml

            pynvml.nvmlInit()
            words = (max(cores) // 64) + 1

            def near(gpu):
                h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
                mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
                return tuple(c for c in cores if (mask[c // 64] >> (c % 64)) & 1)

            sets = [near(g) for g in range(world)]
            peers = [g for g in range(world) if sets[g] == sets[local]]
            node = list(sets[local])
            if node:
                per = max(1, len(node) // len(peers))
                i = peers.index(local)
                mine = node[i * per:(i + 1) * per] or node
        except Exception:
            mine = None
        if not mine:
            per = max(1, len(cores) // world)
            mine = cores[rank * per:(rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        os.environ["B200_HOST_THREADS"] = str(len(mine))
        os.environ["B200_ND_THREADS"] = str(len(mine))
        os.environ.setdefault("OMP_NUM_THREADS", str(len(mine)))
        return len(mine)
    except Exception:
        return None


def run_ours(args, rank, world, local_rank):
    cores_per_rank = pin_rank_to_cores(rank, world)
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    t_start = time.perf_counter()
    main = measure_config(args.config, args, torch, dev, local_rank, rank, world, dist, args.steps, True)
    w, rhs, k = main.pop("w"), main.pop("rhs"), args.cg_iters

    # reference-driven end-to-end leg: the reference's own EQP code over the shipped glue. Every rank runs its own instance
    # on its own GPU at the same time (replicas, like `value`); whole-job rate = ranks / slowest rank's time per step
    ref_e2e = None
    if not args.no_sub:
        kind, size = {0: ("chain", 100), 1: ("poisson", 354), 2: ("chain", 1_000_000), 3: ("poisson3", 48), 4: ("poisson", 128)}[args.config]
        if dist is not None:
            dist.barrier()
        ref_e2e = reference_driver_e2e(kind, size, k, max(3, min(args.steps, 5)), local_rank)
        if world > 1:
            from sleqp_b200 import shard

            ok = bool(ref_e2e) and "ms_per_step" in ref_e2e
            ms = shard.max_over_ranks(float(ref_e2e["ms_per_step"]) if ok else 1e30, dist, dev)
            if ms >= 1e29:
                ref_e2e = {"error": "the reference-driven leg failed on at least one rank: " + str((ref_e2e or {}).get("error", "no output"))[:200]}
            else:
                ref_e2e = dict(ref_e2e, ms_per_step=ms, iters_per_s=world * 1e3 / ms, ranks=world,
                               h2d_bytes_per_step=world * ref_e2e["h2d_bytes_per_step"], d2h_bytes_per_step=world * ref_e2e["d2h_bytes_per_step"],
                               timing="every rank runs oracle/_ref/eqp_step_b200 on its own GPU concurrently (started after a barrier); max over ranks of "
                                      "the per-step time each process measures")
    subs = {}
    if world == 1 and not args.no_sub and args.config == DEFAULT_CONFIG:
        free, _ = torch.cuda.mem_get_info(dev)
        plan = [("config2", 1, {}), ("config4", 3, {"g": args.grid3d or pick_3d_grid(free)})]
        for name, idx, kw in plan:
            if time.perf_counter() - t_start > 240:  # keep the default run within a few minutes
                subs[name] = {"skipped": "time budget of the default run"}
                continue
            try:
                r = measure_config(idx, args, torch, dev, local_rank, rank, world, dist, 3, False, kw)
                r.pop("w"), r.pop("rhs")
                r["steps"] = 3
                if name == "config2":  # the same reference-driven end-to-end leg as the headline (config 4: too long here)
                    rd = reference_driver_e2e("poisson", 354, k, 3, local_rank)
                    if rd and "ms_per_step" in rd:
                        r["e2e"] = {"value": rd["iters_per_s"], "unit": UNIT, "ms_per_step": rd["ms_per_step"], "h2d_bytes_per_step": rd["h2d_bytes_per_step"],
                                    "d2h_bytes_per_step": rd["d2h_bytes_per_step"], "via": "reference-driven (oracle/_ref/eqp_step_b200), like the headline",
                                    "detail": rd, "c_abi": r["e2e"]}
                subs[name] = r
            except Exception as e:  # noqa: BLE001  (a sub-line must not take the headline down)
                subs[name] = {"error": repr(e)[:300]}
                torch.cuda.empty_cache()

    batch_leg = None
    if world > 1 and not args.no_sub:
        batch_leg = batch_step(args, rank, world, local_rank, torch, dev, dist, 64, 4, max(2, min(args.steps, 5)))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = cpu_step_sample(w, rhs, k, with_spd=True)
        cpu = {"value": 1.0 / s["step_s"], "unit": UNIT, "cores": 1, "host_cores_available": host_cores(), "kind": "port",
               "factor_ms": 1e3 * s["factor_s"], "solve_ms": 1e3 * s["solve_s"], "hess_spmv_ms": 1e3 * s["hess_spmv_s"], "jt_spmv_ms": 1e3 * s["jt_spmv_s"],
               "j_spmv_ms": 1e3 * s["j_spmv_s"], "spd_form_ms": 1e3 * s["spd_form_s"], "spd_factor_ms": 1e3 * s["spd_factor_s"], "spd_solve_ms": 1e3 * s["spd_solve_s"],
               "ratios_device": {"factor": 1e3 * s["factor_s"] / main["factor_ms"], "solve": 1e3 * s["solve_s"] / main["solve_ms"],
                                 "factor_vs_spd_stand_in": 1e3 * (s["spd_form_s"] + s["spd_factor_s"]) / main["factor_ms"],
                                 "solve_vs_spd_stand_in": 1e3 * s["spd_solve_s"] / main["solve_ms"]},
               "ratios_e2e": {"factor": 1e3 * s["factor_s"] / main["e2e"]["factor_ms"], "solve": 1e3 * s["solve_s"] / main["e2e"]["solve_ms"]},
               "sample": f"1 SuperLU factorization of K + 4 of {len(rhs)} solves scaled + 1 of {k} Hessian SpMV scaled + J SpMV, one pass (J^T product reported, not in the step); "
                         f"SpMV = {s['spmv_code']}; {CPU_NOTE}"}

    if rank == 0:
        e2e = dict(main["e2e"])
        if ref_e2e and "error" not in ref_e2e:
            # headline e2e: the shipped drop-in as SLEQP would drive it (reference code on top, host vectors)
            e2e = {"value": ref_e2e["iters_per_s"], "unit": UNIT, "h2d_bytes_per_step": ref_e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": ref_e2e["d2h_bytes_per_step"],
                   "ms_per_step": ref_e2e["ms_per_step"], "factor_ms": ref_e2e["set_iterate_ms"], "solve_ms": ref_e2e["solve_ms"],
                   "via": "reference-driven: the reference's own problem / iterate / working-set / aug_jac.c / tr_solver.c / SleqpVec code (oracle/_ref/eqp_step_b200, "
                          "unmodified sources) calling the shipped plugins host/aug_jac/b200_aug_jac.c, host/tr/tr_b200.c, host/sparse/mat_b200.c with host vectors",
                   "detail": ref_e2e, "c_abi": main["e2e"]}
        elif ref_e2e:
            e2e["reference_driver_error"] = ref_e2e["error"]
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": main["config"], "factor_ms": main["factor_ms"], "solve_ms": main["solve_ms"],
            "factor_cold_ms_incl_symbolic": main["factor_cold_ms_incl_symbolic"], "symbolic_ms": main["symbolic_ms"],
            "refine_steps": main["refine_steps"], "probe_residual": main["probe_residual"],
            "e2e": e2e, "gpu_launches": main["launches"], "clocks": main["clocks"],
            "roofline": main["roofline"], "rooflines": main["rooflines"], "device_cg": main["device_cg"],
        }
        if subs:
            line["configs"] = {name: ({kk: vv for kk, vv in r.items() if kk != "launches"} | {"gpu_launches": r.get("launches")}) if "value" in r else r for name, r in subs.items()}
        if batch_leg is not None:
            line["config5_batch"] = batch_leg
        if cores_per_rank is not None:
            line["host_cores_per_rank"] = cores_per_rank
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def batch_step(args, rank, world, local_rank, torch, dev, dist, batch, cfg_idx, steps):
    """Config 5: `batch` independent instances (same pattern family, different seeds) sharded over the ranks
    (instance i -> rank i mod world); on each GPU every instance has its own handle and CUDA stream, so the small
    latency-bound systems overlap. One step = one EQP inner loop of every instance. Strong scaling in the batch."""
    from sleqp_b200 import Fact, _lib, shard

    mine = shard.assign_instances(batch, world, rank)
    k = args.cg_iters
    inst = []
    for i in mine:
        w = make_workload(cfg_idx, seed=i)
        p = w["p"]
        f = Fact(device=local_rank)
        f.set_matrix(p.N, w["cp"], w["ri"], w["v"])
        rhs = step_rhs(w, 2)
        d_val = torch.from_numpy(w["v"]).to(dev)
        d_rhs = []
        for kind, idx, val, b, e in rhs[:4]:  # 4 distinct right-hand sides are cycled
            full = np.zeros(p.N)
            full[idx] = val
            d_rhs.append(torch.from_numpy(full).to(dev))
        inst.append(dict(f=f, d_val=d_val, d_rhs=d_rhs, d_sol=torch.empty(p.N, dtype=torch.float64, device=dev), p=p))
    n_solves = 2 + k
    torch.cuda.synchronize()

    def step():
        for it in inst:
            it["f"].refactor_device(it["d_val"].data_ptr())
        for s_ in range(n_solves):
            for it in inst:
                it["f"].solve_device(it["d_rhs"][s_ % 4].data_ptr(), it["d_sol"].data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    lib = _lib.lib()
    l0 = lib.b200_launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / steps
    clocks = sampler.stop()
    launches = lib.b200_launch_count() - l0
    ms = shard.max_over_ranks(ms, dist, dev)
    p0 = inst[0]["p"]
    st = inst[0]["f"].stats()
    out = {"metric": "sleqp_eqp_instance_iterations_per_s", "value": batch * 1e3 / ms, "unit": "instance-iter/s", "n_gpus": world,
           "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": p0.name + f"_x{batch}", "N": p0.N, "instances": batch, "instances_on_rank0": len(inst), "sharding": "instance i on rank i mod N, no collective",
                      "cg_iters": k, "solves_per_step": n_solves, "timing": "wall clock between device synchronisations (all streams), max over ranks",
                      "symbolic_cached": st["symbolic_cached"]},
           "gpu_launches": int(launches), "clocks": clocks}
    for it in inst:
        it["f"].release()
    return out


def run_batch(args, rank, world, local_rank):
    import torch

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    out = batch_step(args, rank, world, local_rank, torch, dev, dist, args.batch, args.config, args.steps)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=DEFAULT_CONFIG, help="BASELINE.json configs index (default 2: chained Rosenbrock n = 1e6, the headline config)")
    ap.add_argument("--cg-iters", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="only the main workload: no config-2 / config-4 sub-lines, no config-5 leg, no reference-driven e2e")
    ap.add_argument("--grid3d", type=int, default=0, help="grid of the config-4 sub-line (default: largest that fits the free HBM)")
    ap.add_argument("--profile-step", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    ap.add_argument("--batch", type=int, default=0, help="config 5 mode: this many independent instances sharded over the ranks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.batch > 0:
        if args.config == DEFAULT_CONFIG:
            args.config = 4
        run_batch(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
