#!/usr/bin/env python
"""Benchmark of the KKT hot path (BASELINE.json): one "step" = the EQP inner loop of one SQP
iteration on a fixed synthetic KKT system (unit iii):

    1 numeric factorization of [I A_W^T; A_W 0] (symbolic analysis cached)
  + (2 + k) solves   (1 min-norm, 1 LSQ multipliers, k null-space projections of the CG loop)
  + k Hessian SpMV, 1 Jacobian SpMV^T, 1 Jacobian SpMV

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4] [--cg-iters k]

`value` (SLEQP EQP iterations per second) is timed with CUDA events on the handle's stream with every
input resident in HBM; `e2e` is the same step through the reference-facing plugin calls with host
buffers (set_matrix / solve / solution / mult_vec with numpy arrays), copies inside the timed region.
`--impl reference` times the reference's CPU path for the same step on the host cores: the
reference's own SpMV (oracle/_ref, compiled from the unmodified sources) and, for the factorization
arithmetic that lives in absent SuiteSparse, SciPy SuperLU as the stand-in (see oracle/sleqp_oracle.py).
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


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic_per_launch(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of the same workload (profiles/, one sweep = one launch); None if there is no capture
    for this workload."""
    import csv

    if "config2" not in workload:
        return None
    path = os.path.join(ROOT, "profiles", f"r01_ncu_full_{kernel}_config2.csv")
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
        return None


def make_workload(cfg_idx, seed=0):
    from sleqp_b200 import problems

    p = problems.config(cfg_idx, seed=seed)
    cp, ri, v = p.kkt_lower()
    J = p.J.tocsc()
    J.sort_indices()
    H = p.H.tocsc()
    H.sort_indices()
    rng = np.random.default_rng(100 + seed)
    return dict(p=p, cp=cp, ri=ri, v=v, J=J, H=H, rng=rng)


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
def run_reference(args, rank, world):
    """CPU arm: the same step on the host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import ref_lib
    from oracle import sleqp_oracle as orc

    w = make_workload(args.config)
    p, k = w["p"], args.cg_iters
    rhs = step_rhs(w, k)
    n_solve_sample = min(4, len(rhs))
    ref = ref_lib.RefLib("lapack") if ref_lib.available("lapack") else None
    J, H = w["J"], w["H"]
    xd_idx = np.arange(p.n, dtype=np.int32)
    xd = w["rng"].standard_normal(p.n)
    vd_idx = np.arange(p.m, dtype=np.int32)
    vd = w["rng"].standard_normal(p.m)
    try:
        import threadpoolctl

        cores = max(i["num_threads"] for i in threadpoolctl.threadpool_info()) if threadpoolctl.threadpool_info() else 1
    except Exception:
        cores = 1

    def spmv(A, idx, val, trans):
        if ref is not None:
            if trans:
                return ref.mat_mult_vec_trans(A.shape[0], A.shape[1], A.indptr, A.indices, A.data, idx, val, 0.0)
            return ref.mat_mult_vec(A.shape[0], A.shape[1], A.indptr, A.indices, A.data, idx, val)
        if trans:
            return orc.mat_mult_vec_trans(A.shape[1], A.indptr, A.indices, A.data, idx, val, A.shape[0], 0.0)
        return orc.mat_mult_vec(A.shape[0], A.indptr, A.indices, A.data, idx, val)

    def one_step():
        t0 = time.perf_counter()
        lu = orc.SparseLU()
        lu.set_matrix(p.N, w["cp"], w["ri"], w["v"])
        t1 = time.perf_counter()
        for kind, idx, val, b, e in rhs[:n_solve_sample]:
            lu.solve(idx, val)
            lu.solution(b, e, 1e-20)
        t2 = time.perf_counter()
        spmv(H, xd_idx, xd, False)
        t3 = time.perf_counter()
        spmv(J, vd_idx, vd, True)
        spmv(J, xd_idx, xd, False)
        t4 = time.perf_counter()
        factor, solve1, hspmv, jspmv = t1 - t0, (t2 - t1) / n_solve_sample, t3 - t2, t4 - t3
        return factor + solve1 * len(rhs) + hspmv * k + jspmv, factor, solve1

    # keep the whole run within a few minutes whatever --steps/--warmup are: one step is a full sparse LU on the
    # host (seconds); the first step is timed to size the rest
    budget_s = 150.0
    t_probe = time.perf_counter()
    first = one_step()
    t_probe = time.perf_counter() - t_probe
    n_warm = max(0, min(args.warmup - 1, int(0.2 * budget_s / max(t_probe, 1e-3))))
    for _ in range(n_warm):
        one_step()
    n_meas = max(1, min(args.steps, int(0.8 * budget_s / max(t_probe, 1e-3))))
    tot, fac, sol = [], [], []
    if args.warmup == 0:  # the probe step counts as the first measured step
        tot.append(first[0]); fac.append(first[1]); sol.append(first[2])
    while len(tot) < n_meas:
        a, b, c = one_step()
        tot.append(a)
        fac.append(b)
        sol.append(c)
    ms = 1e3 * float(np.mean(tot))
    value = 1e3 / ms
    sample = (f"per step: 1 SuperLU factorization + {n_solve_sample} of {len(rhs)} solves (scaled x{len(rhs)}/{n_solve_sample}) + 1 of {k} Hessian SpMV "
              f"(scaled x{k}) + 1 J^T + 1 J SpMV; SpMV = reference sleqp_mat_mult_vec{'/_trans (oracle/_ref)' if ref is not None else ' (numpy port)'}; "
              "factor/solve = SciPy SuperLU stand-in for the absent Umfpack (sequential code; `cores` = the threads its BLAS calls may use)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(tot), "warmup": (n_warm + 1 if args.warmup else 0),
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["p"].name, "N": p.N, "nnz_K": int(len(w["ri"])), "cg_iters": k, "solves_per_step": len(rhs)},
        "factor_ms": 1e3 * float(np.mean(fac)), "solve_ms": 1e3 * float(np.mean(sol)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch

    from sleqp_b200 import Fact, Mat, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    w = make_workload(args.config, seed=rank if world > 1 else 0)
    p, k = w["p"], args.cg_iters
    rhs = step_rhs(w, k)
    J, H = w["J"], w["H"]
    lib = _lib.lib()

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
    d_rhs = []
    for kind, idx, val, b, e in rhs:
        full = np.zeros(p.N)
        full[idx] = val
        d_rhs.append(torch.from_numpy(full).to(dev))
    d_sol = torch.empty(p.N, dtype=torch.float64, device=dev)
    d_x = torch.from_numpy(w["rng"].standard_normal(p.n)).to(dev)
    d_v = torch.from_numpy(w["rng"].standard_normal(p.m)).to(dev)
    d_hx = torch.empty(p.n, dtype=torch.float64, device=dev)
    d_jx = torch.empty(p.m, dtype=torch.float64, device=dev)
    d_jtv = torch.empty(p.n, dtype=torch.float64, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev)  # 512 MB > 126 MB L2
    torch.cuda.synchronize()

    def device_step():
        fact.refactor_device(d_val.data_ptr())
        fact.solve_device(d_rhs[0].data_ptr(), d_sol.data_ptr())
        fact.solve_device(d_rhs[1].data_ptr(), d_sol.data_ptr())
        mJ.mult_vec_trans_device(d_v.data_ptr(), d_jtv.data_ptr())
        for i in range(k):
            mH.mult_vec_device(d_x.data_ptr(), d_hx.data_ptr())
            fact.solve_device(d_rhs[2 + i].data_ptr(), d_sol.data_ptr())
        mJ.mult_vec_device(d_x.data_ptr(), d_jx.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.b200_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if args.profile_step:
        torch.cuda.profiler.start()  # ncu --profile-from-start off: capture exactly the timed steps
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(1.0)  # L2 flush between timed steps, outside the step's event pair
        ev[s][0].record(stream)
        device_step()
        ev[s][1].record(stream)
    barrier()
    if args.profile_step:
        torch.cuda.profiler.stop()
    launches = lib.b200_launch_count() - launches0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    clocks = sampler.stop()
    from sleqp_b200 import shard

    ms_dev = shard.max_over_ranks(float(step_ms.mean()), dist, dev)
    value = world * 1e3 / ms_dev

    # ---- break-down and roofline of the dominant kernel (live, CUDA events) ----------------------------
    fact.refactor_device(d_val.data_ptr())
    st = fact.stats()
    factor_ms = st["ms_numeric"]
    phases = fact.profile_solve(20)  # pre, fwd, bwd, post
    solve_ms = float(phases.sum())
    peaks, peak_src = load_peaks()
    n_solves = len(rhs)
    # one dataflow kernel launch per sweep (solve.cu: k_flow<forward> / k_flow<backward>); the forward phase also
    # holds the reset of the accumulators (k_flow_reset, ~2 us)
    share = {"numeric_factor(graph)": factor_ms, "k_flow_fwd": phases[1] * n_solves, "k_flow_bwd": phases[2] * n_solves,
             "k_pre+k_post": (phases[0] + phases[3]) * n_solves}
    dom = max(("k_flow_fwd", "k_flow_bwd"), key=lambda x: share[x])
    # algorithmic bytes of one sweep: the factor once (exact nnz(L), 8 B), the row indices of
    # every supernode (4 B), the right-hand side in and out (8 B each), plus the pivots for the backward sweep
    sweep_bytes = 8 * st["nnz_L"] + 4 * st["n_row_idx"] + 16 * st["n_reduced"] + (8 * st["n_reduced"] if dom == "k_flow_bwd" else 0)
    sweep_ms = float(phases[1] if dom == "k_flow_fwd" else phases[2])
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": ncu_traffic_per_launch(dom, p.name), "peak_source": peak_src, "launches_per_sweep": 1,
                "bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_ms, "tree_levels": st["n_levels"],
                "step_share_ms": {k_: float(v_) for k_, v_ in share.items()}}

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
    vh_idx = np.arange(p.m, dtype=np.int32)
    vh = pinned(w["rng"].standard_normal(p.m))
    kv = pinned(w["v"])
    rhs = [(kind, idx, pinned(val), b, e) for kind, idx, val, b, e in rhs]

    buf_n, buf_m = pinned(np.empty(p.n)), pinned(np.empty(p.m))

    def host_step():
        fact.set_matrix(p.N, w["cp"], w["ri"], kv)
        out = None
        for i, (kind, idx, val, b, e) in enumerate(rhs):
            if i == 2:
                mJ.mult_vec_trans(vh_idx, vh, 0.0, out=buf_n)
            if i >= 2:
                mH.mult_vec(xh_idx, xh, out=buf_n)
            fact.solve(idx, val, p.N)
            out = fact.solution(b, e, 1e-20)
        mJ.mult_vec(xh_idx, xh, out=buf_m)
        return out

    h2d = 8 * len(w["v"]) + sum(8 * len(val) for _, _, val, _, _ in rhs) + 8 * p.m + 8 * p.n * (k + 1)
    # a solution slice comes back sparsified: values (8 B) + indices (4 B), copied at full length (fact.cu)
    d2h = sum(12 * (e - b) for _, _, _, b, e in rhs) + 8 * p.n + 8 * p.n * k + 8 * p.m
    e2e_steps = max(3, min(args.steps, 10))
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
    for kind, idx, val, b, e in rhs[2:7]:
        fact.solve(idx, val, p.N)
        fact.solution(b, e, 1e-20)
    e2e_solve_ms = 1e3 * (time.perf_counter() - t0) / len(rhs[2:7])

    # device-resident projected CG: the same inner loop without the per-iteration boundary crossing
    from sleqp_b200 import ProjectedCG

    mH.set_stream(0)
    cgs = ProjectedCG(fact, mH)
    g_idx = np.arange(p.n, dtype=np.int32)
    g_val = w["rng"].standard_normal(p.n)
    cgs.solve(p.n, g_idx, g_val, 1e8, 1e-6, k)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, cg_it, cg_how = cgs.solve(p.n, g_idx, g_val, 1e8, 1e-6, k)
    cg_ms = 1e3 * (time.perf_counter() - t0)
    # the same EQP step end to end with the CG loop on the device: host K in, host step out
    def host_step_device_cg():
        fact.set_matrix(p.N, w["cp"], w["ri"], kv)
        for kind, idx, val, b, e in rhs[:2]:
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
    device_cg = {"e2e_step_ms": e2e_cg_ms, "e2e_value": world * 1e3 / e2e_cg_ms, "iterations": int(cg_it), "exit": int(cg_how), "ms": cg_ms, "ms_per_iteration": cg_ms / max(1, cg_it),
                 "note": "b200_cg_solve: gradient in, step out; 1 SpMV + 1 KKT solve + vector kernels per iteration on the device"}
    cgs.release()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(w, rhs, k)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": p.name, "N": p.N, "n": p.n, "ws_size": p.ws_size, "nnz_K": int(len(w["ri"])), "nnz_L": st["nnz_L"],
                       "cg_iters": k, "solves_per_step": n_solves, "l2": "512 MB buffer written between timed steps",
                       "per_gpu": "each rank runs its own independent instance (replicas, no collective on the data path)"},
            "factor_ms": factor_ms, "solve_ms": solve_ms, "factor_cold_ms_incl_symbolic": cold_ms, "symbolic_ms": st0["ms_symbolic"],
            "refine_steps": st["refine_steps"], "probe_residual": st["probe_residual"],
            "e2e": {"value": world * 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "factor_ms": e2e_factor_ms, "solve_ms": e2e_solve_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "device_cg": device_cg,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    # release the native handles while the CUDA context is certainly alive
    torch.cuda.synchronize()
    mJ.set_stream(0)
    mH.set_stream(0)
    for obj in (mJ, mH, fact):  # (the CG handle was released above)
        obj.release()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_batch(args, rank, world, local_rank):
    """Config 5: `--batch B` independent instances (same pattern family, different seeds) sharded over the ranks
    (instance i -> rank i mod world); on each GPU every instance has its own handle and CUDA stream, so the small
    latency-bound systems overlap. One step = one EQP inner loop of every instance. Strong scaling in B."""
    import torch

    from sleqp_b200 import Fact, _lib, shard

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    mine = shard.assign_instances(args.batch, world, rank)
    k = args.cg_iters
    inst = []
    for i in mine:
        w = make_workload(args.config, seed=i)
        p = w["p"]
        f = Fact(device=local_rank)
        f.set_matrix(p.N, w["cp"], w["ri"], w["v"])
        rhs = step_rhs(w, k)
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
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()
    launches = lib.b200_launch_count() - l0
    ms = shard.max_over_ranks(ms, dist, dev)
    if rank == 0:
        p0 = inst[0]["p"]
        st = inst[0]["f"].stats()
        print(json.dumps({
            "metric": "sleqp_eqp_instance_iterations_per_s", "value": args.batch * 1e3 / ms, "unit": "instance-iter/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": p0.name + f"_x{args.batch}", "N": p0.N, "instances": args.batch, "instances_on_rank0": len(inst),
                       "cg_iters": k, "solves_per_step": n_solves, "timing": "wall clock between device synchronisations (all streams)",
                       "symbolic_cached": st["symbolic_cached"]},
            "gpu_launches": int(launches), "clocks": clocks,
        }), flush=True)
    for it in inst:
        it["f"].release()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_sample(w, rhs, k):
    """Bounded CPU sample of the same step (oracle / reference code; never the thing shipped)."""
    from oracle import ref_lib
    from oracle import sleqp_oracle as orc

    p = w["p"]
    ref = ref_lib.RefLib("lapack") if ref_lib.available("lapack") else None
    J, H = w["J"], w["H"]
    t0 = time.perf_counter()
    lu = orc.SparseLU()
    lu.set_matrix(p.N, w["cp"], w["ri"], w["v"])
    t1 = time.perf_counter()
    ns = min(4, len(rhs))
    for kind, idx, val, b, e in rhs[:ns]:
        lu.solve(idx, val)
        lu.solution(b, e, 1e-20)
    t2 = time.perf_counter()
    xi = np.arange(p.n, dtype=np.int32)
    xv = np.ones(p.n)
    vi = np.arange(p.m, dtype=np.int32)
    vv = np.ones(p.m)
    if ref is not None:
        ref.mat_mult_vec(H.shape[0], H.shape[1], H.indptr, H.indices, H.data, xi, xv)
        t3 = time.perf_counter()
        ref.mat_mult_vec_trans(J.shape[0], J.shape[1], J.indptr, J.indices, J.data, vi, vv, 0.0)
        ref.mat_mult_vec(J.shape[0], J.shape[1], J.indptr, J.indices, J.data, xi, xv)
    else:
        orc.mat_mult_vec(H.shape[0], H.indptr, H.indices, H.data, xi, xv)
        t3 = time.perf_counter()
        orc.mat_mult_vec_trans(J.shape[1], J.indptr, J.indices, J.data, vi, vv, J.shape[0], 0.0)
        orc.mat_mult_vec(J.shape[0], J.indptr, J.indices, J.data, xi, xv)
    t4 = time.perf_counter()
    step = (t1 - t0) + (t2 - t1) / ns * len(rhs) + (t3 - t2) * k + (t4 - t3)
    return {"value": 1.0 / step, "unit": UNIT, "cores": 1, "kind": "port",
            "factor_ms": 1e3 * (t1 - t0), "solve_ms": 1e3 * (t2 - t1) / ns,
            "sample": f"1 SciPy SuperLU factorization (stand-in for the absent Umfpack, single thread) + {ns} of {len(rhs)} solves scaled, "
                      f"1 of {k} Hessian SpMV scaled, J^T and J SpMV with the reference's sleqp_mat_mult_vec{'(_trans) from oracle/_ref' if ref is not None else ' numpy port'}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json configs index (default 1: 2D Poisson control, n~2.5e5)")
    ap.add_argument("--cg-iters", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true", help="cudaProfilerStart/Stop around the timed steps (for ncu --profile-from-start off)")
    ap.add_argument("--batch", type=int, default=0, help="config 5 mode: this many independent instances sharded over the ranks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.batch > 0:
        run_batch(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
