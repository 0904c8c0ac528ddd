"""SpMV / SpMV^T against the HBM roofline on a matrix far larger than L2 (J of a 2D Poisson control problem with
g = 2000: 4e6 x 8e6, 2.4e7 non-zeros, ~300 MB of algorithmic traffic per product), plus the config-2 sizes for
comparison (those fit L2 and are launch-bound). Algorithmic bytes: 12 nnz + 4 (segments + 1) + 8 x + 8 y.
Usage: python profiles/spmv_roofline.py [g]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sleqp_b200 import Mat, problems

peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6545.6))
dev = torch.device("cuda:0")
for g in ([int(sys.argv[1])] if len(sys.argv) > 1 else [354, 2000]):
    p = problems.poisson_control(g, 2)
    J = p.J.tocsc(); J.sort_indices()
    m, n, nnz = J.shape[0], J.shape[1], J.nnz
    M = Mat(device=0)
    M.set(m, n, J.indptr, J.indices, J.data)
    x = torch.randn(n, dtype=torch.float64, device=dev); y = torch.empty(m, dtype=torch.float64, device=dev)
    v = torch.randn(m, dtype=torch.float64, device=dev); z = torch.empty(n, dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)  # the handle launches on this stream; the events are recorded on it too
    M.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    for name, fn, segs in (("y = J x  ", lambda: M.mult_vec_device(x.data_ptr(), y.data_ptr()), m), ("y = J^T v", lambda: M.mult_vec_trans_device(v.data_ptr(), z.data_ptr()), n)):
        for _ in range(5):
            fn()
        reps = 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        byts = 12 * nnz + 4 * (segs + 1) + 8 * (n + m)
        # reference check against scipy on the host
        ref = (J @ x.cpu().numpy()) if "x" in name else (J.T @ v.cpu().numpy())
        got = (y if "x" in name else z).cpu().numpy()
        err = np.abs(got - ref).max() / max(1.0, np.abs(ref).max())
        print(f"g={g} {name} nnz={nnz} {ms*1e3:8.1f} us  {byts/ms/1e6:8.1f} GB/s = {100*byts/ms/1e6/peak:5.1f} % of {peak:.0f} GB/s   max rel err {err:.1e}")
    M.release()
