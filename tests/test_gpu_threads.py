"""Concurrency contract of the reference (src/test/thread_test.c:13,92-110: 8 solver instances on 8 threads):
independent factorization handles on different host threads share no mutable state."""
import threading

import numpy as np
import pytest

from sleqp_b200 import Fact, problems

pytestmark = pytest.mark.gpu


def test_eight_threads_independent_handles():
    errors = []

    def work(tid):
        try:
            p = problems.poisson_control(12 + tid, 2, seed=tid) if tid % 2 else problems.chain_rosenbrock(2000 + 100 * tid, 0.2, seed=tid)
            K = p.kkt_full()
            f = Fact()
            for rep in range(3):
                f.set_matrix(p.N, *p.kkt_lower())
                for kind in ("project_nullspace", "solve_min_norm", "solve_lsq"):
                    idx, val = p.rhs(kind, 10 * tid + rep)
                    f.solve(idx, val, p.N)
                    x = f.solution_dense(0, p.N)
                    b = np.zeros(p.N)
                    b[idx] = val
                    res = np.linalg.norm(K @ x - b) / np.linalg.norm(b)
                    if not res <= 1e-10:
                        errors.append((tid, kind, res))
            f.release()
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
