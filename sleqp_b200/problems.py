"""Synthetic workloads of BASELINE.json: constraint Jacobians, working sets,
KKT matrices in the layout SLEQP's standard augmented Jacobian hands to a factorization backend
(reference: src/main/aug_jac/standard_aug_jac.c:135-237), Hessians and right-hand sides.

Pure numpy/scipy input generation; nothing here is on the measured path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp


@dataclass
class KKTProblem:
    name: str
    n: int  # variables
    m: int  # general constraints
    J: sp.csc_matrix  # m x n constraint Jacobian
    H: sp.csc_matrix  # n x n Hessian of the Lagrangian (for the EQP loop's SpMV)
    active_vars: np.ndarray  # variables at a bound (working set, ascending)
    active_cons: np.ndarray  # constraints in the working set (ascending)
    meta: dict = field(default_factory=dict)

    @property
    def ws_size(self) -> int:
        return int(len(self.active_vars) + len(self.active_cons))

    @property
    def N(self) -> int:
        return self.n + self.ws_size

    def working_rows(self) -> sp.csr_matrix:
        """A_W: active-variable unit rows first, then active constraint rows
        (working_set.c:117-180 index rule)."""
        na = len(self.active_vars)
        E = sp.csr_matrix((np.ones(na), (np.arange(na), self.active_vars)), shape=(na, self.n))
        return sp.vstack([E, self.J.tocsr()[self.active_cons]], format="csr")

    def kkt_lower(self):
        """CSC arrays (colptr, rowidx, val) of tril([I A_W^T; A_W 0]) with int32 indices, rows
        ascending per column, empty columns n..N-1 -- what SLEQP_FACT_FLAGS_LOWER backends get."""
        A = self.working_rows().tocsc()
        A.sort_indices()
        n, N = self.n, self.N
        cnt = 1 + np.diff(A.indptr)
        colptr = np.zeros(N + 1, dtype=np.int64)
        colptr[1 : n + 1] = np.cumsum(cnt)
        colptr[n + 1 :] = colptr[n]
        nnz = int(colptr[n])
        rows = np.empty(nnz, dtype=np.int32)
        val = np.empty(nnz, dtype=np.float64)
        rows[colptr[:n]] = np.arange(n)
        val[colptr[:n]] = 1.0
        mask = np.ones(nnz, dtype=bool)
        mask[colptr[:n]] = False
        rows[mask] = A.indices + n
        val[mask] = A.data
        return colptr.astype(np.int32), rows, val

    def kkt_full(self) -> sp.csc_matrix:
        A = self.working_rows()
        return sp.bmat([[sp.identity(self.n, format="csr"), A.T], [A, None]], format="csc")

    def rhs(self, kind: str, seed: int = 0):
        """(idx, val) of a right-hand side as the aug_jac solve `kind` produces it
        (standard_aug_jac.c:306-435): dense random in the relevant block, dim N."""
        rng = np.random.default_rng(seed)
        if kind in ("project_nullspace", "solve_lsq"):
            idx = np.arange(self.n, dtype=np.int32)
            val = rng.standard_normal(self.n)
        elif kind == "solve_min_norm":
            idx = np.arange(self.n, self.N, dtype=np.int32)
            val = rng.standard_normal(self.ws_size)
        else:
            raise ValueError(kind)
        return idx, val


def _chain_jacobian(x: np.ndarray):
    """c_k = x_{2k} x_{2k+1} + x_{2k+2} - 1, k = 0..m-1 (configs 1 and 3)."""
    n = len(x)
    m = (n - 2) // 2
    k = np.arange(m)
    rows = np.repeat(k, 3)
    cols = np.stack([2 * k, 2 * k + 1, 2 * k + 2], axis=1).ravel()
    vals = np.stack([x[2 * k + 1], x[2 * k], np.ones(m)], axis=1).ravel()
    return sp.csc_matrix((vals, (rows, cols)), shape=(m, n)), m


def _rosenbrock_hessian(x: np.ndarray) -> sp.csc_matrix:
    """Hessian of sum_{i<n-1} 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2: tridiagonal."""
    n = len(x)
    d = np.zeros(n)
    d[:-1] += 1200.0 * x[:-1] ** 2 - 400.0 * x[1:] + 2.0
    d[1:] += 200.0
    off = -400.0 * x[:-1]
    return sp.diags([off, d, off], [-1, 0, 1], format="csc")


def chain_rosenbrock(n: int, active_fraction: float = 0.0, seed: int = 0, name: str | None = None) -> KKTProblem:
    """Configs 1 (n=100) and 3 (n=1e6): chained Rosenbrock objective with sparse nonlinear equality
    constraints, bounds -2 <= x <= 2, x0 ~ U(0.5, 1.5). Active bounds are drawn from the
    even-indexed variables only, so every constraint row keeps its private variable x_{2k+1} and
    the working set has linearly independent rows (pub_working_set.h:42-44 requirement)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.5, 1.5, size=n)
    J, m = _chain_jacobian(x)
    even = np.arange(0, n, 2)
    na = int(round(active_fraction * len(even)))
    active_vars = np.sort(rng.choice(even, size=na, replace=False)) if na else np.zeros(0, dtype=np.int64)
    return KKTProblem(
        name=name or f"chain_rosenbrock_n{n}",
        n=n,
        m=m,
        J=J,
        H=_rosenbrock_hessian(x),
        active_vars=active_vars.astype(np.int64),
        active_cons=np.arange(m, dtype=np.int64),
        meta=dict(x0=x, seed=seed),
    )


def _laplacian(g: int, dim: int) -> sp.csr_matrix:
    """h^2-scaled (2*dim)+1-point Laplacian with Dirichlet boundary on a g^dim grid."""
    T = sp.diags([-np.ones(g - 1), 2.0 * np.ones(g), -np.ones(g - 1)], [-1, 0, 1], format="csr")
    I = sp.identity(g, format="csr")
    if dim == 2:
        return (sp.kron(I, T) + sp.kron(T, I)).tocsr()
    if dim == 3:
        return (sp.kron(sp.kron(I, I), T) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(T, I), I)).tocsr()
    raise ValueError(dim)


def poisson_control(g: int, dim: int = 2, active_fraction: float = 0.1, alpha: float = 1e-2, seed: int = 0, name: str | None = None) -> KKTProblem:
    """Configs 2, 4, 5: min 1/2 |y - y_d|^2 + alpha/2 |u|^2  s.t.  A y - u = f, u_lo <= u <= u_hi,
    x = (y, u), J = [A, -I] with A the h^2-scaled Laplacian (a row scaling of the discretised
    PDE). Working set: all equalities + a seeded fraction of the controls at a bound."""
    rng = np.random.default_rng(seed)
    A = _laplacian(g, dim)
    q = A.shape[0]
    J = sp.hstack([A, -sp.identity(q, format="csr")], format="csc")
    J.sort_indices()
    na = int(round(active_fraction * q))
    active_u = np.sort(rng.choice(q, size=na, replace=False)) if na else np.zeros(0, dtype=np.int64)
    H = sp.diags(np.concatenate([np.ones(q), alpha * np.ones(q)]), 0, format="csc")
    return KKTProblem(
        name=name or f"poisson{dim}d_control_g{g}",
        n=2 * q,
        m=q,
        J=J,
        H=H,
        active_vars=(q + active_u).astype(np.int64),
        active_cons=np.arange(q, dtype=np.int64),
        meta=dict(g=g, dim=dim, alpha=alpha, seed=seed),
    )


def config(idx: int, **kw) -> KKTProblem:
    """The BASELINE.json configs by index (0-based like `configs`)."""
    if idx == 0:
        return chain_rosenbrock(100, active_fraction=kw.pop("active_fraction", 0.2), name="config1_rosenbrock_n100", **kw)
    if idx == 1:
        return poisson_control(kw.pop("g", 354), 2, name="config2_poisson2d_g354", **kw)
    if idx == 2:
        return chain_rosenbrock(kw.pop("n", 1_000_000), name="config3_chain_n1e6", **kw)
    if idx == 3:
        return poisson_control(kw.pop("g", 48), 3, name="config4_poisson3d", **kw)
    if idx == 4:
        return poisson_control(kw.pop("g", 128), 2, name="config5_poisson2d_g128", **kw)
    raise ValueError(idx)


def eqp_harness_problem(n: int = 400) -> tuple[KKTProblem, np.ndarray]:
    """The problem oracle/eqp_harness.c builds in C (same x0 from its xorshift64, same working set: every 20th
    variable at its bound + every constraint). Returns (problem, objective gradient at x0)."""
    mask = (1 << 64) - 1
    state = 88172645463325252
    x = np.empty(n)
    for i in range(n):
        state ^= (state << 13) & mask
        state ^= state >> 7
        state ^= (state << 17) & mask
        x[i] = 0.5 + (state >> 11) / 9007199254740992.0
    J, m = _chain_jacobian(x)
    grad = np.zeros(n)
    grad[:-1] += -400.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2.0 * (1.0 - x[:-1])
    grad[1:] += 200.0 * (x[1:] - x[:-1] ** 2)
    p = KKTProblem(name=f"eqp_harness_n{n}", n=n, m=m, J=J, H=_rosenbrock_hessian(x), active_vars=np.arange(0, n, 20, dtype=np.int64),
                   active_cons=np.arange(m, dtype=np.int64), meta=dict(x0=x))
    return p, grad


def eqp_harness_poisson(g: int = 12, alpha: float = 1e-2) -> tuple[KKTProblem, np.ndarray]:
    """The second problem of oracle/eqp_harness.c ("poisson"): 2D Poisson control on a g x g grid with the same x0
    (xorshift64), target y_d and working set (every constraint + every 7th control at its upper bound)."""
    q = g * g
    n = 2 * q
    mask = (1 << 64) - 1
    state = 88172645463325252
    x = np.empty(n)
    for i in range(n):
        state ^= (state << 13) & mask
        state ^= state >> 7
        state ^= (state << 17) & mask
        x[i] = 0.5 + (state >> 11) / 9007199254740992.0
    A = _laplacian(g, 2)
    J = sp.hstack([A, -sp.identity(q, format="csr")], format="csc")
    J.sort_indices()
    i = np.arange(q)
    yd = np.sin(np.pi * (i % g + 1.0) / (g + 1.0)) * np.sin(np.pi * (i // g + 1.0) / (g + 1.0))
    grad = np.concatenate([x[:q] - yd, alpha * x[q:]])
    H = sp.diags(np.concatenate([np.ones(q), alpha * np.ones(q)]), 0, format="csc")
    p = KKTProblem(name=f"eqp_harness_poisson_g{g}", n=n, m=q, J=J, H=H, active_vars=np.arange(q, n, 7, dtype=np.int64),
                   active_cons=np.arange(q, dtype=np.int64), meta=dict(x0=x, g=g, alpha=alpha))
    return p, grad
