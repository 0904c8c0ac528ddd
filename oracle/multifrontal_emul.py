"""Task-by-task numpy emulation of the device numeric factorization and solve -- TEST
INFRASTRUCTURE ONLY (see oracle/sleqp_oracle.py header for the import rule).

It executes the *same plan* (assembly map, stages, extend-add / panel / update tasks, solve
levels) that the CUDA kernels in sleqp_b200/csrc/numeric.cu and solve.cu execute, one task at a
time, so that schedule or index-map defects show up on the CPU before any GPU time is spent, and
so that GPU intermediates (pivots D) can be compared against it. The update workspace is filled
with NaN first: a missing zero-fill or a read of a dead region poisons the result.

This is not a restatement of reference code (the reference delegates all of this to
Umfpack/CHOLMOD); the parity anchor for numerics is oracle/sleqp_oracle.py.
"""
from __future__ import annotations

import numpy as np

NB, RB, TILE, EA_COLS = 32, 128, 64, 4


def _ld(h):
    """Leading dimension of a column-major panel: the front height rounded up to even (plan.hpp panel_ld)."""
    return (h + 1) & ~1
UPD_INPANEL, UPD_SCHUR, UPD_DIAGCOPY = 0, 1, 2


class Emulated:
    def __init__(self, plan: dict, kval: np.ndarray):
        self.p = plan
        self.kval = np.asarray(kval, dtype=np.float64)
        self.m = int(plan["n_reduced"])
        self.nE = int(plan["n_elim"])
        self.N = int(plan["n"])
        self.n_perturbed = 0
        self.L11 = {}  # unit lower factors of the diagonal blocks, keyed by (supernode, panel step)
        self._read_sst()
        self._factor()
        self._invert()

    # ---------------------------------------------------------------- sparse subtrees (sst.cu)
    def _read_sst(self):
        """SstMeta records of the plan (plan.hpp): supernodes kept with their exact sparse structure, sorted by
        generation; the index structure is read from the 16-bit blob the device stages (sst.cu: sst_stage)."""
        p = self.p
        ns = int(p["n_supernodes"])
        self.is_sst = np.zeros(ns, dtype=bool)
        self.sst = []
        raw = np.asarray(p.get("sst", np.zeros((0, 30), dtype=np.int32))).reshape(-1, 30)
        blob_all = np.asarray(p.get("sst_blob", np.zeros(0, dtype=np.int32)), dtype=np.int64)
        last_gen = 0
        for rec in raw:
            lptr = int(np.array(rec[0:2], dtype=np.int32).view(np.int64)[0])
            uoff = int(np.array(rec[2:4], dtype=np.int32).view(np.int64)[0])
            (sn, first, k, r, rptr, signal, blob, blob_len16, nslev, nseg, nnz, gen, o_segstart, o_seglen, o_colptr, o_rows, ea_begin, ea_end,
             col_ptr, row_ptr, nchild, parent_sst, o_rowptr, o_rcol, o_rpos, _pad) = (int(v) for v in rec[4:30])
            assert blob % 8 == 0 and gen >= last_gen
            last_gen = gen
            b = blob_all[blob : blob + 8 * blob_len16]
            slvl = b[: nslev + 1]
            segstart = b[o_segstart : o_segstart + nseg]
            seglen = b[o_seglen : o_seglen + nseg]
            colptr = b[o_colptr : o_colptr + k + 1]
            rows = b[o_rows : o_rows + nnz]
            assert slvl[0] == 0 and slvl[-1] == nseg and colptr[0] == 0 and colptr[-1] == nnz
            rowptr = b[o_rowptr : o_rowptr + k + r + 1]
            rcol = b[o_rcol : o_rcol + nnz - k]
            rpos = b[o_rpos : o_rpos + nnz - k]
            assert rowptr[0] == 0 and rowptr[-1] == nnz - k
            # the row view holds exactly the off-diagonal entries, every row with ascending columns
            seen = np.zeros(nnz, dtype=bool)
            for i in range(k + r):
                cs = rcol[rowptr[i] : rowptr[i + 1]]
                assert np.all(np.diff(cs) > 0) and np.all(rows[rpos[rowptr[i] : rowptr[i + 1]]] == i)
                assert np.all((colptr[cs] < rpos[rowptr[i] : rowptr[i + 1]]) & (rpos[rowptr[i] : rowptr[i + 1]] < colptr[cs + 1]))
                seen[rpos[rowptr[i] : rowptr[i + 1]]] = True
            assert seen.sum() == nnz - k and not seen[colptr[:-1]].any()
            assert np.array_equal(colptr, p["sst_colptr"][col_ptr : col_ptr + k + 1])
            assert np.array_equal(rows, p["sst_rows"][row_ptr : row_ptr + nnz])
            assert sorted(np.concatenate([np.arange(s0, s0 + n) for s0, n in zip(segstart, seglen)]).tolist()) == list(range(k))
            seg_lev = np.zeros(k, dtype=np.int64)  # segment level of every column
            seg_end = np.zeros(k, dtype=np.int64)
            for lev in range(nslev):
                for q in range(slvl[lev], slvl[lev + 1]):
                    seg_lev[segstart[q] : segstart[q] + seglen[q]] = lev
                    seg_end[segstart[q] : segstart[q] + seglen[q]] = segstart[q] + seglen[q]
            assert p["sn_parent"][sn] == signal and nchild == p["child_ptr"][sn + 1] - p["child_ptr"][sn]
            assert parent_sst == (signal if signal >= 0 and p["sn_sparse"][signal] else -1)
            self.sst.append(dict(Lptr=lptr, Uoff=uoff, sn=sn, first=first, k=k, r=r, Rptr=rptr, signal=signal, colptr=colptr, rows=rows,
                                 slvl=slvl, segstart=segstart, seglen=seglen, nslev=nslev, nnz=nnz, gen=gen, seg_lev=seg_lev, seg_end=seg_end,
                                 ea=(ea_begin, ea_end), nchild=nchild, parent_sst=parent_sst, rowptr=rowptr, rcol=rcol, rpos=rpos))
            self.is_sst[sn] = True
            assert p["sn_sparse"][sn] == 1

    def _sst_segments(self, M, lev):
        return [(int(M["segstart"][q]), int(M["segstart"][q] + M["seglen"][q])) for q in range(M["slvl"][lev], M["slvl"][lev + 1])]

    def _sst_factor(self, M):
        """k_sst_factor: extend-add of the child subtrees, then right-looking sparse LDL^T; one thread per segment,
        segments of one level at a time: a target outside the thread's own segment must belong to a later level."""
        vals = self.L[M["Lptr"] : M["Lptr"] + M["nnz"]]
        k, r, colptr, rows = M["k"], M["r"], M["colptr"], M["rows"]
        Us = np.zeros(r * r)  # column-major like the device block
        for e in range(*M["ea"]):
            dst, v = int(self.p["sst_ea_dst"][e]), self.U[int(self.p["sst_ea_src"][e])]
            if dst >= 0:
                assert dst < M["nnz"]
                vals[dst] += v
            else:
                Us[-1 - dst] += v
        for lev in range(M["nslev"]):
            for j0, j1 in self._sst_segments(M, lev):
                for j in range(j0, j1):
                    p0, p1 = colptr[j], colptr[j + 1]
                    assert rows[p0] == j
                    d = vals[p0]
                    if not (abs(d) >= self.tau) or not np.isfinite(d):
                        d = -self.tau if self.tau > 0 else -1e-300
                        self.n_perturbed += 1
                    self.D[M["first"] + j] = d
                    vals[p0] = d
                    for a in range(p0 + 1, p1):
                        la = vals[a] / d
                        ia = rows[a]
                        for b in range(p0 + 1, a + 1):
                            ib, u = rows[b], -la * vals[b]
                            if ib >= k:
                                Us[(ia - k) + (ib - k) * r] += u
                            else:
                                assert (j < ib < j1) or M["seg_lev"][ib] > lev
                                t = colptr[ib] + int(np.nonzero(rows[colptr[ib] : colptr[ib + 1]] == ia)[0][0])
                                vals[t] += u
                    vals[p0 + 1 : p1] /= d
        if r:
            self.U[M["Uoff"] : M["Uoff"] + r * r] = Us

    def _sst_forward(self, M, yacc, yf):
        """k_sst_forward: every column GATHERS along its row (no atomics inside a subtree): x_j = b_j - sum_c l_jc x_c over
        the entries of row j, which belong to earlier columns of the thread's own segment or to segments of lower levels;
        the tail rows gather after the last level."""
        k, r = M["k"], M["r"]
        rowptr, rcol, rpos = M["rowptr"], M["rcol"], M["rpos"]
        vals = self.L[M["Lptr"] : M["Lptr"] + M["nnz"]]
        x = np.concatenate([yacc[M["first"] : M["first"] + k], np.zeros(r)])
        for lev in range(M["nslev"]):
            for j0, j1 in self._sst_segments(M, lev):
                for j in range(j0, j1):
                    for e in range(rowptr[j], rowptr[j + 1]):
                        c = rcol[e]
                        assert (j0 <= c < j) or M["seg_lev"][c] < lev  # written by this thread or before the last barrier
                        x[j] -= vals[rpos[e]] * x[c]
        for i in range(k, k + r):
            x[i] = -sum(vals[rpos[e]] * x[rcol[e]] for e in range(rowptr[i], rowptr[i + 1]))
        yf[M["first"] : M["first"] + k] = x[:k] / self.D[M["first"] : M["first"] + k]
        tail = self.p["Ridx"][M["Rptr"] : M["Rptr"] + r]
        np.add.at(yacc, tail, x[k:])

    def _sst_backward(self, M, yf, xg):
        k, r, colptr, rows = M["k"], M["r"], M["colptr"], M["rows"]
        vals = self.L[M["Lptr"] : M["Lptr"] + M["nnz"]]
        tail = self.p["Ridx"][M["Rptr"] : M["Rptr"] + r]
        x = np.concatenate([yf[M["first"] : M["first"] + k], xg[tail]])
        for lev in range(M["nslev"] - 1, -1, -1):
            for j0, j1 in self._sst_segments(M, lev):
                for j in range(j1 - 1, j0 - 1, -1):
                    for a in range(colptr[j] + 1, colptr[j + 1]):
                        i = rows[a]
                        assert (j < i < j1) or i >= k or M["seg_lev"][i] > lev
                    x[j] -= sum(vals[a] * x[rows[a]] for a in range(colptr[j] + 1, colptr[j + 1]))
        xg[M["first"] : M["first"] + k] = x[:k]

    # ---------------------------------------------------------------- geometry helpers
    def _geom(self, T):
        p = self.p
        f = int(p["sn_first"][T])
        k = int(p["sn_first"][T + 1]) - f
        r = int(p["Rptr"][T + 1] - p["Rptr"][T])
        return f, k, r, k + r

    def panel(self, T):
        """h x k column-major view of supernode T's panel."""
        f, k, r, h = self._geom(T)
        o, ld = int(self.p["Lptr"][T]), _ld(h)
        return self.L[o : o + ld * k].reshape((k, ld)).T[:h]

    def umat(self, T):
        f, k, r, h = self._geom(T)
        o = int(self.p["Uoff"][T])
        return self.U[o : o + r * r].reshape((r, r)).T

    # ---------------------------------------------------------------- numeric
    def _factor(self):
        p, kv = self.p, self.kval
        nS = len(p["Sdest"])
        self.L = np.zeros(int(p["Lptr"][-1]))
        self.U = np.full(int(p["update_ws_doubles"]), np.nan)
        self.D = np.zeros(self.m)
        # assembly of S = G - A D^-1 A^T into the panels
        val = np.zeros(nS)
        g = p["Sgsrc"]
        val[g >= 0] = kv[g[g >= 0]]
        nt = len(p["Sterm_a"])
        if nt:
            prod = kv[p["Sterm_a"]] * kv[p["Sterm_b"]] / kv[p["Sterm_d"]]
            ent = np.repeat(np.arange(nS), np.diff(p["Sterm_ptr"]))
            np.subtract.at(val, ent, prod)
        self.L[p["Sdest"]] = val
        # static pivot threshold from the largest |S_jj|
        first_ent = np.zeros(self.m, dtype=np.int64)
        # diagonal entry is the first entry of every column of S; recover via Sdest of (j,j)
        smax = 0.0
        for T in range(int(p["n_supernodes"])):
            if self.is_sst[T]:
                continue
            f, k, r, h = self._geom(T)
            smax = max(smax, np.abs(np.diag(self.panel(T)[:k, :k])).max(initial=0.0))
        for M in self.sst:
            smax = max(smax, np.abs(self.L[M["Lptr"] + M["colptr"][:-1]]).max(initial=0.0))
        self.tau = (2.0 ** -26 if int(p.get("n_demoted", 0)) > 0 else 64 * np.finfo(float).eps) * smax  # plan.hpp STATIC_PIVOT_*
        self.n_perturbed = getattr(self, "n_perturbed", 0)
        for M in self.sst:  # leaves of the supernodal tree: before the first stage
            self._sst_factor(M)
        scratch = np.full(max(1, int(p["n_scratch_slots"]) if "n_scratch_slots" in p else 1) * NB * NB, np.nan)
        has_children = np.diff(p["child_ptr"]) > 0
        # Look-ahead on the device (two streams): per stage the panel step and the update tiles of the NEXT panel's
        # columns [ub, um) on the main stream, the remaining tiles [um, ue) on the side stream next to the following
        # panel step. rest(s) waits for panel(s); the look-ahead tiles of stage s and any extend-add wait for
        # rest(s - 1). The emulation runs the LATEST order that allows: rest(s - 1) after panel(s).
        pending = None  # rest tiles of the previous stage not yet executed

        def run_updates(lo, hi):
            for T, t, kind, i0, j0, kb, ke in p["upd_tasks"][lo:hi]:
                self._update(int(T), int(t), int(kind), int(i0), int(j0), scratch, bool(has_children[T]), int(kb), int(ke))

        for st in p["stages"]:
            zb, ze, eb, ee, pb, pe, ub, um, ue, _pad = (int(x) for x in st)
            assert ub <= um <= ue
            if pending is not None and (ee > eb or ze > zb):
                run_updates(*pending)
                pending = None
            for T in p["zero_sn"][zb:ze]:
                self.umat(T)[:, :] = 0.0
            for c, jb in p["ea_tasks"][eb:ee]:
                self._extend_add(int(c), int(jb))
            for T, t, rb, _pad2 in p["pan_tasks"][pb:pe]:
                self._panel(int(T), int(t), int(rb))
            if pending is not None:
                run_updates(*pending)
            run_updates(ub, um)
            pending = (um, ue)
        if pending is not None:
            run_updates(*pending)

    # ---------------------------------------------------------------- selective inversion
    def mpanel(self, T):
        f, k, r, h = self._geom(T)
        o, ld = int(self.p["Lptr"][T]), _ld(h)
        return self.Mt[o : o + ld * k].reshape((k, ld)).T[:h]

    def _invert(self):
        """Minv = [L11^-1; -L21 L11^-1] per supernode, executed tile task by tile task."""
        p = self.p
        self.Mt = np.zeros_like(self.L)
        tmp = np.full(max(1, int(p["Tptr"][-1])), np.nan)
        ns = int(p["n_supernodes"])
        for T in range(ns):  # what k_panel publishes: inverse of every NB x NB diagonal block
            if self.is_sst[T]:
                continue  # sparse subtrees keep their factor (substitution instead of inverse panels)
            f, k, r, h = self._geom(T)
            P, M = self.panel(T), self.mpanel(T)
            for c0 in range(0, k, NB):
                w = min(NB, k - c0)
                L11 = self.L11[(T, c0 // NB)]
                M[c0 : c0 + w, c0 : c0 + w] = np.tril(np.linalg.inv(L11))
        ph = p["inv_phase_ptr"]
        for q in range(len(ph) - 1):
            for T, kind, i0, j0, kb, ke in p["inv_tasks"][int(ph[q]) : int(ph[q + 1])]:
                T, kind, i0, j0, kb, ke = int(T), int(kind), int(i0), int(j0), int(kb), int(ke)
                f, k, r, h = self._geom(T)
                P, M = self.panel(T), self.mpanel(T)
                Tm = tmp[int(p["Tptr"][T]) : int(p["Tptr"][T]) + k * k].reshape((k, k)).T if k > NB else None
                if kind == 0:  # T1: Tmp = L11[C, A] * Ainv ; A columns end at ke
                    i1, j1 = min(i0 + TILE, k), min(j0 + TILE, ke)
                    Tm[i0:i1, j0:j1] = P[i0:i1, kb:ke] @ M[kb:ke, j0:j1]
                elif kind == 1:  # T2: Minv[C, A] = -Cinv * Tmp ; C starts at kb (= end of the A columns)
                    i1, j1 = min(i0 + TILE, k), min(j0 + TILE, kb)
                    M[i0:i1, j0:j1] = -(M[i0:i1, kb:ke] @ Tm[kb:ke, j0:j1])
                else:  # Z: Minv[k + i, j] = -sum_q L21[i, q] Linv[q, j]
                    i1, j1 = min(i0 + TILE, r), min(j0 + TILE, k)
                    M[k + i0 : k + i1, j0:j1] = -(P[k + i0 : k + i1, kb:ke] @ M[kb:ke, j0:j1])

    def row_major_copy(self):
        """Mr as k_transpose leaves it: only the 32x32 tiles of tr_tasks are written, the rest of the buffer is
        whatever cudaMalloc returned (NaN here)."""
        p = self.p
        Mr = np.full(self.Mt.shape, np.nan)
        for T, i0, j0 in p["tr_tasks"]:
            T, i0, j0 = int(T), int(i0), int(j0)
            f, k, r, h = self._geom(T)
            o = int(p["Lptr"][T])
            src = self.Mt[o : o + _ld(h) * k].reshape((k, _ld(h))).T[:h]  # h x k view of the column-major panel
            dst = Mr[o : o + h * k].reshape((h, k))
            i1, j1 = min(h, i0 + 32), min(k, j0 + 32)
            dst[i0:i1, j0:j1] = src[i0:i1, j0:j1]
        return Mr

    def solve_reduced_flow(self, b_new):
        """The dataflow device solve (solve.cu k_flow): warp tasks executed one at a time in ticket order. Asserts
        what the kernels rely on: every counter a task waits for has already reached its target when the tasks run
        sequentially in list order (the order the ticket counters hand them out in, shard by shard => no deadlock for
        any number of resident warps), and the raw panels (no masking of the entries above the diagonal; the row-major copy only
        where k_transpose wrote it) give the right answer."""
        p = self.p
        ns = int(p["n_supernodes"])
        Ridx, Mt = p["Ridx"], self.Mt
        Mr = self.row_major_copy()
        Dinv = 1.0 / self.D

        def fields(t):
            lptr = int(np.array(t[0:2], dtype=np.int32).view(np.int64)[0])
            return (lptr,) + tuple(int(v) for v in t[2:13])

        yacc = np.array(b_new, dtype=np.float64)
        yf = np.zeros(self.m)
        x = np.zeros(self.m)
        cnt = np.zeros(ns, dtype=np.int64)
        for M in self.sst:  # k_sst_forward runs before the dataflow kernel and signals the parents
            assert cnt[M["sn"]] == M["nchild"], "sparse subtree swept before its children"
            self._sst_forward(M, yacc, yf)
            if M["signal"] >= 0:
                cnt[M["signal"]] += 1
        tasks = p["ffl_tasks"]
        for t in range(len(tasks)):
            lptr, rptr, first, k, h, i0, i1, j0, j1, wait_idx, need, signal_idx = fields(tasks[t])
            assert 0 < i1 - i0 <= 32 and 0 < j1 - j0 <= 128 and i1 <= h and j1 <= k
            if wait_idx >= 0:
                assert cnt[wait_idx] == need, "forward task claimed before its producers"
            b = yacc[first + j0 : first + j1]
            for r in range(i0, i1):
                acc = sum(Mt[lptr + j * h + r] * b[j - j0] for j in range(j0, j1))
                if r < k:
                    yf[first + r] += acc * Dinv[first + r]
                else:
                    yacc[Ridx[rptr + r - k]] += acc
            if signal_idx >= 0:
                cnt[signal_idx] += 1
        cnt[:] = 0
        tasks = p["bfl_tasks"]
        for t in range(len(tasks)):
            lptr, rptr, first, k, h, i0, i1, j0, j1, wait_idx, need, signal_idx = fields(tasks[t])
            assert 0 < j1 - j0 <= 32 and j0 % 32 == 0 and 0 < i1 - i0 <= 128 and i0 % 16 == 0 and j1 <= k and i1 <= h
            if wait_idx >= 0:
                assert cnt[wait_idx] == need, "backward task claimed before its producers"
            else:
                assert i1 <= k or int(p["sn_parent"][int(p["sn_of_col"][first])]) < 0
            ii = np.arange(i0, i1)
            v = np.array([yf[first + i] if i < k else x[Ridx[rptr + i - k]] for i in ii])
            for j in range(j0, j1):
                x[first + j] += Mr[lptr + ii * k + j] @ v
            if signal_idx >= 0:
                cnt[signal_idx] += 1
        swept = set()
        for M in reversed(self.sst):  # k_sst_backward: after the dataflow kernel, generations from the top down
            assert M["parent_sst"] < 0 or M["parent_sst"] in swept, "sparse subtree swept before its parent"
            self._sst_backward(M, yf, x)
            swept.add(M["sn"])
        assert np.all(np.isfinite(x))
        return x

    def _extend_add(self, c, jb):
        p = self.p
        par = int(p["sn_parent"][c])
        fc, kc, rc, hc = self._geom(c)
        fp, kp, rp, hp = self._geom(par)
        rel = p["rel"][int(p["Rptr"][c]) : int(p["Rptr"][c + 1])]
        Uc = self.umat(c)
        Lp = self.panel(par)
        Up = self.umat(par) if rp > 0 else None
        for j in range(jb * EA_COLS, min(rc, (jb + 1) * EA_COLS)):
            pj = int(rel[j])
            for i in range(j, rc):
                pi = int(rel[i])
                if pj < kp:
                    Lp[pi, pj] += Uc[i, j]
                else:
                    Up[pi - kp, pj - kp] += Uc[i, j]

    def _factor_diag(self, A):
        """LDL^T of a dense w x w block (lower part of A), returns (unit-lower L, d)."""
        w = A.shape[0]
        A = np.tril(A).copy()
        d = np.zeros(w)
        nper = 0
        for j in range(w):
            dj = A[j, j]
            if abs(dj) < self.tau or not np.isfinite(dj):
                dj = -self.tau if self.tau > 0 else -1e-300
                nper += 1
            d[j] = dj
            A[j, j] = 1.0
            A[j + 1 :, j] /= dj
            for c in range(j + 1, w):
                A[c:, c] -= A[c:, j] * dj * A[c, j]
        return A, d, nper

    def _panel(self, T, t, rb):
        """k_panel: every CTA of a panel step factors the (still unfactored, read-only) diagonal block itself and
        solves for its RB rows of L21; CTA 0 publishes the pivots and the unit lower factor (for the inversion)."""
        f, k, r, h = self._geom(T)
        P = self.panel(T)
        c0 = t * NB
        w = min(NB, k - c0)
        L11, d, nper = self._factor_diag(P[c0 : c0 + w, c0 : c0 + w])
        r0 = c0 + w + rb * RB
        r1 = min(h, r0 + RB)
        if r1 > r0:
            X = P[r0:r1, c0 : c0 + w].copy()
            # solve X_new * D * L11^T = X
            for j in range(w):
                X[:, j] = (X[:, j] - (X[:, :j] * d[:j]) @ L11[j, :j]) / d[j]
            P[r0:r1, c0 : c0 + w] = X
        if rb == 0:
            self.n_perturbed += nper
            self.D[f + c0 : f + c0 + w] = d
            self.L11[(T, t)] = np.tril(L11, -1) + np.eye(w)

    def _update(self, T, t, kind, i0, j0, scratch, accumulate, kb=None, ke=None):
        f, k, r, h = self._geom(T)
        P = self.panel(T)
        if kind == UPD_DIAGCOPY:
            slot = i0
            c0 = t * NB
            w = min(NB, k - c0)
            L11 = scratch[slot * NB * NB : slot * NB * NB + w * w].reshape((w, w)).T
            il = np.tril_indices(w)
            blk = P[c0 : c0 + w, c0 : c0 + w]
            blk[il] = L11[il]
            return
        if kind == UPD_INPANEL:
            d = self.D[f + kb : f + ke]
            i1, j1 = min(h, i0 + TILE), min(t, j0 + TILE)  # t carries the tile's column limit
            Li = P[i0:i1, kb:ke]
            Lj = P[j0:j1, kb:ke]
            upd = (Li * d) @ Lj.T
            mask = (np.arange(i0, i1)[:, None] >= np.arange(j0, j1)[None, :])
            blk = P[i0:i1, j0:j1]
            blk[mask] -= upd[mask]
            return
        if kind == UPD_SCHUR:
            d = self.D[f : f + k]
            Um = self.umat(T)
            sh = k & 1  # shifted tile grid: Schur tiles start at even front rows (symbolic.cpp)
            i1, j1 = min(r, i0 - sh + TILE), min(r, j0 - sh + TILE)
            i0, j0 = max(0, i0 - sh), max(0, j0 - sh)
            Li = P[k + i0 : k + i1, :k]
            Lj = P[k + j0 : k + j1, :k]
            upd = (Li * d) @ Lj.T
            mask = (np.arange(i0, i1)[:, None] >= np.arange(j0, j1)[None, :])
            blk = Um[i0:i1, j0:j1]
            if accumulate:
                blk[mask] -= upd[mask]
            else:
                blk[mask] = -upd[mask]
            return
        raise ValueError(kind)

    # ---------------------------------------------------------------- solve
    def _L11_full(self, T):
        """Unit lower k x k factor of supernode T: off-diagonal blocks from the panel, diagonal blocks as published by
        CTA 0 of every panel step (the panel keeps the assembled values there)."""
        f, k, r, h = self._geom(T)
        out = np.tril(self.panel(T)[:k, :k], -1) + np.eye(k)
        for c0 in range(0, k, NB):
            w = min(NB, k - c0)
            out[c0 : c0 + w, c0 : c0 + w] = self.L11[(T, c0 // NB)]
        return out

    def solve_reduced(self, b_new):
        """Solve S x = b in the permuted (new) labels with the multifrontal front vectors."""
        p = self.p
        ns = int(p["n_supernodes"])
        W = np.full(int(p["Wptr"][-1]), np.nan)
        x = np.zeros(self.m)
        for lv in range(int(p["n_levels"])):
            for T in p["lvl_sn"][int(p["lvl_ptr"][lv]) : int(p["lvl_ptr"][lv + 1])]:
                T = int(T)
                f, k, r, h = self._geom(T)
                w = W[int(p["Wptr"][T]) : int(p["Wptr"][T]) + h]
                w[:k] = b_new[f : f + k]
                w[k:] = 0.0
                for c in p["child_idx"][int(p["child_ptr"][T]) : int(p["child_ptr"][T + 1])]:
                    c = int(c)
                    fc, kc, rc, hc = self._geom(c)
                    rel = p["rel"][int(p["Rptr"][c]) : int(p["Rptr"][c + 1])]
                    wc = W[int(p["Wptr"][c]) + kc : int(p["Wptr"][c]) + hc]
                    np.add.at(w, rel, wc)
                if self.is_sst[T]:
                    M = next(q for q in self.sst if q["sn"] == T)
                    vals = self.L[M["Lptr"] : M["Lptr"] + M["nnz"]]
                    for j in range(k):  # column order = a topological order of the subtree
                        for a in range(M["colptr"][j] + 1, M["colptr"][j + 1]):
                            w[M["rows"][a]] -= vals[a] * w[j]
                    x[f : f + k] = w[:k]
                    continue
                P = self.panel(T)
                L11 = self._L11_full(T)
                y = np.linalg.solve(L11, w[:k])
                w[:k] = y
                w[k:] -= P[k:, :k] @ y
                x[f : f + k] = y
        x /= self.D
        for lv in range(int(p["n_levels"]) - 1, -1, -1):
            for T in p["lvl_sn"][int(p["lvl_ptr"][lv]) : int(p["lvl_ptr"][lv + 1])]:
                T = int(T)
                f, k, r, h = self._geom(T)
                rows = p["Ridx"][int(p["Rptr"][T]) : int(p["Rptr"][T + 1])]
                if self.is_sst[T]:
                    M = next(q for q in self.sst if q["sn"] == T)
                    vals = self.L[M["Lptr"] : M["Lptr"] + M["nnz"]]
                    loc = np.concatenate([x[f : f + k], x[rows]])
                    for j in range(k - 1, -1, -1):
                        loc[j] -= sum(vals[a] * loc[M["rows"][a]] for a in range(M["colptr"][j] + 1, M["colptr"][j + 1]))
                    x[f : f + k] = loc[:k]
                    continue
                P = self.panel(T)
                L11 = self._L11_full(T)
                t = x[f : f + k] - P[k:, :k].T @ x[rows]
                x[f : f + k] = np.linalg.solve(L11.T, t)
        return x

    def solve(self, rhs, refine=1, flow=True):
        """Full K solve in original K indices, block elimination around the reduced system."""
        p, kv = self.p, self.kval
        ke, kr = p["k_of_e"], p["k_of_r"]
        dE = kv[p["dE_src"]]
        import scipy.sparse as sp

        A = sp.csr_matrix((kv[p["Acsr_src"]], p["Acsr_col"], p["Acsr_ptr"]), shape=(self.m, self.nE))
        G = sp.csr_matrix((kv[p["Gsym_src"]], p["Gsym_col"], p["Gsym_ptr"]), shape=(self.m, self.m))

        def once(b):
            t = b[ke] / dE
            bR = b[kr] - A @ t
            lam_new = (self.solve_reduced_flow if flow else self.solve_reduced)(bR[p["perm"]])
            lam = np.empty(self.m)
            lam[p["perm"]] = lam_new
            z = np.empty(self.N)
            z[kr] = lam
            z[ke] = t - (A.T @ lam) / dE
            return z

        def kmul(z):
            out = np.empty(self.N)
            out[ke] = dE * z[ke] + A.T @ z[kr]
            out[kr] = A @ z[ke] + G @ z[kr]
            return out

        z = once(rhs)
        for _ in range(refine):
            z = z + once(rhs - kmul(z))
        return z

    def pivots_full(self):
        """D in factorization order of K: [d_E | D_S]."""
        return np.concatenate([self.kval[self.p["dE_src"]], self.D])
