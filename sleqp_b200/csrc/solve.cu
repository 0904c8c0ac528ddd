// solve.cu -- the repeated solves K z = b behind SleqpFactCallbacks.solve (fact_types.h:12;
// vendor calls replaced: umfpack_di_solve fact_umfpack.c:220, cholmod_l_solve fact_cholmod.c:184).
//
//   K = [D A^T; A G]:  t = b_E / d,  b_R' = b_R - A t,  S y = b_R',  z_R = y,  z_E = (b_E - A^T y) / d
// The reduced system is solved with level-scheduled supernodal forward / diagonal / backward
// sweeps over the inverse panels (selective inversion): every supernode step is a matrix-vector product.
// Iterative refinement runs against the unperturbed K (SURVEY.md hard part 1).
#include "numeric.cuh"

#include <cooperative_groups.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace b200
{

constexpr int SOLVE_THREADS = 128;

// ---- E-block elimination and back-substitution -------------------------------------------------
__global__ void
k_pre(int m,
      const int* __restrict__ k_of_r,
      const int* __restrict__ k_of_e,
      const int* __restrict__ pinv,
      const int* __restrict__ Aptr,
      const int* __restrict__ Acol,
      const double* __restrict__ Aval,
      const double* __restrict__ dE,
      const double* __restrict__ rhs,
      double* __restrict__ bR)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m)
  {
    return;
  }
  double acc = rhs[k_of_r[r]];
  for (int q = Aptr[r]; q < Aptr[r + 1]; ++q)
  {
    const int e = Acol[q];
    acc -= Aval[q] * (rhs[k_of_e[e]] / dE[e]);
  }
  bR[pinv[r]] = acc;
}

__global__ void
k_post_r(int m, const int* __restrict__ k_of_r, const int* __restrict__ pinv, const double* __restrict__ y, double* __restrict__ z)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m)
  {
    z[k_of_r[r]] = y[pinv[r]];
  }
}

__global__ void
k_post_e(int nE,
         const int* __restrict__ k_of_e,
         const int* __restrict__ pinv,
         const int* __restrict__ Aptr,
         const int* __restrict__ Arow,
         const double* __restrict__ Aval,
         const double* __restrict__ dE,
         const double* __restrict__ rhs,
         const double* __restrict__ y,
         double* __restrict__ z)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nE)
  {
    return;
  }
  const int k = k_of_e[e];
  double acc  = rhs[k];
  for (int q = Aptr[e]; q < Aptr[e + 1]; ++q)
  {
    acc -= Aval[q] * y[pinv[Arow[q]]];
  }
  z[k] = acc / dE[e];
}

// res = rhs - K z
__global__ void
k_resid_e(int nE,
          const int* __restrict__ k_of_e,
          const int* __restrict__ k_of_r,
          const int* __restrict__ Aptr,
          const int* __restrict__ Arow,
          const double* __restrict__ Aval,
          const double* __restrict__ dE,
          const double* __restrict__ rhs,
          const double* __restrict__ z,
          double* __restrict__ res)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nE)
  {
    return;
  }
  const int k = k_of_e[e];
  double acc  = rhs[k] - dE[e] * z[k];
  for (int q = Aptr[e]; q < Aptr[e + 1]; ++q)
  {
    acc -= Aval[q] * z[k_of_r[Arow[q]]];
  }
  res[k] = acc;
}

__global__ void
k_resid_r(int m,
          const int* __restrict__ k_of_e,
          const int* __restrict__ k_of_r,
          const int* __restrict__ Aptr,
          const int* __restrict__ Acol,
          const double* __restrict__ Aval,
          const int* __restrict__ Gptr,
          const int* __restrict__ Gcol,
          const double* __restrict__ Gval,
          const double* __restrict__ rhs,
          const double* __restrict__ z,
          double* __restrict__ res)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m)
  {
    return;
  }
  const int k = k_of_r[r];
  double acc  = rhs[k];
  for (int q = Aptr[r]; q < Aptr[r + 1]; ++q)
  {
    acc -= Aval[q] * z[k_of_e[Acol[q]]];
  }
  for (int q = Gptr[r]; q < Gptr[r + 1]; ++q)
  {
    acc -= Gval[q] * z[k_of_r[Gcol[q]]];
  }
  res[k] = acc;
}

__global__ void
k_axpy1(int n, const double* __restrict__ x, double* __restrict__ y)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    y[i] += x[i];
  }
}

// ---- forward sweep ---------------------------------------------------------------------------------
// One CTA per (supernode, row chunk). With the inverse panel Minv = [L11^-1; -L21 L11^-1] the whole
// supernode step is one matrix-vector product  [y_T; delta_tail] = Minv * b_T. The right-hand side lives in
// one global accumulator `yacc`: b_T = yacc[cols of T] already holds every update from T's descendants
// (they ran in earlier launches); tail results are pushed straight to their final rows with native FP64
// atomics (yacc[row] += delta), so there is no per-supernode front vector, no pass-through copying and the
// dependent-load chain of a CTA is task -> yacc -> panel. Rows of the top block go to `yf`.
// Reads the row-major copy Mr of the inverse panel: a warp owns RG rows at a time, lanes stride the
// (contiguous) columns with 8 independent loads in flight; the first round of panel loads is issued
// before b_T is staged. Dynamic shared memory: bT[k].
template <int RG, int U>
__device__ __forceinline__ void
fwd_body(const FwdTask& t, const int* __restrict__ Ridx, const double* __restrict__ Mr, const double* __restrict__ Dinv, double* __restrict__ yacc, double* __restrict__ yf, double* bT)
{
  const int k     = t.k;
  const double* P = Mr + t.Lptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0     = warp * RG;
  const int nvalid = min(RG, t.nrows - g0); // <= 0: this warp has no rows
  const int r0     = t.row0 + g0;
  // the top block is lower triangular: row r only needs columns <= r; rows of one group share the bound
  // of the last row (entries beyond a row's diagonal are zeros)
  const int rlast = r0 + nvalid - 1;
  const int jend  = nvalid <= 0 ? 0 : (rlast < k ? rlast + 1 : k);

  double pre[RG][U];
#pragma unroll
  for (int a = 0; a < RG; ++a)
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int j = lane + 32 * u;
      pre[a][u]   = (a < nvalid && j < jend) ? P[(long long)(r0 + a) * k + j] : 0.0;
    }
  // staged in batches of 4 independent loads per thread
  for (int j = tid; j < k; j += 4 * SOLVE_THREADS)
  {
    double tmp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int jj = j + u * SOLVE_THREADS;
      tmp[u]       = jj < k ? yacc[t.first + jj] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int jj = j + u * SOLVE_THREADS;
      if (jj < k)
      {
        bT[jj] = tmp[u];
      }
    }
  }
  __syncthreads();
  if (nvalid <= 0)
  {
    return; // no barrier follows inside this body
  }
  double acc[RG][U];
#pragma unroll
  for (int a = 0; a < RG; ++a)
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int j = lane + 32 * u;
      acc[a][u]   = j < jend ? pre[a][u] * bT[j] : 0.0;
    }
  int j = lane + 32 * U;
  for (; j + 32 * (U - 1) < jend; j += 32 * U)
  {
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double bj = bT[j + 32 * u];
#pragma unroll
      for (int a = 0; a < RG; ++a)
      {
        if (a < nvalid)
        {
          acc[a][u] += P[(long long)(r0 + a) * k + j + 32 * u] * bj;
        }
      }
    }
  }
  for (; j < jend; j += 32)
  {
    const double bj = bT[j];
#pragma unroll
    for (int a = 0; a < RG; ++a)
    {
      if (a < nvalid)
      {
        acc[a][0] += P[(long long)(r0 + a) * k + j] * bj;
      }
    }
  }
  double mine = 0.0;
#pragma unroll
  for (int a = 0; a < RG; ++a)
  {
    double v = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      v += acc[a][u];
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    if (lane == a)
    {
      mine = v;
    }
  }
  if (lane < nvalid)
  {
    const int r = r0 + lane;
    if (r < k)
    {
      yf[t.first + r] = mine * Dinv[t.first + r]; // D^-1 y: the diagonal solve is folded into the forward sweep
    }
    else
    {
      atomicAdd(yacc + Ridx[t.Rptr + r - k], mine);
    }
  }
}

// Wide fronts (k >= 512): the CTA owns RG rows and its four warps split the columns of each of them, so a CTA
// needs at most two rounds of loads; b_T is not staged, yacc is read directly (it is L2-resident: every CTA of
// the supernode reads the same k values).
template <int RG>
__device__ __forceinline__ void
fwd_wide(const FwdTask& t, const int* __restrict__ Ridx, const double* __restrict__ Mr, const double* __restrict__ Dinv, double* __restrict__ yacc, double* __restrict__ yf, double* red)
{
  constexpr int U  = 8 / RG;
  constexpr int NW = SOLVE_THREADS / 32;
  const int k      = t.k;
  const double* P  = Mr + t.Lptr;
  const double* bT = yacc + t.first;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvalid = min(RG, t.nrows);
  const int r0     = t.row0;
  const int rlast  = r0 + nvalid - 1;
  const int jend   = rlast < k ? rlast + 1 : k;
  double acc[RG];
#pragma unroll
  for (int a = 0; a < RG; ++a)
  {
    acc[a] = 0.0;
  }
  for (int j0 = 0; j0 < jend; j0 += SOLVE_THREADS * U)
  {
    double pv[RG][U], bv[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int j = j0 + u * SOLVE_THREADS + tid;
      bv[u]       = j < jend ? bT[j] : 0.0;
#pragma unroll
      for (int a = 0; a < RG; ++a)
      {
        pv[a][u] = (a < nvalid && j < jend) ? P[(long long)(r0 + a) * k + j] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int a = 0; a < RG; ++a)
      {
        acc[a] += pv[a][u] * bv[u];
      }
  }
#pragma unroll
  for (int a = 0; a < RG; ++a)
  {
    double v = acc[a];
    for (int o = 16; o > 0; o >>= 1)
    {
      v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    if (lane == 0)
    {
      red[a * NW + warp] = v;
    }
  }
  __syncthreads();
  if (tid < nvalid)
  {
    double mine = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      mine += red[tid * NW + w];
    }
    const int r = r0 + tid;
    if (r < k)
    {
      yf[t.first + r] = mine * Dinv[t.first + r];
    }
    else
    {
      atomicAdd(yacc + Ridx[t.Rptr + r - k], mine);
    }
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS)
k_fwd_chunk(const FwdTask* __restrict__ tasks, const int* __restrict__ Ridx, const double* __restrict__ Mr, const double* __restrict__ Dinv, double* __restrict__ yacc, double* __restrict__ yf)
{
  extern __shared__ double bT[];
  const FwdTask t  = tasks[blockIdx.x];
  constexpr int NW = SOLVE_THREADS / 32;
  if (t.wide)
  {
    if (t.nrows == 1)
    {
      fwd_wide<1>(t, Ridx, Mr, Dinv, yacc, yf, bT);
    }
    else
    {
      fwd_wide<2>(t, Ridx, Mr, Dinv, yacc, yf, bT);
    }
    return;
  }
  const int rg     = (t.nrows + NW - 1) / NW; // rows per warp: 1..4
  if (rg == 1)
  {
    fwd_body<1, 8>(t, Ridx, Mr, Dinv, yacc, yf, bT);
  }
  else if (rg == 2)
  {
    fwd_body<2, 4>(t, Ridx, Mr, Dinv, yacc, yf, bT);
  }
  else if (rg <= 4)
  {
    fwd_body<4, 2>(t, Ridx, Mr, Dinv, yacc, yf, bT);
  }
  else
  {
    fwd_body<8, 2>(t, Ridx, Mr, Dinv, yacc, yf, bT); // narrow supernodes: 8 rows per warp
  }
}

// ---- diagonal + backward sweep -----------------------------------------------------------------------
// One CTA per (supernode, column chunk):  x_T = Minv^T [D^-1 y_T; x_rows].  A warp owns CG columns at a
// time (rows are contiguous in the column-major panel), lanes stride the rows with 8 independent loads
// in flight (first round issued before the vector is gathered), shuffle reduction. Dynamic shared memory: v[h].
template <int CG, int U>
__device__ __forceinline__ void
bwd_body(const BwdTask& t,
         const int* __restrict__ Ridx,
         const double* __restrict__ Mt,
         const double* __restrict__ D,
         const double* __restrict__ y,
         double* __restrict__ x,
         double* v)
{
  const int k = t.k, h = t.h;
  const double* P = Mt + t.Lptr;
  const int* rows = Ridx + t.Rptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0     = warp * CG;
  const int nvalid = min(CG, t.ncols - g0);
  const int j0     = t.col0 + g0;
  // column j needs rows >= j; the columns of a group start at the first one's diagonal (the entries above a
  // diagonal are zeros)
  double pre[CG][U];
#pragma unroll
  for (int c = 0; c < CG; ++c)
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int i = j0 + lane + 32 * u;
      pre[c][u]   = (c < nvalid && i < h) ? P[(long long)(j0 + c) * h + i] : 0.0;
    }
  // v = [D^-1 y_T; x_rows], staged in batches of 4 independent (two-hop) loads per thread
  for (int i = t.col0 + tid; i < h; i += 4 * SOLVE_THREADS)
  {
    int src[4];
    double tmp[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int ii = i + u * SOLVE_THREADS;
      src[u]       = ii < k ? t.first + ii : (ii < h ? rows[ii - k] : 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int ii = i + u * SOLVE_THREADS;
      tmp[u]       = ii < k ? y[src[u]] : (ii < h ? x[src[u]] : 0.0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int ii = i + u * SOLVE_THREADS;
      if (ii < h)
      {
        v[ii] = tmp[u];
      }
    }
  }
  __syncthreads();
  if (nvalid <= 0)
  {
    return;
  }
  double acc[CG][U];
#pragma unroll
  for (int c = 0; c < CG; ++c)
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int i = j0 + lane + 32 * u;
      acc[c][u]   = i < h ? pre[c][u] * v[i] : 0.0;
    }
  int i = j0 + lane + 32 * U;
  for (; i + 32 * (U - 1) < h; i += 32 * U)
  {
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const double vi = v[i + 32 * u];
#pragma unroll
      for (int c = 0; c < CG; ++c)
      {
        if (c < nvalid)
        {
          acc[c][u] += P[(long long)(j0 + c) * h + i + 32 * u] * vi;
        }
      }
    }
  }
  for (; i < h; i += 32)
  {
    const double vi = v[i];
#pragma unroll
    for (int c = 0; c < CG; ++c)
    {
      if (c < nvalid)
      {
        acc[c][0] += P[(long long)(j0 + c) * h + i] * vi;
      }
    }
  }
  double mine = 0.0;
#pragma unroll
  for (int c = 0; c < CG; ++c)
  {
    double a = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      a += acc[c][u];
    }
    for (int o = 16; o > 0; o >>= 1)
    {
      a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    if (lane == c)
    {
      mine = a;
    }
  }
  if (lane < nvalid)
  {
    x[t.first + j0 + lane] = mine;
  }
}

// Tall fronts (h >= 512): the CTA owns CG columns and its four warps split the rows of each of them; the input
// vector is not staged: v_i = y_i (top block, already D^-1 y) or x[rows_i] is gathered alongside the panel loads.
template <int CG>
__device__ __forceinline__ void
bwd_tall(const BwdTask& t, const int* __restrict__ Ridx, const double* __restrict__ Mt, const double* __restrict__ y, double* __restrict__ x, double* red)
{
  constexpr int U  = 8 / CG;
  constexpr int NW = SOLVE_THREADS / 32;
  const int k = t.k, h = t.h;
  const double* P = Mt + t.Lptr;
  const int* rows = Ridx + t.Rptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nvalid = min(CG, -t.ncols);
  const int j0     = t.col0;
  double acc[CG];
#pragma unroll
  for (int c = 0; c < CG; ++c)
  {
    acc[c] = 0.0;
  }
  for (int i0 = j0; i0 < h; i0 += SOLVE_THREADS * U)
  {
    int src[U];
    double pv[CG][U], vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int i = i0 + u * SOLVE_THREADS + tid;
      src[u]      = i < k ? t.first + i : (i < h ? rows[i - k] : 0);
#pragma unroll
      for (int c = 0; c < CG; ++c)
      {
        pv[c][u] = (c < nvalid && i < h) ? P[(long long)(j0 + c) * h + i] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int i = i0 + u * SOLVE_THREADS + tid;
      vv[u]       = i < k ? y[src[u]] : (i < h ? x[src[u]] : 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int c = 0; c < CG; ++c)
      {
        acc[c] += pv[c][u] * vv[u];
      }
  }
#pragma unroll
  for (int c = 0; c < CG; ++c)
  {
    double v = acc[c];
    for (int o = 16; o > 0; o >>= 1)
    {
      v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    if (lane == 0)
    {
      red[c * NW + warp] = v;
    }
  }
  __syncthreads();
  if (tid < nvalid)
  {
    double mine = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w)
    {
      mine += red[tid * NW + w];
    }
    x[t.first + j0 + tid] = mine;
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS)
k_bwd_chunk(const BwdTask* __restrict__ tasks,
            const int* __restrict__ Ridx,
            const double* __restrict__ Mt,
            const double* __restrict__ D,
            const double* __restrict__ y,
            double* __restrict__ x)
{
  extern __shared__ double v[];
  const BwdTask t  = tasks[blockIdx.x];
  constexpr int NW = SOLVE_THREADS / 32;
  if (t.ncols < 0)
  {
    if (t.ncols == -1)
    {
      bwd_tall<1>(t, Ridx, Mt, y, x, v);
    }
    else
    {
      bwd_tall<2>(t, Ridx, Mt, y, x, v);
    }
    return;
  }
  const int cg     = (t.ncols + NW - 1) / NW; // columns per warp: 1..4
  if (cg == 1)
  {
    bwd_body<1, 8>(t, Ridx, Mt, D, y, x, v);
  }
  else if (cg == 2)
  {
    bwd_body<2, 4>(t, Ridx, Mt, D, y, x, v);
  }
  else
  {
    bwd_body<4, 2>(t, Ridx, Mt, D, y, x, v);
  }
}

// ---- dataflow sweeps -----------------------------------------------------------------------------------------
// One launch per sweep. Persistent warps draw SweepTasks (plan.hpp) by ticket in topological order and synchronise
// through per-supernode counters instead of kernel boundaries: the panel loads of a task are issued BEFORE its
// warp waits for the producers, so a level of the tree costs a counter poll + an L2 round trip for the vector, not
// a kernel drain + launch + three DRAM round trips.
//
// Both sweeps use the same mapping, "one lane per output, 16 panel entries in flight per lane, no shuffles":
//   forward  reads the column-major inverse panels Mt : lane = front row r,    out_r = sum_j Mt[j h + r] b_j
//   backward reads the row-major copy Mr              : lane = front column j, out_j = sum_i Mr[i k + j] v_i
// so a warp load is 32 consecutive doubles in either sweep. Partial sums go out as FP64 atomics.
//
// What the measurements on B200 decided (profiles/README.md): one ticket counter for all warps serialises the sweep
// (~0.5 same-address atomics with return per ns) -> sharded counters; drawing tickets ahead or in per-CTA batches
// keeps ready tasks hostage behind waiting warps (priority inversion, 30-90 % slower) -> one task per warp at a
// time; the draw of the next ticket is issued before the publication fence so the two round trips overlap.

__device__ __forceinline__ unsigned long long
global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Optional timeline of a sweep (B200_FLOW_TRACE=1, profile entry point only): per level the first claim, the first
// satisfied wait and the last finished task, plus the cycles the warps spent per phase of a task. Every warp
// records into its own rows (no shared words, so tracing does not serialise the sweep; it still costs three extra
// memory round trips per task); the host reduces over the warps. trace == nullptr in every product launch.
// Layout: rec[(warp * 3 + q) * nlevels + level], q = 0 first claim (min), 1 first ready (min), 2 last end (max),
// then 8 phase sums per warp.
struct FlowTrace
{
  unsigned long long* rec;
  int nlevels;
  int pad;
};

__device__ __forceinline__ void
trace_mark(const FlowTrace* tr, int q, int level)
{
  const int warp        = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned long long* p = tr->rec + ((size_t)warp * 3 + q) * tr->nlevels + level;
  const unsigned long long t = global_ns();
  if (q == 2 || t < *p)
  {
    *p = t;
  }
}

__device__ __forceinline__ long long
clk_after(double dep)
{
  long long c;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c) : "d"(dep) : "memory");
  return c;
}

struct FlowPhases
{
  long long fetch = 0, wait = 0, vec = 0, fma = 0, publish = 0, tasks = 0;
};

__device__ __forceinline__ int
ld_acquire(const int* p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ int
ld_relaxed(const int* p)
{
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// a release is all the counter hand-over needs: __threadfence() is fence.sc (MEMBAR.SC)
__device__ __forceinline__ void
fence_acq_rel()
{
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// All lanes poll the same word (one request per poll). While no producer has signalled yet the consumer is far
// from ready: it polls rarely and with relaxed loads (an acquire load invalidates the SM's whole L1, CCTL.IVALL, and
// thousands of warps hold tasks of the narrow top levels for most of the sweep). Once the count moves the
// producers are finishing: tight acquire polls, and the poll that succeeds is the acquire the vector loads need.
__device__ __forceinline__ void
wait_counter(const int* cnt, int need, unsigned far_sleep)
{
  int c = ld_acquire(cnt);
  while (c < need)
  {
    if (c == 0)
    {
      __nanosleep(far_sleep);
      c = ld_relaxed(cnt);
      if (c >= need)
      {
        c = ld_acquire(cnt);
      }
    }
    else
    {
      __nanosleep(32);
      c = ld_acquire(cnt);
    }
  }
}

// Tickets. The task list is dealt over FLOW_SHARDS interleaved sequences with one counter each (128 bytes apart:
// different L2 slices): shard s holds the tasks s, s + FLOW_SHARDS, s + 2 FLOW_SHARDS, ... Warp w of every CTA
// serves shard w, so ANY resident CTA serves all shards in order and the lowest unfinished task is always either
// running or about to be drawn by a free warp: no deadlock however few CTAs are resident. A warp whose shard is
// exhausted helps with the others.
constexpr int FLOW_SHARDS       = FLOW_THREADS / 32;
constexpr int FLOW_TICKET_PITCH = 32; // ints between two counters

struct FlowSched
{
  int ntasks;
  int far_sleep; // poll interval while no producer has signalled yet (ns)
};

// flow_take_issue only issues the atomic of the warp's current shard (its result stays in lane 0);
// flow_take_finish completes the draw, moving on to the other shards when that one is exhausted, and returns the
// task index or -1 when every shard is exhausted.
__device__ __forceinline__ int
flow_take_issue(int* __restrict__ ticket, int lane, int shard)
{
  return lane == 0 ? atomicAdd(ticket + shard * FLOW_TICKET_PITCH, 1) : 0;
}

__device__ __forceinline__ int
flow_take_finish(int issued, int* __restrict__ ticket, int ntasks, int lane, int& shard, int& tried)
{
  int t = -1;
  if (lane == 0)
  {
    int i = issued;
    for (;;)
    {
      const long long cand = (long long)i * FLOW_SHARDS + shard;
      if (cand < ntasks)
      {
        t = (int)cand;
        break;
      }
      shard = (shard + 1) % FLOW_SHARDS;
      if (++tried >= FLOW_SHARDS)
      {
        break;
      }
      i = atomicAdd(ticket + shard * FLOW_TICKET_PITCH, 1);
    }
  }
  return __shfl_sync(0xffffffffu, t, 0);
}

constexpr int FLOW_DEPTH = 16; // panel entries in flight per lane = depth of a task

// One task up to (not including) the publication of its completion.
//   FWD: lanes are rows [i0, i1), depth is columns [j0, j1), panel element (r, j) at Mt[j * h + r].
//   BWD: lanes are columns [j0, j1), depth is rows [i0, i1), panel element (i, j) at Mr[i * k + j].
// vsh: FLOW_DEPTH doubles of shared memory private to the warp (broadcast of the vector).
template <bool FWD, bool TRACE>
__device__ __forceinline__ void
flow_task(const SweepTask& T,
          int lane,
          const int* __restrict__ Ridx,
          const double* __restrict__ M,
          const double* __restrict__ Dinv,
          double* __restrict__ yacc,
          double* __restrict__ yf,
          double* __restrict__ x,
          const int* __restrict__ cnt,
          double* vsh,
          const FlowTrace* trace,
          unsigned far_sleep,
          FlowPhases& ph)
{
  long long c0 = 0, c1 = 0, c2 = 0;
  if (TRACE)
  {
    c0 = clk_after((double)T.k);
  }
  const int k = T.k, h = T.h;
  const int ld    = FWD ? h : k;
  const int o     = (FWD ? T.i0 : T.j0) + lane; // this lane's output index in the front
  const bool ov   = o < (FWD ? T.i1 : T.j1);
  const int d0    = FWD ? T.j0 : T.i0; // depth range
  const int nd    = (FWD ? T.j1 : T.i1) - d0;
  const double* P = M + T.Lptr + (long long)d0 * ld + (ov ? o : (FWD ? T.i0 : T.j0));
  double pre[FLOW_DEPTH];
#pragma unroll
  for (int u = 0; u < FLOW_DEPTH; ++u)
  {
    pre[u] = (ov && u < nd) ? __ldcs(P + (long long)u * ld) : 0.0; // streaming: must not evict the small hot arrays from L2
  }
  // where this lane's result goes, and where its share of the vector comes from (both independent of the producers)
  double* dst  = nullptr;
  double scale = 1.0;
  if (ov)
  {
    if (FWD)
    {
      if (o < k)
      {
        dst   = yf + T.first + o;
        scale = Dinv[T.first + o]; // D^-1 y: the diagonal solve is folded into the forward sweep
      }
      else
      {
        dst = yacc + Ridx[T.Rptr + o - k];
      }
    }
    else
    {
      dst = x + T.first + o;
    }
  }
  const double* vsrc = nullptr;
  bool vplain        = false; // written by an earlier kernel: no coherence concern
  if (lane < nd)
  {
    const int d = d0 + lane;
    if (FWD)
    {
      vsrc = yacc + T.first + d;
    }
    else if (d < k)
    {
      vsrc   = yf + T.first + d; // already D^-1 y
      vplain = true;
    }
    else
    {
      vsrc = x + Ridx[T.Rptr + d - k];
    }
  }
  if (TRACE && lane == 0)
  {
    trace_mark(trace, 0, T.pad0);
  }
  if (T.wait_idx >= 0)
  {
    wait_counter(cnt + T.wait_idx, T.need, far_sleep);
  }
  if (TRACE)
  {
    c1 = clk_after(0.0);
    if (lane == 0)
    {
      trace_mark(trace, 1, T.pad0);
    }
  }
  __syncwarp(); // the previous task's reads of vsh are done
  if (lane < FLOW_DEPTH)
  {
    vsh[lane] = vsrc ? (vplain ? *vsrc : __ldcg(vsrc)) : 0.0;
  }
  __syncwarp();
  if (TRACE)
  {
    c2 = clk_after(vsh[0]);
  }
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
  for (int u = 0; u < FLOW_DEPTH; u += 2)
  {
    acc0 += pre[u] * vsh[u];
    acc1 += pre[u + 1] * vsh[u + 1];
  }
  if (ov)
  {
    atomicAdd(dst, (acc0 + acc1) * scale);
  }
  if (TRACE)
  {
    const long long c3 = clk_after(acc0 + acc1);
    ph.wait += c1 - c0;
    ph.vec += c2 - c1;
    ph.fma += c3 - c2;
    ph.tasks += 1;
    if (lane == 0)
    {
      trace_mark(trace, 2, T.pad0);
    }
  }
}

template <bool FWD, bool TRACE>
__global__ void __launch_bounds__(FLOW_THREADS, FLOW_CTAS_PER_SM)
k_flow(const SweepTask* __restrict__ tasks,
       FlowSched sched,
       const int* __restrict__ Ridx,
       const double* __restrict__ M,
       const double* __restrict__ Dinv,
       double* __restrict__ yacc,
       double* __restrict__ yf,
       double* __restrict__ x,
       int* __restrict__ cnt,
       int* __restrict__ ticket,
       const FlowTrace* __restrict__ trace)
{
  __shared__ double vsh_all[FLOW_THREADS / 32 * FLOW_DEPTH];
  __shared__ __align__(16) SweepTask slot_all[FLOW_THREADS / 32]; // the warp's task record, fetched with cp.async
  const int lane  = threadIdx.x & 31;
  double* vsh     = vsh_all + (threadIdx.x >> 5) * FLOW_DEPTH;
  SweepTask* slot = slot_all + (threadIdx.x >> 5);
  FlowPhases ph;
  long long cf = TRACE ? clk_after(0.0) : 0;
  int shard = threadIdx.x >> 5, tried = 0;
  int cur   = flow_take_finish(flow_take_issue(ticket, lane, shard), ticket, sched.ntasks, lane, shard, tried);
  while (cur >= 0)
  {
    if (lane < 4)
    {
      const unsigned dst = (unsigned)__cvta_generic_to_shared((const char*)slot + 16 * lane);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"((const char*)(tasks + cur) + 16 * lane) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const SweepTask T = *slot;
    __syncwarp(); // the record is in registers: the slot may be overwritten
    if (TRACE)
    {
      ph.fetch += clk_after((double)T.h) - cf;
    }
    flow_task<FWD, TRACE>(T, lane, Ridx, M, Dinv, yacc, yf, x, cnt, vsh, trace, (unsigned)sched.far_sleep, ph);
    if (TRACE)
    {
      cf = clk_after(0.0);
    }
    // publication. The draw of the next ticket is issued first so that its round trip overlaps with the fence;
    // nothing can block between the draw and the signal.
    const bool open  = tried < FLOW_SHARDS;
    const int issued = open ? flow_take_issue(ticket, lane, shard) : 0;
    if (T.signal_idx >= 0)
    {
      fence_acq_rel(); // every lane: its own atomics are visible device-wide before the counter moves
      __syncwarp();
      if (lane == 0)
      {
        atomicAdd(cnt + T.signal_idx, 1);
      }
    }
    cur = open ? flow_take_finish(issued, ticket, sched.ntasks, lane, shard, tried) : -1;
    if (TRACE)
    {
      const long long c = clk_after((double)cur);
      ph.publish += c - cf;
      cf = c;
    }
  }
  if (TRACE && lane == 0)
  {
    const int warp        = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps      = (gridDim.x * blockDim.x) >> 5;
    unsigned long long* p = trace->rec + (size_t)nwarps * 3 * trace->nlevels + (size_t)warp * 8;
    p[0] = ph.fetch;
    p[1] = ph.wait;
    p[2] = ph.vec;
    p[3] = ph.fma;
    p[4] = ph.publish;
    p[5] = ph.tasks;
  }
}

// zeroes what the dataflow sweeps accumulate into (one thread per reduced row / counter)
__global__ void
k_flow_reset(int m, int nflow, double* __restrict__ yf, double* __restrict__ x, int* __restrict__ flow)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
  {
    yf[i] = 0.0;
    x[i]  = 0.0;
  }
  if (i < nflow)
  {
    flow[i] = 0;
  }
}

// ---- fused top of the tree ---------------------------------------------------------------------------
// The upper levels of the supernodal tree have few chunks each (tens to a few hundred CTAs) and cost a full
// kernel launch + drain (~13 us) per level. These cooperative kernels walk all of them in one launch with a
// grid barrier between levels; every CTA strides over the level's tasks.
__device__ __forceinline__ void
fwd_dispatch(const FwdTask& t, const int* __restrict__ Ridx, const double* __restrict__ Mr, const double* __restrict__ Dinv, double* __restrict__ yacc, double* __restrict__ yf, double* bT)
{
  constexpr int NW = SOLVE_THREADS / 32;
  if (t.wide)
  {
    if (t.nrows == 1)
    {
      fwd_wide<1>(t, Ridx, Mr, Dinv, yacc, yf, bT);
    }
    else
    {
      fwd_wide<2>(t, Ridx, Mr, Dinv, yacc, yf, bT);
    }
    return;
  }
  const int rg     = (t.nrows + NW - 1) / NW;
  if (rg == 1)
  {
    fwd_body<1, 8>(t, Ridx, Mr, Dinv, yacc, yf, bT);
  }
  else if (rg == 2)
  {
    fwd_body<2, 4>(t, Ridx, Mr, Dinv, yacc, yf, bT);
  }
  else if (rg <= 4)
  {
    fwd_body<4, 2>(t, Ridx, Mr, Dinv, yacc, yf, bT);
  }
  else
  {
    fwd_body<8, 2>(t, Ridx, Mr, Dinv, yacc, yf, bT); // narrow supernodes: 8 rows per warp
  }
}

__device__ __forceinline__ void
bwd_dispatch(const BwdTask& t, const int* __restrict__ Ridx, const double* __restrict__ Mt, const double* __restrict__ D, const double* __restrict__ y, double* __restrict__ x, double* v)
{
  constexpr int NW = SOLVE_THREADS / 32;
  if (t.ncols < 0)
  {
    if (t.ncols == -1)
    {
      bwd_tall<1>(t, Ridx, Mt, y, x, v);
    }
    else
    {
      bwd_tall<2>(t, Ridx, Mt, y, x, v);
    }
    return;
  }
  const int cg_    = (t.ncols + NW - 1) / NW;
  if (cg_ == 1)
  {
    bwd_body<1, 8>(t, Ridx, Mt, D, y, x, v);
  }
  else if (cg_ == 2)
  {
    bwd_body<2, 4>(t, Ridx, Mt, D, y, x, v);
  }
  else
  {
    bwd_body<4, 2>(t, Ridx, Mt, D, y, x, v);
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS)
k_fwd_top(const FwdTask* __restrict__ tasks, const int* __restrict__ lvl_ptr, int l0, int l1, const int* __restrict__ Ridx, const double* __restrict__ Mr, const double* __restrict__ Dinv, double* __restrict__ yacc, double* __restrict__ yf)
{
  extern __shared__ double smem_top[];
  cg::grid_group grid = cg::this_grid();
  for (int l = l0; l < l1; ++l)
  {
    for (int q = lvl_ptr[l] + blockIdx.x; q < lvl_ptr[l + 1]; q += gridDim.x)
    {
      const FwdTask t = tasks[q];
      fwd_dispatch(t, Ridx, Mr, Dinv, yacc, yf, smem_top);
      __syncthreads(); // the staging buffer is reused by the next task
    }
    grid.sync();
  }
}

__global__ void __launch_bounds__(SOLVE_THREADS)
k_bwd_top(const BwdTask* __restrict__ tasks, const int* __restrict__ lvl_ptr, int l0, int l1, const int* __restrict__ Ridx, const double* __restrict__ Mt, const double* __restrict__ D, const double* __restrict__ y, double* __restrict__ x)
{
  extern __shared__ double smem_top[];
  cg::grid_group grid = cg::this_grid();
  for (int l = l1 - 1; l >= l0; --l)
  {
    for (int q = lvl_ptr[l] + blockIdx.x; q < lvl_ptr[l + 1]; q += gridDim.x)
    {
      const BwdTask t = tasks[q];
      bwd_dispatch(t, Ridx, Mt, D, y, x, smem_top);
      __syncthreads();
    }
    grid.sync();
  }
}

__global__ void
k_coop_probe(int* out)
{
  cg::this_grid().sync();
  if (out && blockIdx.x == 0 && threadIdx.x == 0)
  {
    *out = 1;
  }
}

// ---- small utilities -----------------------------------------------------------------------------------
__global__ void
k_scatter(int nnz, const int* __restrict__ idx, int first, const double* __restrict__ val, double* __restrict__ out)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz)
  {
    out[idx ? idx[q] : first + q] = val[q];
  }
}

__global__ void
k_probe_rhs(int n, double* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    unsigned long long x = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 32;
    out[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}

__global__ void
k_sumsq(int n, const double* __restrict__ x, double* __restrict__ out)
{
  double v = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    v += x[i] * x[i];
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  if ((threadIdx.x & 31) == 0)
  {
    atomicAdd(out, v);
  }
}

__global__ void
k_absrange(int n, const double* __restrict__ x, unsigned long long* __restrict__ mn, unsigned long long* __restrict__ mx)
{
  double lo = __longlong_as_double(0x7ff0000000000000ll), hi = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const double a = fabs(x[i]);
    lo             = fmin(lo, a);
    hi             = fmax(hi, a);
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0)
  {
    atomicMin(mn, (unsigned long long)__double_as_longlong(lo));
    atomicMax(mx, (unsigned long long)__double_as_longlong(hi));
  }
}

__global__ void
k_init_range(double* scal)
{
  scal[2] = __longlong_as_double(0x7ff0000000000000ll);
  scal[3] = 0.0;
}

// ---- host-side enqueue helpers -----------------------------------------------------------------------------
static inline unsigned
nblocks(long long n, int threads)
{
  return (unsigned)((n + threads - 1) / threads);
}

constexpr size_t SOLVE_SMEM_LIMIT = 160 * 1024;

// ---- which levels go into the fused cooperative kernels ----------------------------------------------------
namespace
{
struct TopFusion
{
  int split;    // levels [split, nlevels) are fused; split == nlevels disables the fusion
  int grid_fwd; // cooperative grid sizes
  int grid_bwd;
  size_t smem;
};

int g_flow      = 1;  // dataflow sweeps (default) or one launch per level (B200_SWEEP=level)
int g_sms       = 148;
int g_flow_sleep = 512;
int g_flow_ctas  = FLOW_CTAS_PER_SM; // B200_FLOW_CTAS: fewer resident CTAs per SM (experiments)
int g_coop_ok   = -1; // -1 unknown, 0 unusable, 1 usable (process-wide)
int g_coop_ctas = 0;  // co-resident CTAs of the fused kernels with the worst-case shared memory

// Cooperative launches must be capturable into a CUDA graph and the device must support them; probed once,
// outside of any capture (configure_solve_kernels).
void
probe_cooperative()
{
  if (g_coop_ok >= 0)
  {
    return;
  }
  g_coop_ok = 0;
  // Measured on B200 (profiles/README.md): the fused kernels are slower than one launch per level (a graph
  // node costs only ~0.6 us), so they are opt-in.
  const char* e = std::getenv("B200_TOP_FUSION");
  if (!e || !*e || *e == '0')
  {
    return;
  }
  int dev = 0, coop = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop)
  {
    cudaGetLastError();
    return;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaStream_t s = nullptr;
  cudaGraph_t g  = nullptr;
  cudaGraphExec_t ge = nullptr;
  bool ok = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
  if (ok)
  {
    int* out     = nullptr;
    void* args[] = {&out};
    ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok)
    {
      const bool launched = cudaLaunchCooperativeKernel((void*)k_coop_probe, dim3(2), dim3(32), args, 0, s) == cudaSuccess;
      const bool ended    = cudaStreamEndCapture(s, &g) == cudaSuccess;
      ok                  = launched && ended && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess;
      if (ok)
      {
        ok = cudaGraphLaunch(ge, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
      }
    }
  }
  if (ge)
  {
    cudaGraphExecDestroy(ge);
  }
  if (g)
  {
    cudaGraphDestroy(g);
  }
  if (s)
  {
    cudaStreamDestroy(s);
  }
  cudaGetLastError();
  if (!ok)
  {
    return;
  }
  // worst-case shared memory of the fused kernels: 32 KB (fronts up to 4096 rows); larger levels stay unfused
  int per_sm_f = 0, per_sm_b = 0;
  const size_t smem = 32 * 1024;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f, k_fwd_top, SOLVE_THREADS, smem) != cudaSuccess ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, k_bwd_top, SOLVE_THREADS, smem) != cudaSuccess)
  {
    cudaGetLastError();
    return;
  }
  g_coop_ctas = sms * std::max(1, std::min(std::min(per_sm_f, per_sm_b), 4));
  g_coop_ok   = g_coop_ctas > 0;
}

TopFusion
plan_top_fusion(const Plan& P)
{
  TopFusion tf{P.nlevels, 0, 0, 0};
  if (g_coop_ok != 1 || P.nlevels < 3)
  {
    return tf;
  }
  // fuse the maximal run of top levels whose task counts fit the co-resident grid and whose fronts fit 32 KB
  int split = P.nlevels;
  int maxf = 0, maxb = 0, maxh = 0;
  for (int l = P.nlevels - 1; l >= 1; --l)
  {
    const int nf = P.fwd_ptr[l + 1] - P.fwd_ptr[l], nb = P.bwd_ptr[l + 1] - P.bwd_ptr[l];
    if (nf > 2 * g_coop_ctas || nb > 2 * g_coop_ctas || P.lvl_maxh[l] > 4096)
    {
      break;
    }
    split = l;
    maxf  = std::max(maxf, nf);
    maxb  = std::max(maxb, nb);
    maxh  = std::max(maxh, P.lvl_maxh[l]);
  }
  if (P.nlevels - split < 2)
  {
    return tf;
  }
  tf.split    = split;
  tf.grid_fwd = std::max(1, std::min(g_coop_ctas, maxf));
  tf.grid_bwd = std::max(1, std::min(g_coop_ctas, maxb));
  tf.smem     = sizeof(double) * (size_t)maxh;
  return tf;
}
} // namespace

static void
solve_levels(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc, cudaEvent_t* ev)
{
  // one launch per level of the supernodal tree (B200_SWEEP=level; kept for A/B measurements)
  auto mark = [&](int i) {
    if (ev)
    {
      B200_CUDA(cudaEventRecord(ev[i], stream));
    }
  };
  const Plan& P = *dp.plan;
  const TopFusion tf = plan_top_fusion(P);
  for (int l = 0; l < tf.split; ++l)
  {
    const int cnt     = P.fwd_ptr[l + 1] - P.fwd_ptr[l];
    const size_t smem = sizeof(double) * (size_t)P.lvl_maxh[l];
    k_fwd_chunk<<<cnt, SOLVE_THREADS, smem, stream>>>(dp.fwd_tasks.p + P.fwd_ptr[l], dp.Ridx.p, nb.Mr, nb.Dinv, sb.y, sb.yf);
    lc.tick();
  }
  if (tf.split < P.nlevels)
  {
    const FwdTask* tasks = dp.fwd_tasks.p;
    const int* lp        = dp.fwd_ptr.p;
    int l0 = tf.split, l1 = P.nlevels;
    const int* ridx  = dp.Ridx.p;
    const double *mr = nb.Mr, *di = nb.Dinv;
    double *ya = sb.y, *yf = sb.yf;
    void* args[] = {&tasks, &lp, &l0, &l1, &ridx, &mr, &di, &ya, &yf};
    B200_CUDA(cudaLaunchCooperativeKernel((void*)k_fwd_top, dim3(tf.grid_fwd), dim3(SOLVE_THREADS), args, tf.smem, stream));
    lc.tick();
  }
  mark(2);
  if (tf.split < P.nlevels)
  {
    const BwdTask* tasks = dp.bwd_tasks.p;
    const int* lp        = dp.bwd_ptr.p;
    int l0 = tf.split, l1 = P.nlevels;
    const int* ridx  = dp.Ridx.p;
    const double *mt = nb.Mt, *dd = nb.D, *yf = sb.yf;
    double* xx       = sb.x;
    void* args[] = {&tasks, &lp, &l0, &l1, &ridx, &mt, &dd, &yf, &xx};
    B200_CUDA(cudaLaunchCooperativeKernel((void*)k_bwd_top, dim3(tf.grid_bwd), dim3(SOLVE_THREADS), args, tf.smem, stream));
    lc.tick();
  }
  for (int l = tf.split - 1; l >= 0; --l)
  {
    const int cnt     = P.bwd_ptr[l + 1] - P.bwd_ptr[l];
    const size_t smem = sizeof(double) * (size_t)P.lvl_maxh[l];
    k_bwd_chunk<<<cnt, SOLVE_THREADS, smem, stream>>>(dp.bwd_tasks.p + P.bwd_ptr[l], dp.Ridx.p, nb.Mt, nb.D, sb.yf, sb.x);
    lc.tick();
  }
}

static void
solve_once(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, const double* in, double* out, cudaStream_t stream, LaunchCounter& lc, cudaEvent_t* ev = nullptr)
{
  auto mark = [&](int i) {
    if (ev)
    {
      B200_CUDA(cudaEventRecord(ev[i], stream));
    }
  };
  mark(0);
  const Plan& P = *dp.plan;
  const int T   = 256;
  if (P.m > 0)
  {
    k_pre<<<nblocks(P.m, T), T, 0, stream>>>(P.m, dp.k_of_r.p, dp.k_of_e.p, dp.pinv.p, dp.Acsr_ptr.p, dp.Acsr_col.p, nb.Acsr_val, nb.dE, in, sb.y);
    lc.tick();
    mark(1);
    if (g_flow)
    {
      const int ns    = P.nsuper;
      const int nflow = 2 * ns + 2 * FLOW_SHARDS * FLOW_TICKET_PITCH;
      k_flow_reset<<<nblocks(std::max(P.m, nflow), T), T, 0, stream>>>(P.m, nflow, sb.yf, sb.x, sb.flow);
      lc.tick();
      const int warps_per_cta = FLOW_THREADS / 32;
      const int max_ctas      = g_sms * g_flow_ctas;
      auto grid = [&](size_t ntasks) {
        const long long ctas = ((long long)ntasks + FLOW_SHARDS - 1) / FLOW_SHARDS;
        return (unsigned)std::max<long long>(1, std::min<long long>(max_ctas, ctas));
      };
      const FlowSched fs{(int)P.ffl_tasks.size(), g_flow_sleep};
      const FlowSched bs{(int)P.bfl_tasks.size(), g_flow_sleep};
      int* const tickets_f = sb.flow + 2 * ns;
      int* const tickets_b = tickets_f + FLOW_SHARDS * FLOW_TICKET_PITCH;
      auto kf = sb.trace_fwd ? k_flow<true, true> : k_flow<true, false>;
      auto kb = sb.trace_bwd ? k_flow<false, true> : k_flow<false, false>;
      kf<<<sb.trace_fwd ? max_ctas : grid(P.ffl_tasks.size()), FLOW_THREADS, 0, stream>>>(dp.ffl_tasks.p, fs, dp.Ridx.p, nb.Mt, nb.Dinv, sb.y, sb.yf, sb.x, sb.flow, tickets_f,
                                                                  (const FlowTrace*)sb.trace_fwd);
      lc.tick();
      mark(2);
      kb<<<sb.trace_bwd ? max_ctas : grid(P.bfl_tasks.size()), FLOW_THREADS, 0, stream>>>(dp.bfl_tasks.p, bs, dp.Ridx.p, nb.Mr, nb.Dinv, sb.y, sb.yf, sb.x, sb.flow + ns, tickets_b,
                                                                   (const FlowTrace*)sb.trace_bwd);
      lc.tick();
    }
    else
    {
      solve_levels(dp, nb, sb, stream, lc, ev);
    }
    mark(3);
    k_post_r<<<nblocks(P.m, T), T, 0, stream>>>(P.m, dp.k_of_r.p, dp.pinv.p, sb.x, out);
    lc.tick();
  }
  if (P.nE > 0)
  {
    k_post_e<<<nblocks(P.nE, T), T, 0, stream>>>(P.nE, dp.k_of_e.p, dp.pinv.p, dp.Acsc_ptr.p, dp.Acsc_row.p, nb.Acsc_val, nb.dE, in, sb.x, out);
    lc.tick();
  }
  mark(4);
}

static void
residual(const DevPlan& dp, const NumericBuffers& nb, const double* rhs, const double* z, double* res, cudaStream_t stream, LaunchCounter& lc)
{
  const Plan& P = *dp.plan;
  const int T   = 256;
  if (P.nE > 0)
  {
    k_resid_e<<<nblocks(P.nE, T), T, 0, stream>>>(P.nE, dp.k_of_e.p, dp.k_of_r.p, dp.Acsc_ptr.p, dp.Acsc_row.p, nb.Acsc_val, nb.dE, rhs, z, res);
    lc.tick();
  }
  if (P.m > 0)
  {
    k_resid_r<<<nblocks(P.m, T), T, 0, stream>>>(P.m, dp.k_of_e.p, dp.k_of_r.p, dp.Acsr_ptr.p, dp.Acsr_col.p, nb.Acsr_val, dp.Gsym_ptr.p, dp.Gsym_col.p, nb.Gsym_val, rhs, z, res);
    lc.tick();
  }
}

void
configure_solve_kernels()
{
  // once per process, thread-safe (handles are created concurrently by independent solver threads)
  static std::once_flag once;
  static std::string failure;
  std::call_once(once, [] {
    try
    {
      B200_CUDA(cudaFuncSetAttribute(k_fwd_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM_LIMIT));
      B200_CUDA(cudaFuncSetAttribute(k_bwd_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM_LIMIT));
      probe_cooperative();
      const char* sw = std::getenv("B200_SWEEP");
      g_flow         = !(sw && std::string(sw) == "level");
      if (const char* fc = std::getenv("B200_FLOW_CTAS"))
      {
        g_flow_ctas = std::min(FLOW_CTAS_PER_SM, std::max(1, std::atoi(fc)));
      }
      if (const char* fs = std::getenv("B200_FLOW_SLEEP"))
      {
        g_flow_sleep = std::max(32, std::atoi(fs));
      }
      int dev        = 0;
      B200_CUDA(cudaGetDevice(&dev));
      B200_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    catch (const CudaError& e)
    {
      failure = e.what();
    }
  });
  if (!failure.empty())
  {
    throw CudaError(failure);
  }
}

void
enqueue_solve_phases(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc, cudaEvent_t* ev)
{
  solve_once(dp, nb, sb, sb.rhs, sb.z, stream, lc, ev);
  B200_CUDA(cudaGetLastError());
}

void
enqueue_solve(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, int refine, cudaStream_t stream, LaunchCounter& lc)
{
  const Plan& P = *dp.plan;
  solve_once(dp, nb, sb, sb.rhs, sb.z, stream, lc);
  for (int it = 0; it < refine; ++it)
  {
    residual(dp, nb, sb.rhs, sb.z, sb.res, stream, lc);
    solve_once(dp, nb, sb, sb.res, sb.dz, stream, lc);
    k_axpy1<<<nblocks(P.N, 256), 256, 0, stream>>>(P.N, sb.dz, sb.z);
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

void
enqueue_residual_norms(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc)
{
  const Plan& P = *dp.plan;
  residual(dp, nb, sb.rhs, sb.z, sb.res, stream, lc);
  B200_CUDA(cudaMemsetAsync(nb.scal + 2, 0, 2 * sizeof(double), stream));
  const unsigned blocks = std::min<unsigned>(nblocks(P.N, 256), 1184);
  k_sumsq<<<blocks, 256, 0, stream>>>(P.N, sb.res, nb.scal + 2);
  lc.tick();
  k_sumsq<<<blocks, 256, 0, stream>>>(P.N, sb.rhs, nb.scal + 3);
  lc.tick();
  B200_CUDA(cudaGetLastError());
}

void
enqueue_scatter_rhs(double* rhs, int n, int nnz, const int* d_idx, int first, const double* d_val, cudaStream_t stream, LaunchCounter& lc)
{
  B200_CUDA(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)n, stream));
  if (nnz > 0)
  {
    k_scatter<<<nblocks(nnz, 256), 256, 0, stream>>>(nnz, d_idx, first, d_val, rhs);
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

void
enqueue_probe_rhs(double* rhs, int n, cudaStream_t stream, LaunchCounter& lc)
{
  if (n > 0)
  {
    k_probe_rhs<<<nblocks(n, 256), 256, 0, stream>>>(n, rhs);
    lc.tick();
  }
}

void
enqueue_pivot_range(const DevPlan& dp, const NumericBuffers& nb, cudaStream_t stream, LaunchCounter& lc)
{
  const Plan& P = *dp.plan;
  k_init_range<<<1, 1, 0, stream>>>(nb.scal);
  lc.tick();
  unsigned long long* mn = (unsigned long long*)(nb.scal + 2);
  unsigned long long* mx = (unsigned long long*)(nb.scal + 3);
  if (P.nE > 0)
  {
    k_absrange<<<std::min<unsigned>(nblocks(P.nE, 256), 1184), 256, 0, stream>>>(P.nE, nb.dE, mn, mx);
    lc.tick();
  }
  if (P.m > 0)
  {
    k_absrange<<<std::min<unsigned>(nblocks(P.m, 256), 1184), 256, 0, stream>>>(P.m, nb.D, mn, mx);
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

void
enqueue_copy_pivots(const DevPlan& dp, const NumericBuffers& nb, double* out, cudaStream_t stream, LaunchCounter&)
{
  const Plan& P = *dp.plan;
  if (P.nE > 0)
  {
    B200_CUDA(cudaMemcpyAsync(out, nb.dE, sizeof(double) * (size_t)P.nE, cudaMemcpyDeviceToDevice, stream));
  }
  if (P.m > 0)
  {
    B200_CUDA(cudaMemcpyAsync(out + P.nE, nb.D, sizeof(double) * (size_t)P.m, cudaMemcpyDeviceToDevice, stream));
  }
}

} // namespace b200
