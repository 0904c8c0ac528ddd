// sst.cu -- sparse subtrees (plan.hpp): complete subtrees of the elimination tree whose columns hold a few entries each
// (chains, banded systems) are factored and swept with their exact sparse structure, one CTA per subtree; everything
// of a subtree lives in shared memory while its CTA works.
//
// Why: as dense 32-column supernodes the chain of config 3 (n = 1e6) stored 8.8 M entries for an exact nnz(L) of
// 1.5 M, streamed 129 MB per sweep for 20 MB of factor, and paid one dataflow hand-off per 32-column level (7 levels
// after amalgamation, 15 before). As sparse subtrees it stores 1.005 x nnz(L).
//
// Subtrees nest: generation 0 are the leaves of the supernodal tree, a subtree of generation g sits on top of
// subtrees of earlier generations only (symbolic.cpp) -- for config 3 the whole top of the tree is ONE subtree of
// generation 1, and no dense supernode (hence no dataflow kernel) is left. One launch per generation.
//
// Inside a subtree the unit of work is a SEGMENT: a maximal single-child path of the subtree's own elimination tree
// (consecutive columns). One thread walks a segment column by column, segments of one level run in parallel; the
// number of block-wide synchronisations is the number of segment levels, not the height of the elimination tree.
//
//   k_sst_factor    assembly of the children's update blocks, sparse LDL^T (right-looking), r x r update block
//   k_sst_forward   y = D^-1 L^-1 b inside the subtree, contribution to the right-hand side of the ancestors
//   k_sst_backward  x = L^-T (y - L21^T x_ancestors)
#include "numeric.cuh"

namespace b200
{

namespace
{

typedef unsigned short u16;

constexpr size_t SST_SMEM = SST_SMEM_LIMIT; // upper bound of Plan::sst_smem_bytes (enforced by the analysis)

struct SstShared
{
  double* vals;        // nnz
  double* vec;         // k + r (sweeps) / r * r (factorization)
  const u16* slvl;     // nslev + 1: segments by level
  const u16* segstart; // first column of the segments, ordered by level
  const u16* seglen;
  const u16* colptr;   // k + 1
  const u16* rows;     // nnz, front-local, diagonal first
  const u16* rowptr;   // k + r + 1: the off-diagonal entries by row ...
  const u16* rcol;     // ... their columns (ascending inside a row)
  const u16* rpos;     // ... their positions in vals
};

// Bulk copy global -> shared in 16-byte pieces with eight loads per thread in flight before the first store: the
// staging of a subtree is a few dozen KB per CTA, and with one load per thread and loop trip it is nothing but memory
// latency (measured: 39 us of a 52 us forward sweep). Both pointers 16-byte aligned, n16 = number of 16-byte pieces
// (the plan pads every part so that reading up to the next multiple of 16 bytes is safe).
__device__ __forceinline__ void
stage16(void* dst, const void* __restrict__ src, int n16)
{
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d       = reinterpret_cast<int4*>(dst);
  for (int base = 0; base < n16; base += 8 * SST_THREADS)
  {
    int4 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const int i = base + u * SST_THREADS + threadIdx.x;
      if (i < n16)
      {
        t[u] = __ldcs(s + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const int i = base + u * SST_THREADS + threadIdx.x;
      if (i < n16)
      {
        d[i] = t[u];
      }
    }
  }
}

// carves the dynamic shared memory and loads the values and the parts of the index blob a kernel needs; the caller
// fills vec and synchronises. WHAT: 0 = everything, values by column (factorization); 1 = segments + row view, values by
// row (forward sweep: Mt holds the off-diagonal values in row order); 2 = segments + column view, values by column
// (backward sweep). The blob keeps its layout in shared memory, parts that are not needed are simply not copied.
template <int WHAT>
__device__ __forceinline__ SstShared
sst_stage(double* smem, const SstMeta& M, const u16* __restrict__ blob_all, const double* __restrict__ values, int vec_len)
{
  SstShared S;
  const int nv = (M.nnz + 1) & ~1, nx = (vec_len + 1) & ~1;
  S.vals      = smem;
  S.vec       = S.vals + nv;
  u16* blob   = reinterpret_cast<u16*>(S.vec + nx);
  S.slvl      = blob;
  S.segstart  = blob + M.o_segstart;
  S.seglen    = blob + M.o_seglen;
  S.colptr    = blob + M.o_colptr;
  S.rows      = blob + M.o_rows;
  S.rowptr    = blob + M.o_rowptr;
  S.rcol      = blob + M.o_rcol;
  S.rpos      = blob + M.o_rpos;
  const u16* src = blob_all + M.blob;
  if (WHAT == 0)
  {
    stage16(S.vals, values, nv / 2);
    stage16(blob, src, M.blob_len16);
  }
  else if (WHAT == 1)
  {
    stage16(S.vals, values, ((M.nnz - M.k + 1) & ~1) / 2);
    stage16(blob, src, M.o_colptr / 8);
    stage16(blob + M.o_rowptr, src + M.o_rowptr, (M.o_rpos - M.o_rowptr) / 8);
  }
  else
  {
    stage16(S.vals, values, nv / 2);
    stage16(blob, src, M.o_rowptr / 8);
  }
  return S;
}

__device__ __forceinline__ int
ld_acquire(const int* p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// thread 0 polls a counter until it reaches `need`; the barrier hands the acquired view to the whole CTA (the data
// behind the counter is then read with L2 loads, __ldcg)
__device__ __forceinline__ void
wait_counter(const int* c, int need)
{
  if (threadIdx.x == 0)
  {
    while (ld_acquire(c) < need)
    {
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// Order of work inside a launch that holds several generations: a CTA that has to wait only ever waits for CTAs with
// lower tickets, which are running already -- no assumption on the order the hardware dispatches blocks in, no bound
// on the grid. With one generation (ticket == nullptr) nobody waits and the block index will do.
__device__ __forceinline__ int
take_ticket(int* ticket)
{
  if (ticket == nullptr)
  {
    return blockIdx.x;
  }
  __shared__ int t;
  if (threadIdx.x == 0)
  {
    t = atomicAdd(ticket, 1);
  }
  __syncthreads();
  return t;
}

// Phase stamps of the forward sweep (build with -DB200_SST_TRACE_BUILD, run profiles/prof_driver.py with B200_SST_TRACE=1):
// %globaltimer at the phase boundaries of the CTAs with tickets 0 and 300 and of the first subtree that has children
#ifdef B200_SST_TRACE_BUILD
__device__ unsigned long long g_sst_trace[4 * 8];
__device__ __forceinline__ unsigned long long
gtime()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SST_TR(which, ph)                          \
  if (threadIdx.x == 0 && (which) >= 0)            \
  {                                                \
    g_sst_trace[(which) * 8 + (ph)] = gtime();     \
  }
#else
#define SST_TR(which, ph)
#endif

__device__ __forceinline__ void
smem_add(double* p, double v)
{
  atomicAdd(p, v); // shared-memory FP64 atomic
}

// Shared memory through 32-bit shared-window addresses and explicit ld.shared / st.shared: with pointers derived from the
// dynamic shared array the compiler rebuilt the window base (S2R SR_CgaCtaId, the window of a CTA inside a cluster)
// inside the loops. shared_addr is opaque to the optimiser on purpose (a volatile asm runs once, its result stays in a
// register).
__device__ __forceinline__ unsigned
shared_addr(const void* p)
{
  unsigned a;
  asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(a) : "l"(p));
  return a;
}

__device__ __forceinline__ double
lds64(unsigned a)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

__device__ __forceinline__ int
lds16(unsigned a)
{
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
  return (int)v;
}

__device__ __forceinline__ void
sts64(unsigned a, double v)
{
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

// One thread walks the columns of a segment: x_j -= sum over the entries of column j (backward: rows below the diagonal
// through the column view) or of row j (forward: columns left of the diagonal through the row view) of val * x[idx].
// Measured (clock64 around the walk of thread 0, config 3): 340 cycles per column of a 14-column leaf segment and
// ~1,000 cycles for a single column with 15-26 entries, where the isolated latencies (profiles/micro/chain_latency.cu:
// DFMA 8.7, ld.shared 28.8, st -> ld 39.5 cycles) predict 100-150; hand-pipelined variants (operands one or two columns
// ahead) changed little or lost to register pressure, so the shared-memory round trips evidently take far longer than
// in isolation while the other CTAs of the SM stage their subtrees. What is kept is what needs the fewest round trips
// and instructions: short rows / columns (one or two entries: chains) take a lean scalar path -- address registers
// stepped, the column just finished taken from a register, the next entry range loaded a column ahead; longer ones go
// four entries at a time (indices and values first, then the gathers, then the products).
// ptr / idx / vals / x: shared-window byte addresses of the arrays.
template <bool FWD>
__device__ __forceinline__ void
sst_walk(unsigned ptr, unsigned idx, unsigned vals, unsigned x, int j0, int j1)
{
  int j          = FWD ? j0 : j1 - 1;
  unsigned pj    = ptr + 2u * (unsigned)j; // &ptr[j]
  unsigned xj    = x + 8u * (unsigned)j;   // &x[j]
  int e          = lds16(pj) + (FWD ? 0 : 1);
  int e1         = lds16(pj + 2u);
  double prev    = 0.0;
  int prevc      = -2;
  for (int n = j1 - j0; n > 0; --n)
  {
    // entry range of the next column (forward: it begins where this one ends)
    const unsigned pn = FWD ? pj + 2u : pj - 2u;
    const int en      = n > 1 ? lds16(pn) + (FWD ? 0 : 1) : 0;
    const int en1     = n > 1 ? lds16(pn + 2u) : 0;
    double acc        = lds64(xj);
    const int cnt     = e1 - e;
    unsigned ai = idx + 2u * (unsigned)e, av = vals + 8u * (unsigned)e;
    if (cnt <= 2)
    {
      if (cnt >= 1)
      {
        const int c    = lds16(ai);
        const double v = lds64(av);
        acc -= v * (c == prevc ? prev : lds64(x + 8u * (unsigned)c));
      }
      if (cnt == 2)
      {
        const int c    = lds16(ai + 2u);
        const double v = lds64(av + 8u);
        acc -= v * (c == prevc ? prev : lds64(x + 8u * (unsigned)c));
      }
    }
    else
    {
      for (int left = cnt; left > 0; left -= 4, ai += 8u, av += 32u)
      {
        int c[4];
        double v[4], xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          c[u] = u < left ? lds16(ai + 2u * u) : -1;
          v[u] = u < left ? lds64(av + 8u * u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          xv[u] = c[u] >= 0 ? lds64(x + 8u * (unsigned)c[u]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          acc -= v[u] * xv[u];
        }
      }
    }
    sts64(xj, acc);
    prev  = acc;
    prevc = j;
    j += FWD ? 1 : -1;
    pj = pn;
    xj = FWD ? xj + 8u : xj - 8u;
    e  = en;
    e1 = en1;
  }
}

} // namespace

__global__ void __launch_bounds__(SST_THREADS)
k_sst_factor(const SstMeta* __restrict__ metas,
             const u16* __restrict__ blob_all,
             const long long* __restrict__ ea_src,
             const int* __restrict__ ea_dst,
             double* __restrict__ L,
             double* __restrict__ Mt,
             double* __restrict__ U,
             double* __restrict__ D,
             double* __restrict__ Dinv,
             const double* __restrict__ scal,
             int* __restrict__ n_perturbed)
{
  extern __shared__ double sst_smem[];
  const SstMeta M = metas[blockIdx.x];
  const int k = M.k, r = M.r;
  double* Lg        = L + M.Lptr;
  const SstShared S = sst_stage<0>(sst_smem, M, blob_all, Lg, r * r); // vals = assembled entries of S, zeros in the fill
  double* vals      = S.vals;
  double* Us        = S.vec;
  for (int q = threadIdx.x; q < r * r; q += blockDim.x)
  {
    Us[q] = 0.0;
  }
  const double tau = scal[1];
  int nper         = 0;
  __syncthreads();
  // extend-add of the children (subtrees of earlier generations, factored by earlier launches)
  for (int base = M.ea_begin; base < M.ea_end; base += 4 * SST_THREADS)
  {
    double v[4];
    int dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int e = base + u * SST_THREADS + threadIdx.x;
      if (e < M.ea_end)
      {
        dst[u] = ea_dst[e];
        v[u]   = U[ea_src[e]];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int e = base + u * SST_THREADS + threadIdx.x;
      if (e < M.ea_end)
      {
        smem_add(dst[u] >= 0 ? vals + dst[u] : Us + (-1 - dst[u]), v[u]);
      }
    }
  }
  if (M.ea_end > M.ea_begin)
  {
    __syncthreads();
  }
  for (int lev = 0; lev < M.nslev; ++lev)
  {
    for (int q = S.slvl[lev] + threadIdx.x; q < S.slvl[lev + 1]; q += blockDim.x)
    {
      const int j0 = S.segstart[q], j1 = j0 + S.seglen[q];
      for (int j = j0; j < j1; ++j)
      {
        const int p0 = S.colptr[j], p1 = S.colptr[j + 1];
        double d     = vals[p0];
        if (!(fabs(d) >= tau) || !isfinite(d))
        {
          d = tau > 0.0 ? -tau : -1e-300; // static pivoting, same rule as k_panel
          ++nper;
        }
        const double dinv = 1.0 / d;
        D[M.first + j]    = d;
        Dinv[M.first + j] = dinv;
        vals[p0]          = d;
        // right-looking update: A[ia, ib] -= f_a f_b / d for the entries a >= b of the column. All targets belong to
        // ancestors of j: later columns of this segment (only this thread touches them before the next barrier, but
        // other segments of the level may share targets further up: atomics throughout)
        for (int a = p0 + 1; a < p1; ++a)
        {
          const double la = vals[a] * dinv;
          const int ia    = S.rows[a];
          for (int b = p0 + 1; b <= a; ++b)
          {
            const int ib   = S.rows[b];
            const double u = -la * vals[b];
            if (ib >= k)
            {
              smem_add(Us + (ia - k) + (ib - k) * r, u);
            }
            else
            {
              int t = S.colptr[ib]; // position of row ia in column ib: the structure of an ancestor contains it
              while (S.rows[t] != ia)
              {
                ++t;
              }
              smem_add(vals + t, u);
            }
          }
        }
        for (int a = p0 + 1; a < p1; ++a)
        {
          vals[a] *= dinv; // l_ij
        }
      }
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < M.nnz; q += blockDim.x)
  {
    Lg[q] = vals[q];
  }
  // the off-diagonal entries once more in row order, where the inverse panels of a dense supernode would be: what the
  // forward sweep streams (a column gathers along its row; no indirection through value positions there)
  double* Rg = Mt + M.Lptr;
  for (int q = threadIdx.x; q < M.nnz - k; q += blockDim.x)
  {
    Rg[q] = vals[S.rpos[q]];
  }
  double* Ug = U + M.Uoff;
  for (int q = threadIdx.x; q < r * r; q += blockDim.x)
  {
    Ug[q] = Us[q]; // the whole block: nobody zeroes the update matrix of a sparse subtree
  }
  if (nper)
  {
    atomicAdd(n_perturbed, nper);
  }
}

// forward: inside the subtree y = L^-1 b segment level by segment level. A column GATHERS along its row,
// x_j = b_j - sum_c l_jc x_c: the entries of row j belong to earlier columns of the thread's own segment or to segments
// of lower levels, so nothing inside a subtree needs an atomic (shared-memory FP64 atomics are compare-and-swap loops,
// ATOMS.CAST.SPIN: the scatter form of this loop spent 7.8 us in the seven levels of a 976-column subtree, measured
// with %globaltimer) and the sums of a subtree come out in a fixed order. yf = D^-1 y; the tail rows gather after the
// last level and go to the global accumulator (atomics: several subtrees share an ancestor row), then the parent's
// dependency counter is signalled (a dense parent's dataflow tasks and a parent subtree wait on the same counters)
__global__ void __launch_bounds__(SST_THREADS)
k_sst_forward(const SstMeta* __restrict__ metas,
              const u16* __restrict__ blob_all,
              const int* __restrict__ Ridx,
              const double* __restrict__ L,
              const double* __restrict__ Dinv,
              double* __restrict__ yacc,
              double* __restrict__ yf,
              int* __restrict__ cnt,
              int* __restrict__ ticket)
{
  extern __shared__ double sst_smem[];
#ifdef B200_SST_TRACE_BUILD
  const unsigned long long T0 = gtime();
#endif
  const int tk    = take_ticket(ticket);
  const SstMeta M = metas[tk]; // children (earlier generations) get the lower tickets
  const int k = M.k, r = M.r;
#ifdef B200_SST_TRACE_BUILD
  const int W = tk == 0 ? 0 : (tk == 300 ? 1 : (M.nchild > 0 ? 2 : -1));
  if (threadIdx.x == 0 && W >= 0)
  {
    g_sst_trace[W * 8 + 0] = T0;
  }
#endif
  SST_TR(W, 1) // ticket + record
  const SstShared S = sst_stage<1>(sst_smem, M, blob_all, L + M.Lptr, k + r); // L: here the row-ordered copy (Mt)
  double* x         = S.vec;
  SST_TR(W, 2) // structure and values staged
  if (M.nchild > 0)
  {
    // child subtrees of earlier generations run in this very launch under lower tickets, i.e. they are already
    // running: the structure is staged, now wait for their shares of the right-hand side
    wait_counter(cnt + M.sn, M.nchild);
  }
  // the reciprocal pivots are only needed after the last level: requested here, they arrive while the levels run (the
  // sampled stalls of the kernel had 13 % of all warp samples on the product x * Dinv at the end, waiting for this load)
  double dinv[(SST_MAX_COLS + SST_THREADS - 1) / SST_THREADS];
#pragma unroll
  for (int u = 0; u < (int)(sizeof(dinv) / sizeof(double)); ++u)
  {
    const int q = u * SST_THREADS + threadIdx.x;
    dinv[u]     = q < k ? Dinv[M.first + q] : 0.0;
  }
  {
    double t[(SST_MAX_COLS + SST_MAX_TAIL + SST_THREADS - 1) / SST_THREADS]; // all loads of the right-hand side first
#pragma unroll
    for (int u = 0; u < (int)(sizeof(t) / sizeof(double)); ++u)
    {
      const int q = u * SST_THREADS + threadIdx.x;
      t[u]        = q < k ? __ldcg(yacc + M.first + q) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < (int)(sizeof(t) / sizeof(double)); ++u)
    {
      const int q = u * SST_THREADS + threadIdx.x;
      if (q < k + r)
      {
        x[q] = t[u];
      }
    }
  }
  __syncthreads();
  SST_TR(W, 3) // (children waited for,) right-hand side in shared memory
  {
    const unsigned sv = shared_addr(S.vals);
    const unsigned sx = sv + (unsigned)((const char*)x - (const char*)S.vals), sb = sv + (unsigned)((const char*)S.slvl - (const char*)S.vals);
    const unsigned s_start = sb + 2u * (unsigned)M.o_segstart, s_len = sb + 2u * (unsigned)M.o_seglen, s_ptr = sb + 2u * (unsigned)M.o_rowptr,
                   s_idx = sb + 2u * (unsigned)M.o_rcol;
    int q0 = lds16(sb);
    for (int lev = 0; lev < M.nslev; ++lev)
    {
      const int q1 = lds16(sb + 2u * (unsigned)(lev + 1));
      for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x)
      {
        const int j0 = lds16(s_start + 2u * (unsigned)q);
        sst_walk<true>(s_ptr, s_idx, sv, sx, j0, j0 + lds16(s_len + 2u * (unsigned)q));
      }
      q0 = q1;
      __syncthreads();
    }
  }
  SST_TR(W, 4) // levels
  for (int q = threadIdx.x; q < r; q += blockDim.x)
  {
    double acc = 0.0;
    for (int e = S.rowptr[k + q]; e < S.rowptr[k + q + 1]; ++e)
    {
      acc -= S.vals[e] * x[S.rcol[e]];
    }
    x[k + q] = acc; // only this thread reads it again (below)
  }
#pragma unroll
  for (int u = 0; u < (int)(sizeof(dinv) / sizeof(double)); ++u)
  {
    const int q = u * SST_THREADS + threadIdx.x;
    if (q < k)
    {
      yf[M.first + q] = x[q] * dinv[u];
    }
  }
  for (int q = threadIdx.x; q < r; q += blockDim.x)
  {
    atomicAdd(yacc + Ridx[M.Rptr + q], x[k + q]);
  }
  if (M.signal >= 0)
  {
    __syncthreads();
    if (threadIdx.x == 0)
    {
      // cumulativity: the barrier ordered the CTA's stores and atomics before this thread's release fence
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      atomicAdd(cnt + M.signal, 1);
    }
  }
  SST_TR(W, 5) // tails, stores, signal
}

// backward: x_j = yf_j - sum_i l_ij x_i over the entries of column j; segment levels from the root of the subtree
// down, the columns of a segment from its top to its bottom (no atomics: a column only reads its ancestors)
__global__ void __launch_bounds__(SST_THREADS)
k_sst_backward(const SstMeta* __restrict__ metas,
               const u16* __restrict__ blob_all,
               const int* __restrict__ Ridx,
               const double* __restrict__ L,
               const double* __restrict__ yf,
               double* __restrict__ xg,
               int* __restrict__ done,
               int* __restrict__ ticket)
{
  extern __shared__ double sst_smem[];
  const SstMeta M = metas[gridDim.x - 1 - take_ticket(ticket)]; // parents (later generations) get the lower tickets
  const int k = M.k, r = M.r;
  const SstShared S = sst_stage<2>(sst_smem, M, blob_all, L + M.Lptr, k + r);
  double* x         = S.vec;
  {
    double t[(SST_MAX_COLS + SST_THREADS - 1) / SST_THREADS];
#pragma unroll
    for (int u = 0; u < (int)(sizeof(t) / sizeof(double)); ++u)
    {
      const int q = u * SST_THREADS + threadIdx.x;
      t[u]        = q < k ? yf[M.first + q] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < (int)(sizeof(t) / sizeof(double)); ++u)
    {
      const int q = u * SST_THREADS + threadIdx.x;
      if (q < k)
      {
        x[q] = t[u];
      }
    }
  }
  if (M.parent_sst >= 0)
  {
    wait_counter(done + M.parent_sst, 1); // the parent subtree runs in this very launch (lower ticket: already running)
  }
  for (int q = threadIdx.x; q < r; q += blockDim.x)
  {
    x[k + q] = __ldcg(xg + Ridx[M.Rptr + q]);
  }
  __syncthreads();
  {
    const unsigned sv = shared_addr(S.vals);
    const unsigned sx = sv + (unsigned)((const char*)x - (const char*)S.vals), sb = sv + (unsigned)((const char*)S.slvl - (const char*)S.vals);
    const unsigned s_start = sb + 2u * (unsigned)M.o_segstart, s_len = sb + 2u * (unsigned)M.o_seglen, s_ptr = sb + 2u * (unsigned)M.o_colptr,
                   s_idx = sb + 2u * (unsigned)M.o_rows;
    int q1 = lds16(sb + 2u * (unsigned)M.nslev);
    for (int lev = M.nslev - 1; lev >= 0; --lev)
    {
      const int q0 = lds16(sb + 2u * (unsigned)lev);
      for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x)
      {
        const int j0 = lds16(s_start + 2u * (unsigned)q);
        sst_walk<false>(s_ptr, s_idx, sv, sx, j0, j0 + lds16(s_len + 2u * (unsigned)q));
      }
      q1 = q0;
      __syncthreads();
    }
  }
  for (int q = threadIdx.x; q < k; q += blockDim.x)
  {
    xg[M.first + q] = x[q];
  }
  if (M.nchild > 0)
  {
    __syncthreads();
    if (threadIdx.x == 0)
    {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      atomicAdd(done + M.sn, 1);
    }
  }
}

void
dump_sst_trace()
{
#ifdef B200_SST_TRACE_BUILD
  unsigned long long h[32];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_sst_trace, sizeof(h));
  for (int w = 0; w < 3; ++w)
  {
    std::fprintf(stderr, "[sst trace %d] start %+lld ns:", w, (long long)(h[w * 8] - h[0]));
    for (int ph = 1; ph < 6; ++ph)
    {
      std::fprintf(stderr, " +%lld", (long long)(h[w * 8 + ph] - h[w * 8]));
    }
    std::fprintf(stderr, "  (ticket+record, staged, rhs, levels, end)\n");
  }
#endif
}

void
configure_sst_kernels(int device)
{
  static std::mutex mu;
  static std::vector<int> done;
  std::lock_guard<std::mutex> lock(mu);
  if (std::find(done.begin(), done.end(), device) != done.end())
  {
    return;
  }
  B200_CUDA(cudaFuncSetAttribute(k_sst_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SST_SMEM));
  B200_CUDA(cudaFuncSetAttribute(k_sst_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SST_SMEM));
  B200_CUDA(cudaFuncSetAttribute(k_sst_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SST_SMEM));
  done.push_back(device);
}

void
enqueue_sst_factor(const DevPlan& dp, const NumericBuffers& nb, cudaStream_t stream, LaunchCounter& lc)
{
  const std::vector<int>& gp = dp.plan->sst_gen_ptr;
  for (size_t g = 0; g + 1 < gp.size(); ++g) // children first
  {
    if (gp[g + 1] == gp[g])
    {
      continue;
    }
    k_sst_factor<<<gp[g + 1] - gp[g], SST_THREADS, dp.plan->sst_smem_bytes, stream>>>(dp.sst.p + gp[g], dp.sst_blob.p, dp.sst_ea_src.p, dp.sst_ea_dst.p, nb.L, nb.Mt, nb.U, nb.D, nb.Dinv, nb.scal,
                                                                                      nb.n_perturbed);
    lc.tick("sst");
    B200_CUDA(cudaGetLastError());
  }
}

// One launch per sweep for all generations: the waits inside the kernels order parents and children
void
enqueue_sst_forward(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc)
{
  const int n = (int)dp.plan->sst.size();
  if (n == 0)
  {
    return;
  }
  int* const ticket = dp.plan->sst_gen_ptr.size() > 2 ? sb.flow + sst_ticket_offset(dp.plan->nsuper) : nullptr;
  k_sst_forward<<<n, SST_THREADS, dp.plan->sst_smem_bytes, stream>>>(dp.sst.p, dp.sst_blob.p, dp.Ridx.p, nb.Mt, nb.Dinv, sb.y, sb.yf, sb.flow, ticket);
  lc.tick();
  B200_CUDA(cudaGetLastError());
}

void
enqueue_sst_backward(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc)
{
  const int n = (int)dp.plan->sst.size();
  if (n == 0)
  {
    return;
  }
  int* const ticket = dp.plan->sst_gen_ptr.size() > 2 ? sb.flow + sst_ticket_offset(dp.plan->nsuper) + 4 : nullptr;
  int* const done   = sb.flow + dp.plan->nsuper; // the backward counters of the sweeps; those of sparse subtrees are free
  k_sst_backward<<<n, SST_THREADS, dp.plan->sst_smem_bytes, stream>>>(dp.sst.p, dp.sst_blob.p, dp.Ridx.p, nb.L, sb.yf, sb.x, done, ticket);
  lc.tick();
  B200_CUDA(cudaGetLastError());
}

} // namespace b200
