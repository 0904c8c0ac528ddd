// solve.cu -- the repeated solves K z = b behind SleqpFactCallbacks.solve (fact_types.h:12;
// vendor calls replaced: umfpack_di_solve fact_umfpack.c:220, cholmod_l_solve fact_cholmod.c:184).
//
//   K = [D A^T; A G]:  t = b_E / d,  b_R' = b_R - A t,  S y = b_R',  z_R = y,  z_E = (b_E - A^T y) / d
// The reduced system is solved with level-scheduled supernodal forward / diagonal / backward
// sweeps over the multifrontal front vectors (children are pulled by their parent, one CTA per
// supernode, so the summation order is fixed and the solve is deterministic).
// Iterative refinement runs against the unperturbed K (SURVEY.md hard part 1).
#include "numeric.cuh"

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

// ---- forward sweep: one CTA per supernode of the level -----------------------------------------------
__global__ void __launch_bounds__(SOLVE_THREADS)
k_fwd_level(const int* __restrict__ lvl_sn,
            const SnMeta* __restrict__ sn,
            const int* __restrict__ child_idx,
            const int* __restrict__ rel,
            const double* __restrict__ L,
            const double* __restrict__ b,
            double* __restrict__ y,
            double* __restrict__ W,
            int use_smem)
{
  extern __shared__ double smem[];
  const SnMeta s  = sn[lvl_sn[blockIdx.x]];
  const int k = s.k, h = s.k + s.r;
  const double* P = L + s.Lptr;
  double* Wg      = W + s.Wptr;
  double* w       = use_smem ? smem : Wg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < h; i += SOLVE_THREADS)
  {
    w[i] = i < k ? b[s.first + i] : 0.0;
  }
  __syncthreads();
  for (int q = s.child_begin; q < s.child_end; ++q)
  {
    const SnMeta c   = sn[child_idx[q]];
    const int* rl    = rel + c.Rptr;
    const double* wc = W + c.Wptr + c.k;
    for (int i = tid; i < c.r; i += SOLVE_THREADS)
    {
      w[rl[i]] += wc[i];
    }
    __syncthreads();
  }
  for (int jb = 0; jb < k; jb += 32)
  {
    const int wb = min(32, k - jb);
    if (warp == 0)
    {
      double a[32];
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
      {
        a[jj] = (jj < lane && lane < wb) ? P[(long long)(jb + jj) * h + jb + lane] : 0.0;
      }
      double v = lane < wb ? w[jb + lane] : 0.0;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj)
      {
        const double yj = __shfl_sync(0xffffffffu, v, jj);
        if (lane > jj)
        {
          v -= a[jj] * yj;
        }
      }
      if (lane < wb)
      {
        w[jb + lane] = v;
      }
    }
    __syncthreads();
    for (int i = jb + wb + tid; i < h; i += SOLVE_THREADS)
    {
      double acc = w[i];
      for (int jj = 0; jj < wb; ++jj)
      {
        acc -= P[(long long)(jb + jj) * h + i] * w[jb + jj];
      }
      w[i] = acc;
    }
    __syncthreads();
  }
  for (int i = tid; i < h; i += SOLVE_THREADS)
  {
    if (i < k)
    {
      y[s.first + i] = w[i];
    }
    else if (use_smem)
    {
      Wg[i] = w[i];
    }
  }
}

// ---- diagonal + backward sweep -----------------------------------------------------------------------
__global__ void __launch_bounds__(SOLVE_THREADS)
k_bwd_level(const int* __restrict__ lvl_sn,
            const SnMeta* __restrict__ sn,
            const int* __restrict__ Ridx,
            const double* __restrict__ L,
            const double* __restrict__ D,
            double* __restrict__ y,
            double* __restrict__ W,
            int use_smem)
{
  extern __shared__ double smem[];
  const SnMeta s  = sn[lvl_sn[blockIdx.x]];
  const int k = s.k, h = s.k + s.r;
  const double* P = L + s.Lptr;
  double* w       = use_smem ? smem : W + s.Wptr;
  const int* rows = Ridx + s.Rptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW  = SOLVE_THREADS / 32;
  constexpr int CPW = 32 / NW; // columns per warp inside a 32-column block

  for (int i = tid; i < h; i += SOLVE_THREADS)
  {
    w[i] = i < k ? y[s.first + i] / D[s.first + i] : y[rows[i - k]];
  }
  __syncthreads();
  const int nblk = (k + 31) / 32;
  for (int blk = nblk - 1; blk >= 0; --blk)
  {
    const int jb = blk * 32;
    const int wb = min(32, k - jb);
    // w[c] -= sum_{i >= jb+wb} L[i][c] * w[i] for the block's columns
    {
      double acc[CPW];
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc)
      {
        acc[cc] = 0.0;
      }
      for (int i = jb + wb + lane; i < h; i += 32)
      {
        const double wi = w[i];
#pragma unroll
        for (int cc = 0; cc < CPW; ++cc)
        {
          const int c = jb + warp * CPW + cc;
          if (c < jb + wb)
          {
            acc[cc] += P[(long long)c * h + i] * wi;
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < CPW; ++cc)
      {
        double v = acc[cc];
        for (int o = 16; o > 0; o >>= 1)
        {
          v += __shfl_xor_sync(0xffffffffu, v, o);
        }
        const int c = jb + warp * CPW + cc;
        if (lane == 0 && c < jb + wb)
        {
          w[c] -= v;
        }
      }
    }
    __syncthreads();
    if (warp == 0)
    {
      // L11^T x = t on the block: lane j owns column j of the block
      double col[32];
#pragma unroll
      for (int i = 0; i < 32; ++i)
      {
        col[i] = (i > lane && i < wb) ? P[(long long)(jb + lane) * h + jb + i] : 0.0;
      }
      double v = lane < wb ? w[jb + lane] : 0.0;
#pragma unroll
      for (int jj = 31; jj >= 0; --jj)
      {
        const double xj = __shfl_sync(0xffffffffu, v, jj);
        if (lane < jj)
        {
          v -= col[jj] * xj;
        }
      }
      if (lane < wb)
      {
        w[jb + lane] = v;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < k; i += SOLVE_THREADS)
  {
    y[s.first + i] = w[i];
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
    k_pre<<<nblocks(P.m, T), T, 0, stream>>>(P.m, dp.k_of_r.p, dp.k_of_e.p, dp.pinv.p, dp.Acsr_ptr.p, dp.Acsr_col.p, nb.Acsr_val, nb.dE, in, sb.bR);
    lc.tick();
    mark(1);
    for (int l = 0; l < P.nlevels; ++l)
    {
      const int cnt          = P.lvl_ptr[l + 1] - P.lvl_ptr[l];
      const size_t smem_need = sizeof(double) * (size_t)dp.lvl_maxh[l];
      const int use_smem     = smem_need <= SOLVE_SMEM_LIMIT;
      const size_t smem      = use_smem ? smem_need : 0;
      k_fwd_level<<<cnt, SOLVE_THREADS, smem, stream>>>(dp.lvl_sn.p + P.lvl_ptr[l], dp.sn.p, dp.child_idx.p, dp.rel.p, nb.L, sb.bR, sb.y, sb.W, use_smem);
      lc.tick();
    }
    mark(2);
    for (int l = P.nlevels - 1; l >= 0; --l)
    {
      const int cnt          = P.lvl_ptr[l + 1] - P.lvl_ptr[l];
      const size_t smem_need = sizeof(double) * (size_t)dp.lvl_maxh[l];
      const int use_smem     = smem_need <= SOLVE_SMEM_LIMIT;
      const size_t smem      = use_smem ? smem_need : 0;
      k_bwd_level<<<cnt, SOLVE_THREADS, smem, stream>>>(dp.lvl_sn.p + P.lvl_ptr[l], dp.sn.p, dp.Ridx.p, nb.L, nb.D, sb.y, sb.W, use_smem);
      lc.tick();
    }
    mark(3);
    k_post_r<<<nblocks(P.m, T), T, 0, stream>>>(P.m, dp.k_of_r.p, dp.pinv.p, sb.y, out);
    lc.tick();
  }
  if (P.nE > 0)
  {
    k_post_e<<<nblocks(P.nE, T), T, 0, stream>>>(P.nE, dp.k_of_e.p, dp.pinv.p, dp.Acsc_ptr.p, dp.Acsc_row.p, nb.Acsc_val, nb.dE, in, sb.y, out);
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
  static bool done = false;
  if (!done)
  {
    B200_CUDA(cudaFuncSetAttribute(k_fwd_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM_LIMIT));
    B200_CUDA(cudaFuncSetAttribute(k_bwd_level, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SOLVE_SMEM_LIMIT));
    done = true;
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
