// sst.cu -- sparse subtrees (plan.hpp): complete subtrees of the elimination tree whose columns hold a few entries each
// (chains, banded systems) are factored and swept with their exact sparse structure, one CTA per subtree, level by
// level of the subtree's own elimination tree; everything of a subtree lives in shared memory while its CTA works.
//
// Why: as dense 32-column supernodes the chain of config 3 (n = 1e6) stored 8.8 M entries for an exact nnz(L) of
// 1.5 M, streamed 129 MB per sweep for 20 MB of factor, and paid one dataflow hand-off per 32-column level (7 levels
// after amalgamation, 15 before). As 512 sparse subtrees of ~1000 columns it stores 1.005 x nnz(L) and the tree above
// them is three small dense levels.
//
//   k_sst_factor    sparse LDL^T of every subtree (right-looking, columns of one level in parallel, shared-memory
//                   atomics for the updates) + its r x r contribution block for the parent front
//   k_sst_forward   y = D^-1 L^-1 b inside the subtree, contribution of the subtree to the right-hand side of its ancestors
//   k_sst_backward  x = L^-T (y - L21^T x_ancestors)
#include "numeric.cuh"

namespace b200
{

namespace
{

typedef unsigned short u16;

// Shared memory of one CTA: the values of the subtree, the front vector (sweeps) or the update block (factorization),
// and the WHOLE index structure of the subtree (column pointers, front-local rows, columns by level as 16-bit
// integers, level pointers): a level of a subtree is a few dozen columns with two or three entries each, so anything
// fetched from global memory inside the level loop costs a full memory latency per level (measured: 60 us for the
// forward sweep of config 3 with the indices in global memory, against 13 us for the three dense levels above).
constexpr int SST_VEC      = SST_MAX_COLS + SST_MAX_TAIL > SST_MAX_TAIL * SST_MAX_TAIL ? SST_MAX_COLS + SST_MAX_TAIL : SST_MAX_TAIL * SST_MAX_TAIL;
constexpr size_t SST_SMEM = sizeof(double) * (size_t)(SST_MAX_NNZ + SST_VEC + 8) + sizeof(int) * (size_t)(SST_MAX_COLS + 8)
                            + sizeof(u16) * (size_t)(SST_MAX_NNZ + 2 * SST_MAX_COLS + 16);

struct SstShared
{
  double* vals; // nnz
  double* vec;  // k + r (sweeps) / r * r (factorization)
  int* lptr;    // nlev + 1
  u16* colptr;  // k + 1
  u16* rows;    // nnz
  u16* lcol;    // k
};

// Bulk copy global -> shared in 16-byte pieces with eight loads per thread in flight before the first store: the
// staging of a subtree is a few dozen KB per CTA, and with one load per thread and loop trip it is nothing but memory
// latency (measured: 39 us of a 52 us forward sweep). Both pointers 16-byte aligned, n16 = number of 16-byte pieces
// (the plan pads every segment so that reading up to the next multiple of 16 bytes is safe).
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

// carves the dynamic shared memory and loads the index structure and the values (segments padded to 16 bytes, see
// symbolic.cpp); the caller fills vec and synchronises
__device__ __forceinline__ SstShared
sst_stage(double* smem, const SstMeta& M, const u16* __restrict__ colptr_all, const u16* __restrict__ rows_all, const int* __restrict__ lvl_ptr_all,
          const u16* __restrict__ lvl_col_all, const double* __restrict__ Lg, int vec_len)
{
  SstShared S;
  const int nv = (M.nnz + 1) & ~1, nx = (vec_len + 1) & ~1, nl = (M.nlev + 1 + 3) & ~3, nc = (M.k + 1 + 7) & ~7, nr = (M.nnz + 7) & ~7, nk = (M.k + 7) & ~7;
  S.vals   = smem;
  S.vec    = S.vals + nv;
  S.lptr   = reinterpret_cast<int*>(S.vec + nx);
  S.colptr = reinterpret_cast<u16*>(S.lptr + nl);
  S.rows   = S.colptr + nc;
  S.lcol   = S.rows + nr;
  stage16(S.vals, Lg, nv / 2);
  stage16(S.lptr, lvl_ptr_all + M.lvl_ptr, nl / 4);
  stage16(S.colptr, colptr_all + M.col_ptr, nc / 8);
  stage16(S.rows, rows_all + M.row_ptr, nr / 8);
  stage16(S.lcol, lvl_col_all + M.lvl_col, nk / 8);
  return S;
}

__device__ __forceinline__ void
smem_add(double* p, double v)
{
  atomicAdd(p, v); // shared-memory FP64 atomic
}

} // namespace

__global__ void __launch_bounds__(SST_THREADS)
k_sst_factor(const SstMeta* __restrict__ metas,
             const u16* __restrict__ colptr_all,
             const u16* __restrict__ rows_all,
             const int* __restrict__ lvl_ptr_all,
             const u16* __restrict__ lvl_col_all,
             double* __restrict__ L,
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
  const SstShared S = sst_stage(sst_smem, M, colptr_all, rows_all, lvl_ptr_all, lvl_col_all, Lg, r * r); // vals = assembled entries of S, zeros in the fill
  double* vals      = S.vals;
  double* Us        = S.vec;
  for (int q = threadIdx.x; q < r * r; q += blockDim.x)
  {
    Us[q] = 0.0;
  }
  const double tau = scal[1];
  int nper         = 0;
  __syncthreads();
  for (int lev = 0; lev < M.nlev; ++lev)
  {
    for (int q = S.lptr[lev] + threadIdx.x; q < S.lptr[lev + 1]; q += blockDim.x)
    {
      const int j  = S.lcol[q];
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
      // right-looking update: A[ia, ib] -= f_a f_b / d for the entries a >= b of the column (all targets belong to
      // ancestors of j, i.e. to later levels or to the update block; columns of one level may share a target: atomics)
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
    __syncthreads();
  }
  for (int q = threadIdx.x; q < M.nnz; q += blockDim.x)
  {
    Lg[q] = vals[q];
  }
  double* Ug = U + M.Uoff;
  for (int q = threadIdx.x; q < r * r; q += blockDim.x)
  {
    Ug[q] = Us[q]; // the whole block: nobody zeroes the update matrix of a leaf
  }
  if (nper)
  {
    atomicAdd(n_perturbed, nper);
  }
}

// forward: inside the subtree y = L^-1 b level by level (a column that is done pushes its multiples down its entries),
// yf = D^-1 y; the rows of the ancestors receive their share through atomics on the global accumulator, then the
// parent's dependency counter is signalled (the dataflow kernel that follows waits on it like on any child)
__global__ void __launch_bounds__(SST_THREADS)
k_sst_forward(const SstMeta* __restrict__ metas,
              const u16* __restrict__ colptr_all,
              const u16* __restrict__ rows_all,
              const int* __restrict__ lvl_ptr_all,
              const u16* __restrict__ lvl_col_all,
              const int* __restrict__ Ridx,
              const double* __restrict__ L,
              const double* __restrict__ Dinv,
              double* __restrict__ yacc,
              double* __restrict__ yf,
              int* __restrict__ cnt)
{
  extern __shared__ double sst_smem[];
  const SstMeta M = metas[blockIdx.x];
  const int k = M.k, r = M.r;
  const SstShared S = sst_stage(sst_smem, M, colptr_all, rows_all, lvl_ptr_all, lvl_col_all, L + M.Lptr, k + r);
  double* x         = S.vec;
  {
    double t[(SST_MAX_COLS + SST_MAX_TAIL + SST_THREADS - 1) / SST_THREADS]; // all loads of the right-hand side first
#pragma unroll
    for (int u = 0; u < (int)(sizeof(t) / sizeof(double)); ++u)
    {
      const int q = u * SST_THREADS + threadIdx.x;
      t[u]        = q < k ? yacc[M.first + q] : 0.0;
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
  for (int lev = 0; lev < M.nlev; ++lev)
  {
    for (int q = S.lptr[lev] + threadIdx.x; q < S.lptr[lev + 1]; q += blockDim.x)
    {
      const int j    = S.lcol[q];
      const double y = x[j];
      for (int a = S.colptr[j] + 1; a < S.colptr[j + 1]; ++a)
      {
        smem_add(x + S.rows[a], -S.vals[a] * y);
      }
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < k; q += blockDim.x)
  {
    yf[M.first + q] = x[q] * Dinv[M.first + q];
  }
  for (int q = threadIdx.x; q < r; q += blockDim.x)
  {
    atomicAdd(yacc + Ridx[M.Rptr + q], x[k + q]);
  }
  if (M.parent >= 0)
  {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
    {
      atomicAdd(cnt + M.parent, 1);
    }
  }
}

// backward: x_j = yf_j - sum_i l_ij x_i over the entries of column j, levels from the root of the subtree down
__global__ void __launch_bounds__(SST_THREADS)
k_sst_backward(const SstMeta* __restrict__ metas,
               const u16* __restrict__ colptr_all,
               const u16* __restrict__ rows_all,
               const int* __restrict__ lvl_ptr_all,
               const u16* __restrict__ lvl_col_all,
               const int* __restrict__ Ridx,
               const double* __restrict__ L,
               const double* __restrict__ yf,
               double* __restrict__ xg)
{
  extern __shared__ double sst_smem[];
  const SstMeta M = metas[blockIdx.x];
  const int k = M.k, r = M.r;
  const SstShared S = sst_stage(sst_smem, M, colptr_all, rows_all, lvl_ptr_all, lvl_col_all, L + M.Lptr, k + r);
  double* x         = S.vec;
  {
    double t[(SST_MAX_COLS + SST_MAX_TAIL + SST_THREADS - 1) / SST_THREADS];
#pragma unroll
    for (int u = 0; u < (int)(sizeof(t) / sizeof(double)); ++u)
    {
      const int q = u * SST_THREADS + threadIdx.x;
      t[u]        = q < k ? yf[M.first + q] : (q < k + r ? xg[Ridx[M.Rptr + q - k]] : 0.0);
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
  for (int lev = M.nlev - 1; lev >= 0; --lev)
  {
    for (int q = S.lptr[lev] + threadIdx.x; q < S.lptr[lev + 1]; q += blockDim.x)
    {
      const int j = S.lcol[q];
      double s    = x[j];
      for (int a = S.colptr[j] + 1; a < S.colptr[j + 1]; ++a)
      {
        s -= S.vals[a] * x[S.rows[a]];
      }
      x[j] = s;
    }
    __syncthreads();
  }
  for (int q = threadIdx.x; q < k; q += blockDim.x)
  {
    xg[M.first + q] = x[q];
  }
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
  const int n = (int)dp.plan->sst.size();
  if (n == 0)
  {
    return;
  }
  k_sst_factor<<<n, SST_THREADS, dp.plan->sst_smem_bytes, stream>>>(dp.sst.p, dp.sst_colptr16.p, dp.sst_rows16.p, dp.sst_lvl_ptr.p, dp.sst_lvl_col16.p, nb.L, nb.U, nb.D, nb.Dinv, nb.scal,
                                                     nb.n_perturbed);
  lc.tick("sst");
  B200_CUDA(cudaGetLastError());
}

void
enqueue_sst_forward(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc)
{
  const int n = (int)dp.plan->sst.size();
  if (n == 0)
  {
    return;
  }
  k_sst_forward<<<n, SST_THREADS, dp.plan->sst_smem_bytes, stream>>>(dp.sst.p, dp.sst_colptr16.p, dp.sst_rows16.p, dp.sst_lvl_ptr.p, dp.sst_lvl_col16.p, dp.Ridx.p, nb.L, nb.Dinv, sb.y, sb.yf,
                                                  sb.flow);
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
  k_sst_backward<<<n, SST_THREADS, dp.plan->sst_smem_bytes, stream>>>(dp.sst.p, dp.sst_colptr16.p, dp.sst_rows16.p, dp.sst_lvl_ptr.p, dp.sst_lvl_col16.p, dp.Ridx.p, nb.L, sb.yf, sb.x);
  lc.tick();
  B200_CUDA(cudaGetLastError());
}

} // namespace b200
