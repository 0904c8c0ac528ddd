// cg.cu -- device-resident Steihaug projected CG: the EQP trust-region
// solve  min g^T p + 1/2 p^T H p  s.t.  A_W p = 0, |p| <= Delta  with every vector kept in HBM.
//
// Restates the reference's loop, src/main/tr/steihaug_solver.c:223-496 (projection =
// sleqp_aug_jac_project_nullspace, standard_aug_jac.c:396-435; boundary step =
// sleqp_tr_compute_bdry_sol, tr/tr_util.c:9-50), iterate for iterate: same recurrences, same three
// exits (interior / boundary / negative curvature) and the same behaviour at the iteration cap (the
// reference leaves the step at zero, steihaug_solver.c:302-305). What changes is where the data lives:
// the reference pays, per iteration, a sparse-vector round trip through the SleqpFact boundary (H2D
// right-hand side, D2H solution, host sparsification); here one iteration is 1 SpMV + 1 KKT solve +
// a few fused vector kernels on the factorization's stream; the scalars of the recurrences stay on the
// device and the host reads nine of them once per iteration to check the exits.
#include "cg.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

using namespace b200;

namespace b200
{

// out[0] += x.y, out[1] += x.x, out[2] += y.y   (out zeroed by the caller)
__global__ void
k_dot3(int n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out)
{
  double xy = 0.0, xx = 0.0, yy = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const double a = x[i], b = y[i];
    xy += a * b;
    xx += a * a;
    yy += b * b;
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    xy += __shfl_xor_sync(0xffffffffu, xy, o);
    xx += __shfl_xor_sync(0xffffffffu, xx, o);
    yy += __shfl_xor_sync(0xffffffffu, yy, o);
  }
  if ((threadIdx.x & 31) == 0)
  {
    atomicAdd(out + 0, xy);
    atomicAdd(out + 1, xx);
    atomicAdd(out + 2, yy);
  }
}

// out = a * x + b * y (out may alias x or y)
__global__ void
k_axpby(int n, double a, const double* x, double b, const double* y, double* out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out[i] = a * x[i] + b * y[i];
  }
}

// out = a * x + (*b) * y with the second coefficient in device memory (out may alias x or y)
__global__ void
k_axpby_dev(int n, double a, const double* x, const double* __restrict__ b, const double* y, double* out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out[i] = a * x[i] + (*b) * y[i];
  }
}

// the scalars of one CG iteration stay on the device (layout: see b200_cg_solve)
__global__ void
k_cg_alpha(double* S)
{
  S[10] = S[9] / S[0]; // alpha = r.g / d.Bd                            (steihaug_solver.c:405)
}

__global__ void
k_cg_beta(double* S)
{
  S[11] = S[6] / S[9]; // beta = (r.g)_new / (r.g)_old                   (:467-469)
  S[9]  = S[6];
}

// ---- device-controlled loop (Hessian on the device): the host enqueues iterations ahead, the exit tests run in the
// last block of the reductions and every kernel of the CG checks the exit flag first, so that the state of the exit
// iteration (z, d, B d) survives whatever has been enqueued behind it.
// Scalars S: [0..2] d.Bd, d.d, Bd.Bd   [3..5] z+.d, z+.z+, d.d   [6..8] r.g, r.r, g.g   [9] current r.g   [10] alpha
// [11] beta   [12] |z|^2   [13] min Rayleigh   [14] max Rayleigh   [15] iterations completed
// Exit record E (ints): [0] flag (0 running, 1 + B200_CG_* otherwise)   [1] iteration of the exit
// Exit scalars X: [0] d.Bd   [1] d.d   [2] |z|^2 at the exit
constexpr int CGD_BLOCKS = 1184; // 8 per SM; four elements per thread and trip
enum
{
  CGD_CURVATURE, // (x, y) = (d, B d): Rayleigh bounds (:150-173), negative-curvature exit (:349), alpha (:405)
  CGD_RADIUS,    // z+ = z + alpha d written, (x, y) = (z+, d): boundary exit (:419)
  CGD_TAIL       // (x, y) = (r, g): beta (:467-469), |z|^2 of the accepted iterate, iteration count, and the
                 // convergence test the reference makes at the top of the next pass (:318-327)
};

// (x.y, x.x, y.y) -> S[3 * MODE ...] and the scalar logic that follows, in one launch: every block leaves its partial
// sums in `part`, the block that arrives last adds them up in a fixed order (the result does not depend on which block
// that is, nor on the order of arrival: bit-reproducible, unlike atomics on the three sums) and runs the exit test.
template <int MODE>
__global__ void __launch_bounds__(256)
k_cgd_reduce(int n,
             const double* __restrict__ x,
             const double* __restrict__ y,
             double* __restrict__ xout,
             double* __restrict__ res,       // CGD_RADIUS: r += alpha * Bd rides along (same alpha, same pass) ...
             const double* __restrict__ Bd,  // ... (:449-456); the reference updates r after the boundary test, but r is dead on that exit
             double* __restrict__ S,
             int* __restrict__ E,
             double* __restrict__ X,
             double* __restrict__ part,
             unsigned* __restrict__ arrived,
             double param, // CGD_RADIUS: radius^2   CGD_TAIL: tol^2, negative = no convergence test (iteration cap behind)
             const int nblocks)
{
  if (*E)
  {
    return;
  }
  __shared__ double red[8][3];
  __shared__ bool last;
  double xy = 0.0, xx = 0.0, yy = 0.0;
  const double alpha = MODE == CGD_RADIUS ? S[10] : 0.0;
  // four elements per thread and trip, all loads before the first use (the grid covers n in one or two trips: with one
  // element per trip a thread sat out a full DRAM latency per element, 11-15 us per reduction for 16-48 MB)
  const int stride = gridDim.x * blockDim.x;
  for (int i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride)
  {
    double a[4], b[4], rr[4], bd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int i = i0 + u * stride;
      a[u]        = i < n ? x[i] : 0.0;
      b[u]        = i < n ? y[i] : 0.0;
      if (MODE == CGD_RADIUS)
      {
        rr[u] = i < n ? res[i] : 0.0;
        bd[u] = i < n ? Bd[i] : 0.0;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int i = i0 + u * stride;
      if (MODE == CGD_RADIUS)
      {
        a[u] += alpha * b[u];
        if (i < n)
        {
          xout[i] = a[u];
          res[i]  = rr[u] + alpha * bd[u];
        }
      }
      xy += a[u] * b[u];
      xx += a[u] * a[u];
      yy += b[u] * b[u];
    }
  }
  auto block_sum = [&]() { // fixed shape: butterfly inside the warp, the eight warps in order
    for (int o = 16; o > 0; o >>= 1)
    {
      xy += __shfl_xor_sync(0xffffffffu, xy, o);
      xx += __shfl_xor_sync(0xffffffffu, xx, o);
      yy += __shfl_xor_sync(0xffffffffu, yy, o);
    }
    if ((threadIdx.x & 31) == 0)
    {
      red[threadIdx.x >> 5][0] = xy;
      red[threadIdx.x >> 5][1] = xx;
      red[threadIdx.x >> 5][2] = yy;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      xy = xx = yy = 0.0;
      for (int w = 0; w < 8; ++w)
      {
        xy += red[w][0];
        xx += red[w][1];
        yy += red[w][2];
      }
    }
  };
  block_sum();
  if (threadIdx.x == 0)
  {
    part[3 * blockIdx.x + 0] = xy;
    part[3 * blockIdx.x + 1] = xx;
    part[3 * blockIdx.x + 2] = yy;
    __threadfence();
    last = atomicAdd(arrived, 1u) == (unsigned)nblocks - 1;
  }
  __syncthreads();
  if (!last)
  {
    return;
  }
  __threadfence();
  xy = xx = yy = 0.0;
  for (int q = threadIdx.x; q < nblocks; q += blockDim.x)
  {
    xy += __ldcg(part + 3 * q + 0);
    xx += __ldcg(part + 3 * q + 1);
    yy += __ldcg(part + 3 * q + 2);
  }
  __syncthreads(); // red is reused
  block_sum();
  if (threadIdx.x != 0)
  {
    return;
  }
  *arrived        = 0;
  S[3 * MODE + 0] = xy;
  S[3 * MODE + 1] = xx;
  S[3 * MODE + 2] = yy;
  const int iteration = (int)S[15];
  if (MODE == CGD_CURVATURE)
  {
    const double dBd = xy, dd = xx;
    if (dd != 0.0)
    {
      S[13] = fmin(S[13], dBd / dd);
      S[14] = fmax(S[14], dBd / dd);
    }
    if (dBd <= 0.0)
    {
      E[0] = 1 + B200_CG_NEG_CURVATURE;
      E[1] = iteration;
      X[0] = dBd;
      X[1] = dd;
      X[2] = S[12];
      return;
    }
    S[10] = S[9] / dBd;
  }
  if (MODE == CGD_RADIUS)
  {
    if (xx >= param)
    {
      E[0] = 1 + B200_CG_BOUNDARY;
      E[1] = iteration;
      X[0] = S[0];
      X[1] = S[1];
      X[2] = S[12];
    }
  }
  if (MODE == CGD_TAIL)
  {
    S[11] = xy / S[9];
    S[9]  = xy;
    S[12] = S[4];
    S[15] += 1.0;
    if (param >= 0.0 && fabs(xy) < param) // what the reference tests at the top of the next pass
    {
      E[0] = 1 + B200_CG_INTERIOR;
      E[1] = iteration + 1;
    }
  }
}

__global__ void
k_cgd_axpby(int n, double a, const double* x, const double* __restrict__ b, const double* y, double* out, const int* __restrict__ flag)
{
  if (*flag)
  {
    return;
  }
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out[i] = a * x[i] + (*b) * y[i];
  }
}

} // namespace b200

struct b200_cg
{
  b200_fact* fact = nullptr;
  b200_mat* hess  = nullptr; // Hessian of the Lagrangian resident on the device, or
  b200_hess_prod_fn hess_cb = nullptr; // ... a host callback (the reference's matrix-free SLEQP_FUNC_HESS_PROD, pub_func.h:168)
  void* hess_ctx            = nullptr;
  PinnedBuf<double> h_d, h_Bd; // staging of the callback mode
  int device      = 0;
  cudaStream_t stream = nullptr;
  int n = 0, N = 0;
  DevBuf<double> z, znext, d, dnext, Bd, scal;
  DevBuf<int> g_idx, exit_rec, cp_idx, cp_cnt; // cp_*: compaction of the sparse result
  DevBuf<double> cp_val;
  DevBuf<double> part;       // partial sums of the reductions of the device-controlled loop
  DevBuf<unsigned> arrived;  // ... and their arrival counter (zero between launches)
  PinnedBuf<int> h_exit;
  DevBuf<double> g_val;
  PinnedBuf<double> h_scal, h_step, h_val;
  PinnedBuf<int> h_idx;
};

namespace
{

template <typename F>
int
guarded(F&& f)
{
  try
  {
    return f();
  }
  catch (const CudaError& e)
  {
    return set_error(B200_ERR_CUDA, e.what());
  }
  catch (const std::exception& e)
  {
    return set_error(B200_ERR_CUDA, e.what());
  }
}

inline unsigned
nb(int n)
{
  return (unsigned)((n + 255) / 256);
}

// host <- (x.y, x.x, y.y); one stream synchronisation
void
dot3(b200_cg* C, const double* x, const double* y, double out[3])
{
  B200_CUDA(cudaMemsetAsync(C->scal.p, 0, 3 * sizeof(double), C->stream));
  k_dot3<<<std::min(nb(C->n), 592u), 256, 0, C->stream>>>(C->n, x, y, C->scal.p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaMemcpyAsync(C->h_scal.p, C->scal.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, C->stream));
  B200_CUDA(cudaStreamSynchronize(C->stream));
  out[0] = C->h_scal.p[0];
  out[1] = C->h_scal.p[1];
  out[2] = C->h_scal.p[2];
}

void
axpby(b200_cg* C, double a, const double* x, double b, const double* y, double* out)
{
  k_axpby<<<nb(C->n), 256, 0, C->stream>>>(C->n, a, x, b, y, out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

} // namespace

extern "C" {

int
b200_cg_create(b200_cg** handle, b200_fact* fact, b200_mat* hess)
{
  if (!handle || !fact)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  *handle = nullptr;
  return guarded([&]() {
    std::unique_ptr<b200_cg> C(new b200_cg());
    C->fact   = fact;
    C->hess   = hess;
    C->stream = (cudaStream_t)b200_fact_stream(fact);
    C->device = b200_fact_device(fact); // everything of the CG lives next to the factorization it projects with
    B200_CUDA(cudaSetDevice(C->device));
    // Hessian products run on the factorization's stream: they alternate with its solves
    if (hess)
    {
      int rc = b200_mat_set_stream(hess, (void*)C->stream);
      if (rc != B200_OK)
      {
        return rc;
      }
    }
    C->scal.reserve(32);
    C->h_scal.reserve(32);
    *handle = C.release();
    return (int)B200_OK;
  });
}

int
b200_cg_set_hess_callback(b200_cg* C, b200_hess_prod_fn fn, void* ctx)
{
  if (!C)
  {
    return set_error(B200_ERR_ARG, "null handle");
  }
  C->hess_cb  = fn;
  C->hess_ctx = ctx;
  return B200_OK;
}

int
b200_cg_solve(b200_cg* C,
              int n,
              int nnz_g,
              const int* g_idx,
              const double* g_val,
              double trust_radius,
              double rel_tol,
              int max_iter,
              double* step_out,
              int* iterations,
              int* termination)
{
  return b200_cg_solve_ex(C, n, nnz_g, g_idx, g_val, trust_radius, rel_tol, max_iter, step_out, iterations, termination, nullptr, nullptr, nullptr);
}

namespace
{
// sparse result (b200_cg_solve_sparse): entries with |v| > zero_eps, ascending, compacted on the device
struct SparseStep
{
  double zero_eps;
  int* idx;
  double* val;
  int* nnz;
};

int
cg_solve_impl(b200_cg* C,
              int n,
              int nnz_g,
              const int* g_idx,
              const double* g_val,
              double trust_radius,
              double rel_tol,
              int max_iter,
              double* step_out,
              const SparseStep* sparse,
              int* iterations,
              int* termination,
              double* tr_dual,
              double* min_rayleigh,
              double* max_rayleigh)
{
  if (!C || (!step_out && !sparse) || (sparse && (!sparse->idx || !sparse->val || !sparse->nnz)) || n <= 0 || nnz_g < 0 || nnz_g > n || (nnz_g > 0 && (!g_idx || !g_val)))
  {
    return set_error(B200_ERR_ARG, "bad argument");
  }
  if (!C->hess && !C->hess_cb)
  {
    return set_error(B200_ERR_STATE, "no Hessian: neither a device matrix nor a host callback is set");
  }
  // steihaug_solver.c:234-235, 246: both Rayleigh bounds start at 1, the dual of the trust region is only set on the
  // boundary exit
  double ray_min = 1.0, ray_max = 1.0;
  if (tr_dual)
  {
    *tr_dual = NAN; // "not computed" (the reference: SLEQP_NONE); the glue translates
  }
  b200_stats st;
  int rc = b200_fact_stats(C->fact, &st);
  if (rc != B200_OK)
  {
    return rc;
  }
  if (n > st.n)
  {
    return set_error(B200_ERR_ARG, "number of variables exceeds the order of the factorized KKT matrix");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(C->device));
    const int N = st.n;
    C->n = n;
    C->N = N;
    cudaStream_t s = C->stream;
    C->z.reserve((size_t)n);
    C->znext.reserve((size_t)n);
    C->d.reserve((size_t)n);
    C->dnext.reserve((size_t)n);
    C->Bd.reserve((size_t)n);
    // [r; 0], the right-hand side of the projection (tail stays zero), and K^-1 [r; 0], whose first n entries are
    // g = P r, live in the buffers the solve graph of the factorization works on: no copies around the projection
    double *rfull = nullptr, *gfull = nullptr;
    {
      int brc = b200_fact_device_buffers(C->fact, &rfull, &gfull);
      if (brc != B200_OK)
      {
        return brc;
      }
    }
    C->h_step.reserve((size_t)n);
    C->h_exit.reserve(4);
    B200_CUDA(cudaStreamSynchronize(s));
    B200_CUDA(cudaMemsetAsync(C->z.p, 0, sizeof(double) * (size_t)n, s));
    B200_CUDA(cudaMemsetAsync(rfull, 0, sizeof(double) * (size_t)N, s));
    // r0 = gradient (sparse host vector -> dense device vector)
    bool g_contig = false;
    int g_first   = 0;
    if (nnz_g > 0)
    {
      C->h_val.reserve((size_t)nnz_g);
      C->h_idx.reserve((size_t)nnz_g);
      C->g_val.reserve((size_t)nnz_g);
      C->g_idx.reserve((size_t)nnz_g);
      const double* src_val = g_val;
      const int* src_idx    = g_idx;
      // ascending indices (pub_vec.h:13-14): a dense gradient is a contiguous run and its indices stay on the host
      g_contig = (g_idx[nnz_g - 1] - g_idx[0]) == nnz_g - 1;
      g_first  = g_idx[0];
      const bool pinned_in = is_pinned_host(g_val, sizeof(double) * (size_t)nnz_g) && (g_contig || is_pinned_host(g_idx, sizeof(int) * (size_t)nnz_g));
      if (!pinned_in) // page-locked caller arrays are DMA'd as they are
      {
        std::memcpy(C->h_val.p, g_val, sizeof(double) * (size_t)nnz_g);
        src_val = C->h_val.p;
        if (!g_contig)
        {
          std::memcpy(C->h_idx.p, g_idx, sizeof(int) * (size_t)nnz_g);
          src_idx = C->h_idx.p;
        }
      }
      B200_CUDA(cudaMemcpyAsync(C->g_val.p, src_val, sizeof(double) * (size_t)nnz_g, cudaMemcpyHostToDevice, s));
      if (!g_contig)
      {
        B200_CUDA(cudaMemcpyAsync(C->g_idx.p, src_idx, sizeof(int) * (size_t)nnz_g, cudaMemcpyHostToDevice, s));
      }
      LaunchCounter lc;
      enqueue_scatter_rhs(rfull, N, nnz_g, g_contig ? nullptr : C->g_idx.p, g_first, C->g_val.p, s, lc);
    }
    double* r = rfull;
    double* g = gfull;
    auto project = [&]() -> int { return b200_fact_solve_device(C->fact, rfull, gfull); };
    // out = H v: a device SpMV, or -- matrix-free Hessians -- one round trip through the host callback
    auto hess_prod = [&](const double* v_dev, double* out_dev) -> int {
      if (C->hess)
      {
        return b200_mat_mult_vec_device(C->hess, v_dev, out_dev);
      }
      C->h_d.reserve((size_t)n);
      C->h_Bd.reserve((size_t)n);
      B200_CUDA(cudaMemcpyAsync(C->h_d.p, v_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      if (C->hess_cb(C->hess_ctx, n, C->h_d.p, C->h_Bd.p) != 0)
      {
        return set_error(B200_ERR_ARG, "Hessian callback failed");
      }
      B200_CUDA(cudaMemcpyAsync(out_dev, C->h_Bd.p, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, s));
      return (int)B200_OK;
    };
    auto report_rayleigh = [&]() {
      if (min_rayleigh)
      {
        *min_rayleigh = ray_min;
      }
      if (max_rayleigh)
      {
        *max_rayleigh = ray_max;
      }
    };
    auto finish  = [&](const double* p_dev, int iters, int term) {
      if (sparse)
      {
        *sparse->nnz = 0;
      }
      if (p_dev && sparse && is_pinned_host(sparse->idx, sizeof(int) * (size_t)n) && is_pinned_host(sparse->val, sizeof(double) * (size_t)n))
      {
        // sparsified on the device (sleqp_vec_set_from_raw, vec.c:72-104), both arrays DMA'd into the caller's buffers
        const int nchunks = compact_chunks(n);
        C->cp_idx.reserve((size_t)n);
        C->cp_val.reserve((size_t)n);
        C->cp_cnt.reserve((size_t)nchunks + 1);
        LaunchCounter eager;
        enqueue_compact(p_dev, n, sparse->zero_eps, C->cp_cnt.p, C->cp_idx.p, C->cp_val.p, s, eager);
        B200_CUDA(cudaMemcpyAsync(C->h_exit.p + 2, C->cp_cnt.p + nchunks, sizeof(int), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaMemcpyAsync(sparse->val, C->cp_val.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaMemcpyAsync(sparse->idx, C->cp_idx.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        *sparse->nnz = C->h_exit.p[2];
      }
      else if (p_dev)
      {
        B200_CUDA(cudaMemcpyAsync(C->h_step.p, p_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        if (sparse)
        {
          int nnz = 0;
          for (int i = 0; i < n; ++i)
          {
            const double v = C->h_step.p[i];
            if (std::fabs(v) > sparse->zero_eps)
            {
              sparse->idx[nnz] = i;
              sparse->val[nnz] = v;
              ++nnz;
            }
          }
          *sparse->nnz = nnz;
        }
        else
        {
          std::memcpy(step_out, C->h_step.p, sizeof(double) * (size_t)n);
        }
      }
      else if (step_out)
      {
        std::memset(step_out, 0, sizeof(double) * (size_t)n);
      }
      if (iterations)
      {
        *iterations = iters;
      }
      if (termination)
      {
        *termination = term;
      }
      report_rayleigh();
      return (int)B200_OK;
    };
    // dual of the trust-region constraint on the boundary exit (steihaug_tr_dual, steihaug_solver.c:187-221):
    // max(0, -(p^T H p + p^T grad)) / radius^2 with one more Hessian product
    auto boundary_dual = [&](const double* p_dev, double* scratch_grad, double* scratch_Hp) -> int {
      if (!tr_dual)
      {
        return (int)B200_OK;
      }
      int hrc = hess_prod(p_dev, scratch_Hp);
      if (hrc != B200_OK)
      {
        return hrc;
      }
      B200_CUDA(cudaMemsetAsync(scratch_grad, 0, sizeof(double) * (size_t)n, s));
      if (nnz_g > 0)
      {
        LaunchCounter lc;
        enqueue_scatter_rhs(scratch_grad, n, nnz_g, g_contig ? nullptr : C->g_idx.p, g_first, C->g_val.p, s, lc);
      }
      double a[3], b[3];
      dot3(C, p_dev, scratch_Hp, a);
      dot3(C, p_dev, scratch_grad, b);
      const double comb = a[0] + b[0];
      *tr_dual          = comb < 0 ? -comb / (trust_radius * trust_radius) : 0.0;
      return (int)B200_OK;
    };

    // g0 = P[r0], d0 = -g0                                              (steihaug_solver.c:270-277)
    int prc = project();
    if (prc != B200_OK)
    {
      return prc;
    }
    axpby(C, -1.0, g, 0.0, g, C->d.p);
    double sc[3];
    dot3(C, r, g, sc); // r.g, r.r, g.g
    double r_dot_g         = sc[0];
    const double d_nrm_sq0 = sc[2];
    const double tol_sq    = rel_tol * rel_tol;
    if (d_nrm_sq0 < tol_sq) //                                              (:280-285)
    {
      return finish(C->z.p, 0, B200_CG_INTERIOR);
    }
    if (C->hess)
    {
      // ---- Hessian on the device: the whole loop runs ahead of the host, CHUNK iterations per synchronisation ----
      constexpr int CHUNK = 8;
      double* S = C->scal.p;
      C->exit_rec.reserve(4);
      C->h_exit.reserve(4);
      int* E    = C->exit_rec.p;
      double* X = S + 16;
      {
        double init[20] = {0};
        init[9]  = r_dot_g;
        init[13] = 1.0;
        init[14] = 1.0;
        std::memcpy(C->h_scal.p, init, sizeof(init));
        B200_CUDA(cudaMemcpyAsync(S, C->h_scal.p, sizeof(init), cudaMemcpyHostToDevice, s));
        B200_CUDA(cudaMemsetAsync(E, 0, 4 * sizeof(int), s));
      }
      if (max_iter != 0 && std::fabs(r_dot_g) < tol_sq)
      {
        return finish(C->z.p, 0, B200_CG_INTERIOR); // the test at the top of the first pass (:318-327), known on the host
      }
      double* zb[2] = {C->z.p, C->znext.p};
      double* db[2] = {C->d.p, C->dnext.p};
      const unsigned gb = nb(n);
      if (C->part.cap < (size_t)(3 * CGD_BLOCKS))
      {
        C->part.reserve(3 * CGD_BLOCKS);
        C->arrived.reserve(2);
        B200_CUDA(cudaMemsetAsync(C->arrived.p, 0, 2 * sizeof(unsigned), s));
      }
      double* const part      = C->part.p;
      unsigned* const arrived = C->arrived.p;
      const int gr            = (int)std::min((nb(n) + 3u) / 4u, (unsigned)CGD_BLOCKS);
      auto enqueue_iteration = [&](int it) -> int {
        double* z_cur = zb[it & 1];
        double* z_new = zb[(it + 1) & 1];
        double* d_cur = db[it & 1];
        double* d_new = db[(it + 1) & 1];
        int hrc = b200_mat_mult_vec_device_if(C->hess, d_cur, C->Bd.p, E); //       (:339)
        if (hrc != B200_OK)
        {
          return hrc;
        }
        k_cgd_reduce<CGD_CURVATURE><<<gr, 256, 0, s>>>(n, d_cur, C->Bd.p, nullptr, nullptr, nullptr, S, E, X, part, arrived, 0.0, gr);
        // z+ = z + alpha d, r += alpha B d
        k_cgd_reduce<CGD_RADIUS><<<gr, 256, 0, s>>>(n, z_cur, d_cur, z_new, r, C->Bd.p, S, E, X, part, arrived, trust_radius * trust_radius, gr);
        int prc2 = project();                                              // g = P[r]        (:459)
        if (prc2 != B200_OK)
        {
          return prc2;
        }
        // the convergence test of the next pass rides on this reduction, unless the iteration cap comes first (:302)
        const bool cap_next = max_iter >= 0 && it + 1 >= max_iter;
        k_cgd_reduce<CGD_TAIL><<<gr, 256, 0, s>>>(n, r, g, nullptr, nullptr, nullptr, S, E, X, part, arrived, cap_next ? -1.0 : tol_sq, gr);
        k_cgd_axpby<<<gb, 256, 0, s>>>(n, -1.0, g, S + 11, d_cur, d_new, E); // d+ = -g + beta d (:472-479)
        g_launches.fetch_add(4, std::memory_order_relaxed);
        return (int)B200_OK;
      };
      int enq = 0;
      for (;;)
      {
        const int upto = max_iter >= 0 ? std::min(max_iter, enq + CHUNK) : enq + CHUNK;
        for (; enq < upto; ++enq)
        {
          int erc = enqueue_iteration(enq);
          if (erc != B200_OK)
          {
            return erc;
          }
        }
        if (max_iter >= 0 && enq == max_iter)
        {
          // the loop-top test of the iteration behind the cap (an interior exit there still counts, :318 precedes :302
          // only for the NEXT pass: the reference tests the cap first, so nothing more is enqueued)
        }
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaMemcpyAsync(C->h_exit.p, E, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaMemcpyAsync(C->h_scal.p, S, 20 * sizeof(double), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        const int flag = C->h_exit.p[0];
        ray_min = C->h_scal.p[13];
        ray_max = C->h_scal.p[14];
        if (flag)
        {
          const int it   = C->h_exit.p[1];
          const int kind = flag - 1;
          double* z_cur  = zb[it & 1];
          double* z_new  = zb[(it + 1) & 1];
          double* d_cur  = db[it & 1];
          double* d_new  = db[(it + 1) & 1];
          if (kind == B200_CG_INTERIOR)
          {
            return finish(z_cur, it, B200_CG_INTERIOR);
          }
          const double dBd = C->h_scal.p[16], d_nrm_sq = C->h_scal.p[17], z_nrm_sq_x = C->h_scal.p[18];
          if (kind == B200_CG_NEG_CURVATURE) //                                 (:349-402)
          {
            double zs[3];
            dot3(C, z_cur, d_cur, zs);
            const double z_dot_d = zs[0];
            const double inner   = z_dot_d * z_dot_d - d_nrm_sq * (z_nrm_sq_x - trust_radius * trust_radius);
            const double tau_min = 1. / d_nrm_sq * (-z_dot_d - std::sqrt(inner));
            const double tau_max = 1. / d_nrm_sq * (-z_dot_d + std::sqrt(inner));
            double gs[3], zbd[3];
            B200_CUDA(cudaMemsetAsync(z_new, 0, sizeof(double) * (size_t)n, s));
            if (nnz_g > 0)
            {
              LaunchCounter lc;
              enqueue_scatter_rhs(z_new, n, nnz_g, g_contig ? nullptr : C->g_idx.p, g_first, C->g_val.p, s, lc);
            }
            dot3(C, z_new, d_cur, gs);
            dot3(C, z_cur, C->Bd.p, zbd);
            const double gd_ = gs[0], zBd = zbd[0];
            const double tau_min_obj = tau_min * ((gd_ + zBd) + 0.5 * tau_min * dBd);
            const double tau_max_obj = tau_max * ((gd_ + zBd) + 0.5 * tau_max * dBd);
            const double tau         = (tau_min_obj < tau_max_obj) ? tau_min : tau_max;
            axpby(C, 1.0, z_cur, tau, d_cur, z_new);
            return finish(z_new, it, B200_CG_NEG_CURVATURE);
          }
          // boundary (:419-441, tr_util.c:9-50)
          double zs[3];
          dot3(C, z_cur, d_cur, zs);
          const double prev_dot_d = zs[0];
          const double p_norm = std::sqrt(zs[1]), d_norm = std::sqrt(zs[2]);
          const double inner  = prev_dot_d * prev_dot_d - d_norm * d_norm * (p_norm * p_norm - trust_radius * trust_radius);
          const double factor = 1. / (d_norm * d_norm) * (-prev_dot_d + std::sqrt(inner));
          axpby(C, 1.0, z_cur, factor, d_cur, z_new);
          int drc = boundary_dual(z_new, d_new, C->Bd.p);
          if (drc != B200_OK)
          {
            return drc;
          }
          return finish(z_new, it, B200_CG_BOUNDARY);
        }
        if (max_iter >= 0 && enq >= max_iter)
        {
          // cap reached without an exit: the reference tests the cap before the convergence test of the next pass
          // and returns the zero step (:302-305)
          return finish(nullptr, max_iter, B200_CG_MAX_ITER);
        }
      }
    }
    // ---- matrix-free Hessian (host callback): one host synchronisation per iteration ------------------------------
    // The scalars of the recurrences stay on the device:
    //   S[0..2] = (d.Bd, d.d, Bd.Bd)   S[3..5] = (z+.d, z+.z+, d.d)   S[6..8] = (r.g, r.r, g.g) after the update
    //   S[9] = current r.g   S[10] = alpha   S[11] = beta
    // so the whole iteration -- SpMV, step, residual update, projection, new direction -- is enqueued without
    // waiting for the host, which then reads S[0..8] once and checks the three exits in the reference's order. The
    // iterate and the direction are double-buffered: what an exit needs (z, d, Bd of the iteration) is still intact
    // although the work after the exit test has already run (and is discarded).
    double* S = C->scal.p;
    B200_CUDA(cudaMemcpyAsync(S + 9, C->h_scal.p, sizeof(double), cudaMemcpyHostToDevice, s)); // h_scal[0] = r.g (dot3 above)
    double* z_cur = C->z.p;
    double* z_new = C->znext.p;
    double* d_cur = C->d.p;
    double* d_new = C->dnext.p;
    auto dot3_async = [&](const double* a, const double* b, double* out) {
      k_dot3<<<std::min(nb(n), 592u), 256, 0, s>>>(n, a, b, out);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    };
    auto axpby_dev = [&](double a, const double* xx, const double* bdev, const double* yy, double* out) {
      k_axpby_dev<<<nb(n), 256, 0, s>>>(n, a, xx, bdev, yy, out);
      g_launches.fetch_add(1, std::memory_order_relaxed);
    };
    double z_nrm_sq = 0.0;
    for (int it = 0;; ++it)
    {
      if (max_iter >= 0 && it >= max_iter) // cap: the reference returns the zero step (:302-305)
      {
        return finish(nullptr, it, B200_CG_MAX_ITER);
      }
      if (std::fabs(r_dot_g) < tol_sq) //                                     (:318-327)
      {
        return finish(z_cur, it, B200_CG_INTERIOR);
      }
      B200_CUDA(cudaMemsetAsync(S, 0, 9 * sizeof(double), s));
      int hrc = hess_prod(d_cur, C->Bd.p); //                                   (:339)
      if (hrc != B200_OK)
      {
        return hrc;
      }
      dot3_async(d_cur, C->Bd.p, S); // d.Bd, d.d, Bd.Bd
      k_cg_alpha<<<1, 1, 0, s>>>(S);
      axpby_dev(1.0, z_cur, S + 10, d_cur, z_new); // z+ = z + alpha d
      dot3_async(z_new, d_cur, S + 3);             // z+.d, z+.z+
      axpby_dev(1.0, r, S + 10, C->Bd.p, r);       // r += alpha B d           (:449-456)
      prc = project();                             // g = P[r]                  (:459)
      if (prc != B200_OK)
      {
        return prc;
      }
      dot3_async(r, g, S + 6);
      k_cg_beta<<<1, 1, 0, s>>>(S);
      axpby_dev(-1.0, g, S + 11, d_cur, d_new); // d+ = -g + beta d           (:472-479)
      g_launches.fetch_add(2, std::memory_order_relaxed);
      B200_CUDA(cudaMemcpyAsync(C->h_scal.p, S, 9 * sizeof(double), cudaMemcpyDeviceToHost, s));
      B200_CUDA(cudaStreamSynchronize(s));
      const double dBd = C->h_scal.p[0], d_nrm_sq = C->h_scal.p[1];
      const double z_next_nrm_sq = C->h_scal.p[4], r_dot_g_new = C->h_scal.p[6];
      if (d_nrm_sq != 0.) // steihaug_collect_rayleigh (:150-173), right after the Hessian product (:345)
      {
        ray_min = std::min(ray_min, dBd / d_nrm_sq);
        ray_max = std::max(ray_max, dBd / d_nrm_sq);
      }
      if (dBd <= 0.0) // negative curvature                                  (:349-402)
      {
        double zs[3];
        dot3(C, z_cur, d_cur, zs); // z.d
        const double z_dot_d = zs[0];
        const double inner   = z_dot_d * z_dot_d - d_nrm_sq * (z_nrm_sq - trust_radius * trust_radius);
        const double tau_min = 1. / d_nrm_sq * (-z_dot_d - std::sqrt(inner));
        const double tau_max = 1. / d_nrm_sq * (-z_dot_d + std::sqrt(inner));
        double gs[3], zb[3];
        // gradient . d with the original sparse gradient (kept on the device)
        B200_CUDA(cudaMemsetAsync(z_new, 0, sizeof(double) * (size_t)n, s));
        if (nnz_g > 0)
        {
          LaunchCounter lc;
          enqueue_scatter_rhs(z_new, n, nnz_g, g_contig ? nullptr : C->g_idx.p, g_first, C->g_val.p, s, lc);
        }
        dot3(C, z_new, d_cur, gs);
        dot3(C, z_cur, C->Bd.p, zb);
        const double gd = gs[0], zBd = zb[0];
        const double tau_min_obj = tau_min * ((gd + zBd) + 0.5 * tau_min * dBd);
        const double tau_max_obj = tau_max * ((gd + zBd) + 0.5 * tau_max * dBd);
        const double tau         = (tau_min_obj < tau_max_obj) ? tau_min : tau_max;
        axpby(C, 1.0, z_cur, tau, d_cur, z_new);
        return finish(z_new, it, B200_CG_NEG_CURVATURE);
      }
      if (z_next_nrm_sq >= trust_radius * trust_radius) // boundary       (:419-441, tr_util.c:9-50)
      {
        double zs[3];
        dot3(C, z_cur, d_cur, zs);
        const double prev_dot_d = zs[0];
        const double p_norm = std::sqrt(zs[1]), d_norm = std::sqrt(zs[2]);
        const double inner  = prev_dot_d * prev_dot_d - d_norm * d_norm * (p_norm * p_norm - trust_radius * trust_radius);
        const double factor = 1. / (d_norm * d_norm) * (-prev_dot_d + std::sqrt(inner));
        axpby(C, 1.0, z_cur, factor, d_cur, z_new);
        int drc = boundary_dual(z_new, d_new, C->Bd.p); // d_new / Bd: scratch from here on
        if (drc != B200_OK)
        {
          return drc;
        }
        return finish(z_new, it, B200_CG_BOUNDARY);
      }
      std::swap(z_cur, z_new);
      std::swap(d_cur, d_new);
      z_nrm_sq = z_next_nrm_sq;
      r_dot_g  = r_dot_g_new;
    }
  });
}
} // namespace

int
b200_cg_solve_ex(b200_cg* C,
                 int n,
                 int nnz_g,
                 const int* g_idx,
                 const double* g_val,
                 double trust_radius,
                 double rel_tol,
                 int max_iter,
                 double* step_out,
                 int* iterations,
                 int* termination,
                 double* tr_dual,
                 double* min_rayleigh,
                 double* max_rayleigh)
{
  if (!step_out)
  {
    return set_error(B200_ERR_ARG, "bad argument");
  }
  return cg_solve_impl(C, n, nnz_g, g_idx, g_val, trust_radius, rel_tol, max_iter, step_out, nullptr, iterations, termination, tr_dual, min_rayleigh, max_rayleigh);
}

int
b200_cg_solve_sparse(b200_cg* C,
                     int n,
                     int nnz_g,
                     const int* g_idx,
                     const double* g_val,
                     double trust_radius,
                     double rel_tol,
                     int max_iter,
                     double zero_eps,
                     int* step_idx,
                     double* step_val,
                     int* step_nnz,
                     int* iterations,
                     int* termination,
                     double* tr_dual,
                     double* min_rayleigh,
                     double* max_rayleigh)
{
  const SparseStep sp{zero_eps, step_idx, step_val, step_nnz};
  return cg_solve_impl(C, n, nnz_g, g_idx, g_val, trust_radius, rel_tol, max_iter, nullptr, &sp, iterations, termination, tr_dual, min_rayleigh, max_rayleigh);
}

int
b200_cg_free(b200_cg** handle)
{
  if (!handle || !*handle)
  {
    return B200_OK;
  }
  b200_cg* C = *handle;
  if (C->stream)
  {
    cudaStreamSynchronize(C->stream);
  }
  if (C->hess)
  {
    b200_mat_set_stream(C->hess, nullptr);
  }
  delete C;
  *handle = nullptr;
  return B200_OK;
}

} // extern "C"
