// fact.cu -- the b200_fact_* C-ABI: one handle per SleqpFact, owning a CUDA stream, the device
// copy of the cached plan, the factor, and the solve workspaces. Host glue: ../host/fact_b200.c.
#include "numeric.cuh"

#include <cuda.h> // CUtensorMap and the cuTensorMapEncodeTiled prototype (resolved at run time, no link against libcuda)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace b200
{
std::atomic<int64_t> g_launches{0};
}

using namespace b200;

namespace
{

struct Graph
{
  cudaGraphExec_t exec = nullptr;
  int64_t launches     = 0;
  void reset()
  {
    if (exec)
    {
      cudaGraphExecDestroy(exec);
      exec = nullptr;
    }
    launches = 0;
  }
};

constexpr int MAX_REFINE = 3;

} // namespace

struct b200_fact
{
  int device          = 0;
  int sms             = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_copy = nullptr;
  NumericOverlap overlap = {}; // side stream + events of the factorization graph (look-ahead)

  DevPlan dp;
  // factor
  DevBuf<double> val, L, Mt, Mr, tmp, U, D, Dinv, scratch, scal, dE, Acsc_val, Acsr_val, Acsr_sval, Gsym_val;
  DevBuf<int> nper;
  DevBuf<double> jval; // values of the constraint Jacobian (set_kkt: the KKT values are gathered from them on the device)
  // solve
  DevBuf<double> rhs, z, res, dz, bR, y, yf, x;
  DevBuf<int> rhs_idx, flow, cp_idx, cp_cnt; // cp_*: device-side sparsification of a solution slice
  DevBuf<double> cp_val;
  DevBuf<double> rhs_val;
  PinnedBuf<int> h_rhs_idx;
  PinnedBuf<double> h_rhs_val, h_sol, h_scal;
  PinnedBuf<int> h_nper;

  Graph g_numeric;
  Graph g_solve[MAX_REFINE + 1];

  bool factored = false, solved = false;
  bool timed_numeric = false, timed_solve = false;
  int refine          = 0;
  int n_perturbed     = 0;
  double probe_res    = 0;
  double rcond        = 0;
  bool symbolic_cached = false;
  double ms_symbolic  = 0;

  NumericBuffers nbuf() const
  {
    NumericBuffers nb;
    nb.val         = val.p;
    nb.L           = L.p;
    nb.Mt          = Mt.p;
    nb.tmp         = tmp.p;
    nb.Mr          = Mr.p;
    nb.U           = U.p;
    nb.D           = D.p;
    nb.Dinv        = Dinv.p;
    nb.scratch     = scratch.p;
    nb.scal        = scal.p;
    nb.n_perturbed = nper.p;
    nb.dE          = dE.p;
    nb.Acsc_val    = Acsc_val.p;
    nb.Acsr_val    = Acsr_val.p;
    nb.Acsr_sval   = Acsr_sval.p;
    nb.Gsym_val    = Gsym_val.p;
    return nb;
  }
  SolveBuffers sbuf() const
  {
    SolveBuffers sb;
    sb.rhs = rhs.p;
    sb.z   = z.p;
    sb.res = res.p;
    sb.dz  = dz.p;
    sb.bR  = bR.p;
    sb.y   = y.p;
    sb.yf  = yf.p;
    sb.x   = x.p;
    sb.flow = flow.p;
    sb.sms  = sms;
    return sb;
  }
  void drop_graphs()
  {
    g_numeric.reset();
    for (auto& g : g_solve)
    {
      g.reset();
    }
  }
};

namespace
{

int
pick_device(int device)
{
  if (device >= 0)
  {
    return device;
  }
  for (const char* name : {"B200_DEVICE", "LOCAL_RANK"})
  {
    const char* v = std::getenv(name);
    if (v && *v)
    {
      return std::atoi(v);
    }
  }
  return 0;
}

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
  catch (const std::bad_alloc&)
  {
    return set_error(B200_ERR_CUDA, "host allocation failed");
  }
  catch (const std::exception& e)
  {
    return set_error(B200_ERR_CUDA, e.what());
  }
}

// 2-D tensor map of a column-major panel (h rows x k columns of doubles, leading dimension ld, even): boxes of
// 16 x 16 elements in the 128-byte swizzle, rows / columns outside the panel read as zero.
void
encode_panel_map(PanelTensorMap* out, double* panel, int h, int k, int ld)
{
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
    {
      cudaGetLastError();
      fn = nullptr;
    }
    return (EncodeTiled)fn;
  }();
  if (!encode)
  {
    throw CudaError("cuTensorMapEncodeTiled is not available from this driver");
  }
  static_assert(sizeof(PanelTensorMap) == sizeof(CUtensorMap), "CUtensorMap is 128 bytes");
  const cuuint64_t dims[2]    = {(cuuint64_t)h, (cuuint64_t)k};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  const cuuint32_t box[2]     = {16, 16};
  const cuuint32_t estr[2]    = {1, 1};
  const CUresult rc = encode((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)panel, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS)
  {
    throw CudaError("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)rc) + " (h = " + std::to_string(h) + ", k = " + std::to_string(k) + ")");
  }
}

void
upload_plan(b200_fact* F, std::shared_ptr<const Plan> plan)
{
  const Plan& P  = *plan;
  DevPlan& dp    = F->dp;
  cudaStream_t s = F->stream;
  // the captured graphs have the old buffers baked in; until every upload and reservation below has succeeded the
  // handle has NO current plan (a failure midway must not leave a half-uploaded plan that the next set_matrix with
  // the same pattern would take for current)
  F->drop_graphs();
  dp.plan.reset();
  F->factored = F->solved = false;
  dp.Ridx.upload(P.Ridx, s);
  dp.rel.upload(P.rel, s);
  dp.child_idx.upload(P.child_idx, s);
  static_assert(sizeof(long long) == sizeof(i64), "i64");
  auto up64 = [&](DevBuf<long long>& b, const std::vector<i64>& v) {
    b.reserve(v.size());
    if (!v.empty())
    {
      B200_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(i64), cudaMemcpyHostToDevice, s));
    }
  };
  up64(dp.Sdest, P.Sdest);
  up64(dp.Sterm_ptr, P.Sterm_ptr);
  up64(dp.Sdiag, P.Sdiag);
  dp.Sgsrc.upload(P.Sgsrc, s);
  dp.Sterm_a.upload(P.Sterm_a, s);
  dp.Sterm_b.upload(P.Sterm_b, s);
  dp.Sterm_d.upload(P.Sterm_d, s);
  dp.zero_sn.upload(P.zero_sn, s);
  dp.ea_tasks.upload(P.ea_tasks, s);
  dp.pan_tasks.upload(P.pan_tasks, s);
  dp.upd_tasks.upload(P.upd_tasks, s);
  dp.lvl_sn.upload(P.lvl_sn, s);
  dp.inv_tasks.upload(P.inv_tasks, s);
  dp.tr_tasks.upload(P.tr_tasks, s);
  dp.ffl_tasks.upload(P.ffl_tasks, s);
  dp.bfl_tasks.upload(P.bfl_tasks, s);
  dp.k_of_e.upload(P.k_of_e, s);
  dp.k_of_r.upload(P.k_of_r, s);
  dp.pinv.upload(P.pinv, s);
  dp.perm.upload(P.perm, s);
  dp.dE_src.upload(P.dE_src, s);
  dp.Acsc_ptr.upload(P.Acsc_ptr, s);
  dp.Acsc_row.upload(P.Acsc_row, s);
  dp.Acsc_src.upload(P.Acsc_src, s);
  dp.Acsr_ptr.upload(P.Acsr_ptr, s);
  dp.Acsr_col.upload(P.Acsr_col, s);
  dp.Acsr_src.upload(P.Acsr_src, s);
  dp.Acsr_k.upload(P.Acsr_k, s);
  dp.Acsr_dsrc.upload(P.Acsr_dsrc, s);
  dp.Acsc_p.upload(P.Acsc_p, s);
  dp.Gsym_ptr.upload(P.Gsym_ptr, s);
  dp.Gsym_col.upload(P.Gsym_col, s);
  dp.Gsym_src.upload(P.Gsym_src, s);
  dp.Ksrc.upload(P.Ksrc, s);
  dp.sst.upload(P.sst, s);
  dp.sst_blob.upload(P.sst_blob, s);
  dp.sst_ea_src.upload(P.sst_ea_src, s);
  dp.sst_ea_dst.upload(P.sst_ea_dst, s);
  // the uploads read pageable host vectors owned by the (shared, immutable) plan: safe, but
  // finish them before anything else touches the stream
  B200_CUDA(cudaStreamSynchronize(s));

  const size_t N = (size_t)P.N, m = (size_t)P.m;
  F->L.reserve((size_t)P.Lptr[P.nsuper] + 8);
  F->Mt.reserve((size_t)P.Lptr[P.nsuper] + 8);
  F->tmp.reserve((size_t)P.Tptr[P.nsuper] + 8);
  F->Mr.reserve((size_t)P.Lptr[P.nsuper] + 8);
  F->U.reserve((size_t)P.Utotal + 8);
  F->D.reserve(m + 8);
  F->Dinv.reserve(m + 8);
  F->scratch.reserve((size_t)std::max(1, P.n_scratch_slots) * NB * NB);
  F->scal.reserve(8);
  F->nper.reserve(2);
  F->dE.reserve((size_t)P.nE + 8);
  F->Acsc_val.reserve(P.Acsc_src.size() + 8);
  F->Acsr_val.reserve(P.Acsr_src.size() + 8);
  F->Acsr_sval.reserve(P.Acsr_src.size() + 8);
  F->Gsym_val.reserve(P.Gsym_src.size() + 8);
  F->rhs.reserve(N + 8);
  F->z.reserve(N + 8);
  F->res.reserve(N + 8);
  F->dz.reserve(N + 8);
  F->bR.reserve(m + 8);
  F->y.reserve(m + 8);
  F->yf.reserve(m + 8);
  F->x.reserve(m + 8);
  F->flow.reserve(2 * (size_t)P.nsuper + 2 * (FLOW_THREADS / 32) * 32 + 8);
  F->h_sol.reserve(N + 8);
  F->h_scal.reserve(8);
  F->h_nper.reserve(2);
  std::vector<SnMeta> meta((size_t)P.nsuper);
  for (int T = 0; T < P.nsuper; ++T)
  {
    SnMeta& m     = meta[T];
    m.Lptr        = P.Lptr[T];
    m.Uoff        = P.Uoff[T];
    m.Rptr        = P.Rptr[T];
    m.Tptr        = P.Tptr[T];
    m.first       = P.sn_first[T];
    m.k           = P.sn_first[T + 1] - P.sn_first[T];
    m.r           = (int)(P.Rptr[T + 1] - P.Rptr[T]);
    m.parent      = P.sn_parent[T];
    m.child_begin = P.child_ptr[T];
    m.child_end   = P.child_ptr[T + 1];
    m.ld          = (int)panel_ld(m.k + m.r);
    m.tmap        = -1;
    m.pad2 = m.pad3 = m.pad4 = m.pad5 = 0;
  }
  // TMA tensor maps of the panels of large fronts (numeric.cu: tile_update_tma): L is allocated by now
  {
    const char* e       = std::getenv("B200_TMA_MIN_FRONT"); // tests lower it so that small fronts take the TMA path too
    const int min_front = e ? std::max(1, std::atoi(e)) : TMA_MIN_FRONT;
    std::vector<PanelTensorMap> maps;
    for (int T = 0; T < P.nsuper; ++T)
    {
      SnMeta& m = meta[T];
      if (m.k + m.r < min_front)
      {
        continue;
      }
      PanelTensorMap tm;
      encode_panel_map(&tm, F->L.p + m.Lptr, m.k + m.r, m.k, m.ld);
      m.tmap = (int)maps.size();
      maps.push_back(tm);
    }
    dp.tmaps.upload(maps, s);
    dp.sn.upload(meta, s);
    B200_CUDA(cudaStreamSynchronize(s)); // `maps` and `meta` are locals
  }
  dp.plan = plan;
}

template <typename Enqueue>
void
run_graph(b200_fact* F, Graph& g, Enqueue&& enqueue)
{
  if (!g.exec)
  {
    LaunchCounter lc;
    int64_t captured = 0;
    lc.captured      = &captured;
    cudaGraph_t graph = nullptr;
    B200_CUDA(cudaStreamBeginCapture(F->stream, cudaStreamCaptureModeThreadLocal));
    try
    {
      enqueue(lc);
    }
    catch (...)
    {
      cudaStreamEndCapture(F->stream, &graph);
      if (graph)
      {
        cudaGraphDestroy(graph);
      }
      throw;
    }
    B200_CUDA(cudaStreamEndCapture(F->stream, &graph));
    cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    B200_CUDA(e);
    g.launches = captured;
  }
  B200_CUDA(cudaGraphLaunch(g.exec, F->stream));
  g_launches.fetch_add(g.launches, std::memory_order_relaxed);
}

void
launch_solve(b200_fact* F, int refine)
{
  const NumericBuffers nb = F->nbuf();
  const SolveBuffers sb   = F->sbuf();
  run_graph(F, F->g_solve[refine], [&](LaunchCounter& lc) { enqueue_solve(F->dp, nb, sb, refine, F->stream, lc); });
}

// The part of a factorization that follows the plan lookup and the upload of the values (F->val): numeric graph,
// pivot range, probe solve that fixes the number of refinement steps (shared by set_matrix and set_kkt).
void
launch_numeric(b200_fact* F)
{
  const NumericBuffers nb = F->nbuf();
  run_graph(F, F->g_numeric, [&](LaunchCounter& lc) {
    enqueue_numeric(F->dp, nb, F->stream, lc, &F->overlap);
    enqueue_pivot_range(F->dp, nb, F->stream, lc);
  });
}

template <typename Lap>
int
factor_and_probe(b200_fact* F, Lap&& lap, bool numeric_enqueued = false)
{
  {
    const Plan& P = *F->dp.plan;
    if (P.N == 0)
    {
      F->factored = true;
      F->refine   = 0;
      F->rcond    = 1.0;
      return (int)B200_OK;
    }
    if (!numeric_enqueued)
    {
      launch_numeric(F);
      B200_CUDA(cudaEventRecord(F->ev_b, F->stream));
    }
    F->timed_numeric = true;
    lap("enqueue copy + numeric graph");
    // pivot range and perturbation count
    B200_CUDA(cudaMemcpyAsync(F->h_scal.p, F->scal.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, F->stream));
    B200_CUDA(cudaMemcpyAsync(F->h_nper.p, F->nper.p, sizeof(int), cudaMemcpyDeviceToHost, F->stream));
    B200_CUDA(cudaStreamSynchronize(F->stream));
    const double dmin = F->h_scal.p[2], dmax = F->h_scal.p[3];
    F->rcond       = (dmax > 0.0 && std::isfinite(dmax) && std::isfinite(dmin)) ? dmin / dmax : 0.0;
    F->n_perturbed = P.m > 0 ? F->h_nper.p[0] : 0;
    lap("numeric factorization (wait)");

    // probe solve: choose the number of refinement steps every solve of this factor performs
    const NumericBuffers nb = F->nbuf();
    const SolveBuffers sb   = F->sbuf();
    LaunchCounter eager;
    double best = INFINITY, prev = INFINITY, best_r = 0.0, best_x = 0.0, best_b = 0.0;
    int refine  = 0;
    for (int r = 0; r <= MAX_REFINE; ++r)
    {
      enqueue_probe_rhs(sb.rhs, P.N, F->stream, eager);
      launch_solve(F, r);
      enqueue_residual_norms(F->dp, nb, sb, F->stream, eager);
      B200_CUDA(cudaMemcpyAsync(F->h_scal.p, F->scal.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, F->stream));
      B200_CUDA(cudaStreamSynchronize(F->stream));
      const double rnorm = std::sqrt(F->h_scal.p[2]), bnorm = std::sqrt(F->h_scal.p[3]), xnorm = std::sqrt(F->h_scal.p[4]);
      const double rr = rnorm / bnorm;
      if (std::isfinite(rr) && rr < best)
      {
        best   = rr;
        refine = r;
        best_r = rnorm;
        best_x = xnorm;
        best_b = bnorm;
      }
      // good enough, broken, or refinement stopped paying
      if (!std::isfinite(rr) || rr <= 1e-13 || (r > 0 && rr > 0.25 * prev))
      {
        break;
      }
      prev = rr;
    }
    F->refine    = refine;
    F->probe_res = best;
    lap("probe solve(s) + residual");
    // Verdict. A residual of 1e-6 |b| or better (refinement is then chosen so that every solve meets 1e-10): fine.
    // Otherwise the system is accepted only if no pivot fell below the noise threshold (so the rows of the working set
    // are independent at working precision) and the computed solution is the exact solution of a system within 1e-13
    // of K relative to |K| -- nearly dependent rows, where |x| >> |b| and no backward-stable solver (Umfpack's LU
    // included) gets a small residual relative to |b|.
    double best_eta = INFINITY;
    if (!(best <= 1e-6) && F->n_perturbed == 0 && std::isfinite(best))
    {
      // normwise backward error with ||K||_2 bounded from below by its largest entry (i.e. the error from above)
      enqueue_abs_range(F->val.p, P.nnzK_input, F->scal.p, F->stream, eager);
      B200_CUDA(cudaMemcpyAsync(F->h_scal.p, F->scal.p, 7 * sizeof(double), cudaMemcpyDeviceToHost, F->stream));
      B200_CUDA(cudaStreamSynchronize(F->stream));
      best_eta = best_r / (F->h_scal.p[6] * best_x + best_b);
    }
    static const bool timing = std::getenv("B200_TIMING") != nullptr;
    if (timing)
    {
      std::fprintf(stderr, "[b200 probe] residual %.3e backward error %.3e perturbed %d refine %d\n", best, best_eta, F->n_perturbed, refine);
    }
    if (!(best <= 1e-6) && !(best_eta <= 1e-13))
    {
      return set_error(B200_ERR_SINGULAR,
                       "KKT matrix is numerically singular (probe residual " + std::to_string(best) + ", " + std::to_string(F->n_perturbed) +
                         " perturbed pivots): the working set rows are not linearly independent");
    }
    F->factored = true;
    return (int)B200_OK;
  }
}

} // namespace

extern "C" {

int
b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int
b200_host_pin(void* ptr, size_t bytes)
{
  if (!ptr || bytes == 0)
  {
    return set_error(B200_ERR_ARG, "null buffer");
  }
  const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess)
  {
    cudaGetLastError(); // not sticky, but the next cudaGetLastError() of an unrelated call would report it
    if (e == cudaErrorHostMemoryAlreadyRegistered && is_pinned_host(ptr, bytes))
    {
      return B200_OK; // another owner page-locked the same array: nothing to do
    }
    return set_error(B200_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  }
  return B200_OK;
}

int
b200_host_unpin(void* ptr)
{
  if (!ptr)
  {
    return set_error(B200_ERR_ARG, "null buffer");
  }
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    return set_error(B200_ERR_CUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
  }
  return B200_OK;
}

int64_t
b200_launch_count(void)
{
  return g_launches.load();
}

int
b200_fact_create(b200_fact** handle, int device)
{
  if (!handle)
  {
    return set_error(B200_ERR_ARG, "null handle pointer");
  }
  *handle = nullptr;
  return guarded([&]() {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
    {
      cudaGetLastError();
      return set_error(B200_ERR_CUDA, "no CUDA device available: the B200 backend has no CPU fallback");
    }
    const int dev = pick_device(device);
    if (dev >= count)
    {
      return set_error(B200_ERR_CUDA, "requested device " + std::to_string(dev) + " but only " + std::to_string(count) + " visible");
    }
    B200_CUDA(cudaSetDevice(dev));
    std::unique_ptr<b200_fact> F(new b200_fact());
    F->device = dev;
    B200_CUDA(cudaStreamCreateWithFlags(&F->stream, cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreate(&F->ev_a));
    B200_CUDA(cudaEventCreate(&F->ev_b));
    B200_CUDA(cudaEventCreate(&F->ev_c));
    B200_CUDA(cudaEventCreate(&F->ev_d));
    B200_CUDA(cudaEventCreateWithFlags(&F->ev_copy, cudaEventDisableTiming));
    B200_CUDA(cudaStreamCreateWithFlags(&F->overlap.side, cudaStreamNonBlocking));
    for (cudaEvent_t& e : F->overlap.panel_done)
    {
      B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    for (cudaEvent_t& e : F->overlap.rest_done)
    {
      B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    B200_CUDA(cudaDeviceGetAttribute(&F->sms, cudaDevAttrMultiProcessorCount, dev));
    configure_solve_kernels();
    configure_numeric_kernels(dev);
    configure_sst_kernels(dev);
    *handle = F.release();
    return (int)B200_OK;
  });
}

int
b200_fact_set_matrix(b200_fact* F, int n_rows, int n_cols, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only)
{
  if (!F)
  {
    return set_error(B200_ERR_ARG, "null handle");
  }
  if (n_rows != n_cols)
  {
    return set_error(B200_ERR_ARG, "matrix must be square (fact_umfpack.c:127 asserts the same)");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    F->factored = F->solved = false;
    // B200_TIMING=1: wall-clock of the host-side phases of this call on stderr
    static const bool timing = std::getenv("B200_TIMING") != nullptr;
    auto t_last              = std::chrono::steady_clock::now();
    auto lap                 = [&](const char* what) {
      if (timing)
      {
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[b200 set_matrix] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
      }
    };
    // the values go to the device while the host hashes the pattern and looks the plan up
    if (nnz < 0 || (nnz > 0 && !val))
    {
      return set_error(B200_ERR_ARG, "malformed CSC values");
    }
    if ((size_t)nnz + 8 > F->val.cap)
    {
      // DevBuf::reserve reallocates without keeping the contents and the captured graphs have val.p baked in: they
      // must not survive the old buffer (also when get_plan fails below and the previous plan stays current)
      B200_CUDA(cudaStreamSynchronize(F->stream));
      F->drop_graphs();
    }
    F->val.reserve((size_t)nnz + 8);
    B200_CUDA(cudaEventRecord(F->ev_a, F->stream));
    if (nnz > 0)
    {
      B200_CUDA(cudaMemcpyAsync(F->val.p, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, F->stream));
    }
    std::shared_ptr<const Plan> plan;
    bool cached = false;
    int rc      = get_plan(n_rows, nnz, colptr, rowidx, val, lower_only, plan, cached);
    lap("pattern hash + plan lookup");
    if (rc != B200_OK)
    {
      // `val` is borrowed for the call only: the copy above must not be in flight when we return
      cudaStreamSynchronize(F->stream);
      return rc;
    }
    F->symbolic_cached = cached;
    F->ms_symbolic     = cached ? 0.0 : plan->ms_symbolic;
    if (F->dp.plan != plan)
    {
      upload_plan(F, plan);
    }
    return factor_and_probe(F, lap);
  });
}

int
b200_fact_set_kkt(b200_fact* F,
                  int num_vars,
                  int num_cons,
                  int nnz_jac,
                  const int* jac_cols,
                  const int* jac_rows,
                  const double* jac_data,
                  const int* var_index,
                  const int* cons_index,
                  int working_set_size)
{
  if (!F)
  {
    return set_error(B200_ERR_ARG, "null handle");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    const bool was_factored = F->factored;
    F->factored = F->solved = false;
    static const bool timing = std::getenv("B200_TIMING") != nullptr;
    auto t_last              = std::chrono::steady_clock::now();
    auto lap                 = [&](const char* what) {
      if (timing)
      {
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[b200 set_kkt] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
        t_last = now;
      }
    };
    if (nnz_jac < 0 || (nnz_jac > 0 && !jac_data))
    {
      return set_error(B200_ERR_ARG, "malformed Jacobian values");
    }
    // the Jacobian values go to the device while the host hashes the pattern + working set and looks the plan up;
    // jval is not part of any captured graph (the gather below is launched eagerly)
    F->jval.reserve((size_t)nnz_jac + 8);
    B200_CUDA(cudaEventRecord(F->ev_a, F->stream));
    if (nnz_jac > 0)
    {
      B200_CUDA(cudaMemcpyAsync(F->jval.p, jac_data, sizeof(double) * (size_t)nnz_jac, cudaMemcpyHostToDevice, F->stream));
    }
    // Speculation: iterate after iterate the working set and the Jacobian pattern usually stay what they were. The
    // numeric factorization under the handle's current plan is enqueued right behind the copy, and the host hashes
    // the key (0.7 ms at config 3) while the device works; a different plan simply redoes the work (the values are on
    // the device already, the gather below reads them through the new plan's map).
    bool speculated = false;
    if (was_factored && F->dp.plan && F->dp.plan->kkt_keyed && F->dp.plan->key_aux == nnz_jac && F->dp.plan->N == num_vars + working_set_size && F->dp.plan->N > 0)
    {
      if (F->dp.plan->nnzK_input > 0)
      {
        enqueue_gather_kkt((int)F->dp.plan->nnzK_input, F->dp.Ksrc.p, F->jval.p, F->val.p, F->stream);
      }
      launch_numeric(F);
      B200_CUDA(cudaEventRecord(F->ev_b, F->stream));
      speculated = true;
    }
    std::shared_ptr<const Plan> plan;
    bool cached = false;
    int rc      = get_plan_kkt(num_vars, num_cons, nnz_jac, jac_cols, jac_rows, jac_data, var_index, cons_index, working_set_size, plan, cached);
    lap("key hash + plan lookup");
    if (speculated && rc == B200_OK && plan == F->dp.plan)
    {
      F->symbolic_cached = cached;
      F->ms_symbolic     = 0.0;
      return factor_and_probe(F, lap, /*numeric_enqueued=*/true);
    }
    if (rc != B200_OK)
    {
      cudaStreamSynchronize(F->stream); // `jac_data` is borrowed for the call only
      return rc;
    }
    F->symbolic_cached = cached;
    F->ms_symbolic     = cached ? 0.0 : plan->ms_symbolic;
    const Plan& P      = *plan;
    if ((size_t)P.nnzK_input + 8 > F->val.cap)
    {
      B200_CUDA(cudaStreamSynchronize(F->stream)); // the captured graphs have val.p baked in
      F->drop_graphs();
      F->dp.plan.reset();
      F->val.reserve((size_t)P.nnzK_input + 8);
    }
    if (F->dp.plan != plan)
    {
      upload_plan(F, plan);
    }
    if (P.nnzK_input > 0)
    {
      enqueue_gather_kkt((int)P.nnzK_input, F->dp.Ksrc.p, F->jval.p, F->val.p, F->stream);
    }
    return factor_and_probe(F, lap);
  });
}

int
b200_fact_refactor_device(b200_fact* F, const double* d_val)
{
  if (!F || !d_val)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "refactor needs a previous successful set_matrix with the same pattern");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    const Plan& P = *F->dp.plan;
    if (P.N == 0)
    {
      return (int)B200_OK;
    }
    B200_CUDA(cudaEventRecord(F->ev_a, F->stream));
    B200_CUDA(cudaMemcpyAsync(F->val.p, d_val, sizeof(double) * (size_t)P.nnzK_input, cudaMemcpyDeviceToDevice, F->stream));
    const NumericBuffers nb = F->nbuf();
    run_graph(F, F->g_numeric, [&](LaunchCounter& lc) {
      enqueue_numeric(F->dp, nb, F->stream, lc, &F->overlap);
      enqueue_pivot_range(F->dp, nb, F->stream, lc);
    });
    B200_CUDA(cudaEventRecord(F->ev_b, F->stream));
    F->timed_numeric = true;
    return (int)B200_OK;
  });
}

int
b200_fact_profile_numeric(b200_fact* F, double* ms_out)
{
  if (!F || !ms_out)
  {
    return set_error(B200_ERR_ARG, "bad argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "profile needs a factorization");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    struct Ctx
    {
      cudaStream_t stream;
      std::vector<cudaEvent_t> ev;
      std::vector<const char*> tag;
    } ctx;
    ctx.stream = F->stream;
    auto mark  = [](void* c, const char* tag) {
      Ctx* x = (Ctx*)c;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, x->stream);
      x->ev.push_back(e);
      x->tag.push_back(tag);
    };
    const NumericBuffers nb = F->nbuf();
    LaunchCounter lc;
    lc.trace     = mark;
    lc.trace_ctx = &ctx;
    B200_CUDA(cudaStreamSynchronize(F->stream));
    mark(&ctx, "start");
    enqueue_numeric(F->dp, nb, F->stream, lc);
    mark(&ctx, "rest");
    B200_CUDA(cudaStreamSynchronize(F->stream));
    const char* names[8] = {"assemble", "zero", "extend_add", "panel", "update", "inv_gemm", "transpose", "rest"};
    for (int i = 0; i < 8; ++i)
    {
      ms_out[i] = 0.0;
    }
    for (size_t i = 1; i < ctx.ev.size(); ++i)
    {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx.ev[i - 1], ctx.ev[i]);
      for (int n = 0; n < 8; ++n)
      {
        if (std::strcmp(ctx.tag[i], names[n]) == 0)
        {
          ms_out[n] += ms;
        }
      }
    }
    for (cudaEvent_t e : ctx.ev)
    {
      cudaEventDestroy(e);
    }
    return (int)B200_OK;
  });
}

int
b200_fact_profile_solve(b200_fact* F, int reps, double* ms_out)
{
  if (!F || !ms_out || reps <= 0)
  {
    return set_error(B200_ERR_ARG, "bad argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "profile needs a factorization");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    const NumericBuffers nb = F->nbuf();
    SolveBuffers sb         = F->sbuf();
    LaunchCounter eager;
    double acc[4] = {0, 0, 0, 0};
    // optional per-level timeline of the dataflow sweeps (stderr)
    const char* tr_env = std::getenv("B200_FLOW_TRACE");
    const bool tracing = tr_env && *tr_env && *tr_env != '0';
    const int nl       = F->dp.plan->nlevels;
    struct TraceRec
    {
      unsigned long long* rec;
      int nlevels;
      int pad;
    };
    int sms = 0;
    B200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, F->device));
    int ctas_per_sm = FLOW_CTAS_PER_SM;
    if (const char* fc = std::getenv("B200_FLOW_CTAS"))
    {
      ctas_per_sm = std::min(FLOW_CTAS_PER_SM, std::max(1, std::atoi(fc)));
    }
    const size_t nwarps = (size_t)sms * ctas_per_sm * (FLOW_THREADS / 32);
    const size_t per    = nwarps * (3 * (size_t)std::max(nl, 1) + 8); // level records, then 8 phase sums per warp
    DevBuf<unsigned long long> tr_data;
    DevBuf<TraceRec> tr_rec;
    std::vector<unsigned long long> tr_init;
    if (tracing && nl > 0)
    {
      tr_init.assign(2 * per, 0ull);
      for (int sweep = 0; sweep < 2; ++sweep)
      {
        for (size_t w = 0; w < nwarps; ++w)
        {
          std::fill(tr_init.begin() + sweep * per + w * 3 * nl, tr_init.begin() + sweep * per + (w * 3 + 2) * nl, ~0ull); // the two minima
        }
      }
      tr_data.reserve(tr_init.size());
      std::vector<TraceRec> rec = {{tr_data.p, nl, 0}, {tr_data.p + per, nl, 0}};
      tr_rec.upload(rec, F->stream);
      B200_CUDA(cudaStreamSynchronize(F->stream));
      sb.trace_fwd = tr_rec.p;
      sb.trace_bwd = tr_rec.p + 1;
    }
    cudaEvent_t ev[5];
    for (auto& e : ev)
    {
      B200_CUDA(cudaEventCreate(&e));
    }
    for (int it = 0; it < reps + 1; ++it)
    {
      if (tracing && nl > 0)
      {
        B200_CUDA(cudaMemcpyAsync(tr_data.p, tr_init.data(), sizeof(unsigned long long) * tr_init.size(), cudaMemcpyHostToDevice, F->stream));
      }
      enqueue_solve_phases(F->dp, nb, sb, F->stream, eager, ev);
      B200_CUDA(cudaStreamSynchronize(F->stream));
      if (tracing && nl > 0 && it == reps)
      {
        std::vector<unsigned long long> t(tr_init.size());
        B200_CUDA(cudaMemcpy(t.data(), tr_data.p, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost));
        for (int sweep = 0; sweep < 2; ++sweep)
        {
          std::vector<unsigned long long> c(3 * (size_t)nl);
          for (int q = 0; q < 3; ++q)
          {
            for (int l = 0; l < nl; ++l)
            {
              unsigned long long v = q == 2 ? 0ull : ~0ull;
              for (size_t w = 0; w < nwarps; ++w)
              {
                const unsigned long long e = t[sweep * per + (w * 3 + q) * nl + l];
                v                          = q == 2 ? std::max(v, e) : std::min(v, e);
              }
              c[(size_t)q * nl + l] = v;
            }
          }
          unsigned long long t0 = ~0ull;
          for (int l = 0; l < nl; ++l)
          {
            t0 = std::min(t0, c[l]);
          }
          std::fprintf(stderr, "[flow trace] %s sweep, us since the first claim: level  first-claim  first-ready  last-end\n", sweep ? "backward" : "forward");
          for (int l = 0; l < nl; ++l)
          {
            std::fprintf(stderr, "[flow trace]   %3d  %9.2f  %9.2f  %9.2f\n", l, (c[l] - t0) * 1e-3, (c[nl + l] - t0) * 1e-3, (c[2 * nl + l] - t0) * 1e-3);
          }
          double ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
          for (size_t w = 0; w < nwarps; ++w)
          {
            for (int q = 0; q < 8; ++q)
            {
              ph[q] += (double)t[sweep * per + nwarps * 3 * nl + w * 8 + q];
            }
          }
          const double nt = std::max(1.0, ph[5]);
          std::fprintf(stderr, "[flow trace]   %.0f tasks; cycles per task: record fetch %.0f  wait %.0f  vector (+ rest of the panel) %.0f  fma %.0f  publish + draw %.0f\n", ph[5],
                       ph[0] / nt, ph[1] / nt, ph[2] / nt, ph[3] / nt, ph[4] / nt);
        }
      }
      if (it == 0)
      {
        continue; // warm-up
      }
      for (int p = 0; p < 4; ++p)
      {
        float ms = 0.f;
        B200_CUDA(cudaEventElapsedTime(&ms, ev[p], ev[p + 1]));
        acc[p] += ms;
      }
    }
    for (auto& e : ev)
    {
      cudaEventDestroy(e);
    }
    for (int p = 0; p < 4; ++p)
    {
      ms_out[p] = acc[p] / reps;
    }
    return (int)B200_OK;
  });
}

int
b200_fact_solve(b200_fact* F, int nnz_rhs, const int* idx, const double* val, int dim)
{
  return b200_fact_solve_offset(F, nnz_rhs, idx, val, 0, dim);
}

int
b200_fact_solve_offset(b200_fact* F, int nnz_rhs, const int* idx, const double* val, int offset, int dim)
{
  if (!F)
  {
    return set_error(B200_ERR_ARG, "null handle");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "solve called before a successful set_matrix");
  }
  const Plan& P = *F->dp.plan;
  if (dim != P.N)
  {
    return set_error(B200_ERR_ARG, "rhs dimension " + std::to_string(dim) + " != matrix order " + std::to_string(P.N));
  }
  if (nnz_rhs < 0 || nnz_rhs > dim || (nnz_rhs > 0 && (!idx || !val)))
  {
    return set_error(B200_ERR_ARG, "malformed sparse rhs");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    if (P.N == 0)
    {
      F->solved = true;
      return (int)B200_OK;
    }
    // the previous solve may still be reading the staging buffers
    B200_CUDA(cudaStreamSynchronize(F->stream));
    LaunchCounter eager;
    bool copy_pending = false;
    B200_CUDA(cudaEventRecord(F->ev_c, F->stream));
    if (nnz_rhs > 0)
    {
      if (offset < 0 || idx[0] < 0 || (long long)idx[nnz_rhs - 1] + offset >= dim)
      {
        return set_error(B200_ERR_ARG, "rhs index out of range");
      }
      const bool contiguous = (idx[nnz_rhs - 1] - idx[0]) == nnz_rhs - 1; // ascending indices (pub_vec.h:13-14)
      F->rhs_val.reserve((size_t)nnz_rhs);
      // the right-hand side is borrowed for the duration of the call: pageable memory is staged through the
      // handle's pinned buffer, page-locked memory is read by the copy engine directly (and waited for below)
      bool direct = is_pinned_host(val, sizeof(double) * (size_t)nnz_rhs);
      if (direct)
      {
        B200_CUDA(cudaMemcpyAsync(F->rhs_val.p, val, sizeof(double) * (size_t)nnz_rhs, cudaMemcpyHostToDevice, F->stream));
      }
      else
      {
        F->h_rhs_val.reserve((size_t)nnz_rhs);
        std::memcpy(F->h_rhs_val.p, val, sizeof(double) * (size_t)nnz_rhs);
        B200_CUDA(cudaMemcpyAsync(F->rhs_val.p, F->h_rhs_val.p, sizeof(double) * (size_t)nnz_rhs, cudaMemcpyHostToDevice, F->stream));
      }
      const int* d_idx = nullptr;
      if (!contiguous)
      {
        F->h_rhs_idx.reserve((size_t)nnz_rhs);
        F->rhs_idx.reserve((size_t)nnz_rhs);
        std::memcpy(F->h_rhs_idx.p, idx, sizeof(int) * (size_t)nnz_rhs);
        B200_CUDA(cudaMemcpyAsync(F->rhs_idx.p, F->h_rhs_idx.p, sizeof(int) * (size_t)nnz_rhs, cudaMemcpyHostToDevice, F->stream));
        d_idx = F->rhs_idx.p;
      }
      if (direct)
      {
        B200_CUDA(cudaEventRecord(F->ev_copy, F->stream));
      }
      copy_pending = direct;
      enqueue_scatter_rhs(F->rhs.p, P.N, nnz_rhs, d_idx, idx[0], F->rhs_val.p, F->stream, eager, offset);
    }
    else
    {
      enqueue_scatter_rhs(F->rhs.p, P.N, 0, nullptr, 0, nullptr, F->stream, eager);
    }
    launch_solve(F, F->refine);
    B200_CUDA(cudaEventRecord(F->ev_d, F->stream));
    if (copy_pending)
    {
      B200_CUDA(cudaEventSynchronize(F->ev_copy)); // the caller may reuse its buffer as soon as we return
    }
    F->timed_solve = true;
    F->solved      = true;
    return (int)B200_OK;
  });
}

int
b200_fact_solution_ptr(b200_fact* F, int begin, int end, const double** out)
{
  if (!F || !out)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->solved)
  {
    return set_error(B200_ERR_STATE, "solution requested before solve");
  }
  const Plan& P = *F->dp.plan;
  if (begin < 0 || end < begin || end > P.N)
  {
    return set_error(B200_ERR_ARG, "solution slice out of range");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    if (end > begin)
    {
      B200_CUDA(cudaMemcpyAsync(F->h_sol.p, F->z.p + begin, sizeof(double) * (size_t)(end - begin), cudaMemcpyDeviceToHost, F->stream));
    }
    B200_CUDA(cudaStreamSynchronize(F->stream));
    *out = F->h_sol.p;
    return (int)B200_OK;
  });
}

int
b200_fact_solution(b200_fact* F, int begin, int end, double* out_dense)
{
  if (F && F->solved && out_dense && end > begin && begin >= 0 && F->dp.plan && end <= F->dp.plan->N && is_pinned_host(out_dense, sizeof(double) * (size_t)(end - begin)))
  {
    return guarded([&]() {
      B200_CUDA(cudaSetDevice(F->device));
      B200_CUDA(cudaMemcpyAsync(out_dense, F->z.p + begin, sizeof(double) * (size_t)(end - begin), cudaMemcpyDeviceToHost, F->stream));
      B200_CUDA(cudaStreamSynchronize(F->stream));
      return (int)B200_OK;
    });
  }
  const double* p = nullptr;
  int rc          = b200_fact_solution_ptr(F, begin, end, &p);
  if (rc != B200_OK)
  {
    return rc;
  }
  if (end > begin)
  {
    if (!out_dense)
    {
      return set_error(B200_ERR_ARG, "null output");
    }
    std::memcpy(out_dense, p, sizeof(double) * (size_t)(end - begin));
  }
  return B200_OK;
}

int
b200_fact_solution_sparse(b200_fact* F, int begin, int end, double zero_eps, int* idx_out, double* val_out, int* nnz_out)
{
  // Page-locked outputs: the slice is sparsified on the device (sleqp_vec_set_from_raw, vec.c:72-104) and both
  // arrays are DMA'd straight into the caller's buffers -- no host pass over the values. (The arrays are copied at
  // full length so that one synchronisation serves the count and the data.)
  if (F && F->solved && nnz_out && idx_out && val_out && end > begin && begin >= 0 && F->dp.plan && end <= F->dp.plan->N && is_pinned_host(val_out, sizeof(double) * (size_t)(end - begin)) && is_pinned_host(idx_out, sizeof(int) * (size_t)(end - begin)))
  {
    return guarded([&]() {
      B200_CUDA(cudaSetDevice(F->device));
      const int n       = end - begin;
      const int nchunks = compact_chunks(n);
      F->cp_idx.reserve((size_t)n);
      F->cp_val.reserve((size_t)n);
      F->cp_cnt.reserve((size_t)nchunks + 1);
      LaunchCounter eager;
      enqueue_compact(F->z.p + begin, n, zero_eps, F->cp_cnt.p, F->cp_idx.p, F->cp_val.p, F->stream, eager);
      B200_CUDA(cudaMemcpyAsync(F->h_nper.p + 1, F->cp_cnt.p + nchunks, sizeof(int), cudaMemcpyDeviceToHost, F->stream));
      B200_CUDA(cudaMemcpyAsync(val_out, F->cp_val.p, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, F->stream));
      B200_CUDA(cudaMemcpyAsync(idx_out, F->cp_idx.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, F->stream));
      B200_CUDA(cudaStreamSynchronize(F->stream));
      *nnz_out = F->h_nper.p[1];
      return (int)B200_OK;
    });
  }
  const double* p = nullptr;
  int rc          = b200_fact_solution_ptr(F, begin, end, &p);
  if (rc != B200_OK)
  {
    return rc;
  }
  if (!nnz_out || (end > begin && (!idx_out || !val_out)))
  {
    return set_error(B200_ERR_ARG, "null output");
  }
  // compaction in blocks of 64: solutions are mostly dense, so a block without dropped entries (the common
  // case) is a straight copy; only mixed blocks take the entry-by-entry path
  const int n = end - begin;
  int nnz     = 0;
  for (int i0 = 0; i0 < n; i0 += 64)
  {
    const int len = std::min(64, n - i0);
    int keep      = 0;
    for (int i = 0; i < len; ++i)
    {
      keep += std::fabs(p[i0 + i]) > zero_eps;
    }
    if (keep == len)
    {
      std::memcpy(val_out + nnz, p + i0, sizeof(double) * (size_t)len);
      for (int i = 0; i < len; ++i)
      {
        idx_out[nnz + i] = i0 + i;
      }
      nnz += len;
    }
    else if (keep > 0)
    {
      for (int i = 0; i < len; ++i)
      {
        const double v = p[i0 + i];
        idx_out[nnz]   = i0 + i;
        val_out[nnz]   = v;
        nnz += std::fabs(v) > zero_eps;
      }
    }
  }
  *nnz_out = nnz;
  return B200_OK;
}

int
b200_fact_solve_device(b200_fact* F, const double* d_rhs, double* d_sol)
{
  if (!F || !d_rhs || !d_sol)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "solve called before a successful set_matrix");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    const Plan& P = *F->dp.plan;
    if (P.N == 0)
    {
      return (int)B200_OK;
    }
    B200_CUDA(cudaEventRecord(F->ev_c, F->stream));
    if (d_rhs != F->rhs.p)
    {
      B200_CUDA(cudaMemcpyAsync(F->rhs.p, d_rhs, sizeof(double) * (size_t)P.N, cudaMemcpyDeviceToDevice, F->stream));
    }
    launch_solve(F, F->refine);
    if (d_sol != F->z.p)
    {
      B200_CUDA(cudaMemcpyAsync(d_sol, F->z.p, sizeof(double) * (size_t)P.N, cudaMemcpyDeviceToDevice, F->stream));
    }
    B200_CUDA(cudaEventRecord(F->ev_d, F->stream));
    F->timed_solve = true;
    F->solved      = true;
    return (int)B200_OK;
  });
}

int
b200_fact_device_buffers(b200_fact* F, double** rhs, double** solution)
{
  if (!F || !rhs || !solution)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "device buffers requested before a successful set_matrix");
  }
  *rhs      = F->rhs.p;
  *solution = F->z.p;
  return B200_OK;
}

int
b200_fact_rcond(b200_fact* F, double* rcond)
{
  if (!F || !rcond)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "condition requested before a successful set_matrix");
  }
  *rcond = F->rcond;
  return B200_OK;
}

int
b200_fact_stats(b200_fact* F, b200_stats* stats)
{
  if (!F || !stats)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->dp.plan)
  {
    return set_error(B200_ERR_STATE, "no matrix set");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    fill_stats_from_plan(*F->dp.plan, stats);
    stats->symbolic_cached = F->symbolic_cached;
    stats->ms_symbolic     = F->ms_symbolic;
    stats->n_perturbed     = F->n_perturbed;
    stats->refine_steps    = F->refine;
    stats->probe_residual  = F->probe_res;
    B200_CUDA(cudaStreamSynchronize(F->stream));
    float ms = 0.f;
    if (F->timed_numeric)
    {
      B200_CUDA(cudaEventElapsedTime(&ms, F->ev_a, F->ev_b));
      stats->ms_numeric = ms;
    }
    if (F->timed_solve)
    {
      B200_CUDA(cudaEventElapsedTime(&ms, F->ev_c, F->ev_d));
      stats->ms_solve = ms;
    }
    return (int)B200_OK;
  });
}

int
b200_fact_structure(b200_fact* F, int* perm, int* parent, int* colcount, int* n_super_total, int* super_first)
{
  if (!F || !F->dp.plan)
  {
    return set_error(B200_ERR_STATE, "no matrix set");
  }
  full_structure(*F->dp.plan, perm, parent, colcount, n_super_total, super_first);
  return B200_OK;
}

int
b200_fact_pivots(b200_fact* F, double* d_out)
{
  if (!F || !d_out)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  if (!F->factored)
  {
    return set_error(B200_ERR_STATE, "pivots requested before a successful set_matrix");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(F->device));
    const Plan& P = *F->dp.plan;
    if (P.nE > 0)
    {
      B200_CUDA(cudaMemcpyAsync(d_out, F->dE.p, sizeof(double) * (size_t)P.nE, cudaMemcpyDeviceToHost, F->stream));
    }
    if (P.m > 0)
    {
      B200_CUDA(cudaMemcpyAsync(d_out + P.nE, F->D.p, sizeof(double) * (size_t)P.m, cudaMemcpyDeviceToHost, F->stream));
    }
    B200_CUDA(cudaStreamSynchronize(F->stream));
    return (int)B200_OK;
  });
}

void*
b200_fact_stream(b200_fact* F)
{
  return F ? (void*)F->stream : nullptr;
}

int
b200_fact_device(b200_fact* F)
{
  return F ? F->device : -1;
}

int
b200_fact_free(b200_fact** handle)
{
  if (!handle || !*handle)
  {
    return B200_OK;
  }
  b200_fact* F = *handle;
  cudaSetDevice(F->device);
  if (F->stream)
  {
    cudaStreamSynchronize(F->stream);
  }
  F->drop_graphs();
  for (cudaEvent_t e : {F->ev_a, F->ev_b, F->ev_c, F->ev_d, F->ev_copy, F->overlap.panel_done[0], F->overlap.panel_done[1], F->overlap.rest_done[0], F->overlap.rest_done[1],
                        F->overlap.rest_done[2]})
  {
    if (e)
    {
      cudaEventDestroy(e);
    }
  }
  if (F->overlap.side)
  {
    cudaStreamDestroy(F->overlap.side);
  }
  if (F->stream)
  {
    cudaStreamDestroy(F->stream);
  }
  delete F;
  *handle = nullptr;
  return B200_OK;
}

} // extern "C"
