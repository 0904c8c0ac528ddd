// numeric.cuh -- interfaces between the C-ABI layer (fact.cu) and the kernel files.
#pragma once
#include "device.cuh"

namespace b200
{

// Raw device pointers of one factorization (owned by the handle).
struct NumericBuffers
{
  const double* val; // values of tril(K), same order as the caller's CSC
  double* L;         // supernodal panels [L11; L21]
  double* Mt;        // inverse panels [L11^-1; -L21 L11^-1] (what the solves read)
  double* Mr;        // row-major copy of the inverse panels (the backward sweep reads it)
  double* tmp;       // k x k scratch of the selective inversion
  double* U;         // update-matrix workspace
  double* D;         // pivots of S (new labels)
  double* Dinv;      // their reciprocals (the diagonal solve is a multiply in the forward sweep)
  double* scratch;   // diagonal-block scratch slots
  double* scal;      // [0] max|S_jj|, [1] pivot threshold, [2..3] reduction results
  int* n_perturbed;
  double* dE;        // pivots of the E block
  double* Acsc_val;
  double* Acsr_val;
  double* Acsr_sval; // A entries divided by the pivot of their column (k_pre)
  double* Gsym_val;
};

struct SolveBuffers
{
  double* rhs;  // N, original K indexing
  double* z;    // N, solution
  double* res;  // N, residual / correction right-hand side
  double* dz;   // N, correction
  double* bR;   // m, reduced right-hand side (new labels)
  double* y;    // m, right-hand side of the reduced system, accumulates the updates of the forward sweep
  double* yf;   // m, forward result (new labels)
  double* x;    // m, solution of the reduced system (new labels)
  const void* trace_fwd = nullptr; // device FlowTrace records (profile entry point with B200_FLOW_TRACE=1 only)
  const void* trace_bwd = nullptr;
  int sms = 148; // SM count of the handle's device (grid of the persistent sweep kernels)
  int* flow;    // dataflow sweeps: [0, ns) forward counters, [ns, 2 ns) backward counters, then the two ticket counters
};

// configuration of the kernels before any capture: the experiment knobs once per process, the opt-in to large dynamic
// shared memory once per device (`device` is current)
void configure_solve_kernels();
void configure_numeric_kernels(int device);

// Second stream + events of the look-ahead: the update tiles of a stage that do not touch the next panel's columns
// run next to the next panel step (numeric.cu).
struct NumericOverlap
{
  cudaStream_t side;
  cudaEvent_t panel_done[2]; // rings over the stages
  cudaEvent_t rest_done[3];
};

// sparse subtrees (sst.cu): factorization before the first stage of the dense schedule, forward sweep before the
// forward dataflow kernel (signals the parents' counters), backward sweep after the backward dataflow kernel
void configure_sst_kernels(int device);
void dump_sst_trace(); // -DB200_SST_TRACE_BUILD only
// the two ticket counters of the sparse-subtree sweeps inside SolveBuffers::flow (after the counters and tickets of
// the dataflow sweeps; zeroed by k_pre with the rest)
inline int
sst_ticket_offset(int nsuper)
{
  return 2 * nsuper + 2 * (FLOW_THREADS / 32) * 32;
}
void enqueue_sst_factor(const DevPlan& dp, const NumericBuffers& nb, cudaStream_t stream, LaunchCounter& lc);
void enqueue_sst_forward(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc);
void enqueue_sst_backward(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc);

// ov == nullptr: everything on `stream` in stage order (profiling entry point: per-class timings)
void enqueue_numeric(const DevPlan& dp, const NumericBuffers& nb, cudaStream_t stream, LaunchCounter& lc, const NumericOverlap* ov = nullptr);

// One solve K z = rhs with `refine` refinement steps, everything on `stream`.
void enqueue_solve(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, int refine, cudaStream_t stream, LaunchCounter& lc);

// one unrefined solve launched eagerly with events around its four phases:
// ev[0] pre ev[1] forward sweep ev[2] backward sweep ev[3] post ev[4]
void enqueue_solve_phases(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc, cudaEvent_t* ev);

// scal[2] = ||rhs - K z||_2^2, scal[3] = ||rhs||_2^2
void enqueue_residual_norms(const DevPlan& dp, const NumericBuffers& nb, const SolveBuffers& sb, cudaStream_t stream, LaunchCounter& lc);

// rhs[i] = 0 for all i, then rhs[idx[q]] = val[q]; idx == nullptr means the contiguous range
// [first, first + nnz).
void enqueue_scatter_rhs(double* rhs, int n, int nnz, const int* d_idx, int first, const double* d_val, cudaStream_t stream, LaunchCounter& lc, int offset = 0);

// val[q] = src[q] < 0 ? 1 : jval[src[q]]: the values of tril(K) gathered from the constraint Jacobian's (Plan::Ksrc)
void enqueue_gather_kkt(int nnz, const int* d_src, const double* d_jval, double* d_val, cudaStream_t stream);

// deterministic pseudo-random probe right-hand side
void enqueue_probe_rhs(double* rhs, int n, cudaStream_t stream, LaunchCounter& lc);

// min/max |d| over both pivot sets -> scal[2], scal[3]
void enqueue_pivot_range(const DevPlan& dp, const NumericBuffers& nb, cudaStream_t stream, LaunchCounter& lc);

// Sparsification of x[0, n) on the device (sleqp_vec_set_from_raw, vec.c:72-104): the entries with |x_i| > eps in
// ascending order -> (idx_out, val_out), their number -> chunk_cnt[compact_chunks(n)]. chunk_cnt: compact_chunks(n) + 1 ints.
void enqueue_abs_range(const double* v, long long n, double* scal, cudaStream_t stream, LaunchCounter& lc);
void enqueue_compact(const double* x, int n, double eps, int* chunk_cnt, int* idx_out, double* val_out, cudaStream_t stream, LaunchCounter& lc);
int compact_chunks(int n);

// out[i] = dE[i] for i < nE, D[i - nE] otherwise
void enqueue_copy_pivots(const DevPlan& dp, const NumericBuffers& nb, double* out, cudaStream_t stream, LaunchCounter& lc);

} // namespace b200
