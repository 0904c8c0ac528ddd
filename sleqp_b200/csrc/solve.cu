// solve.cu -- the repeated solves K z = b behind SleqpFactCallbacks.solve (fact_types.h:12;
// vendor calls replaced: umfpack_di_solve fact_umfpack.c:220, cholmod_l_solve fact_cholmod.c:184).
//
//   K = [D A^T; A G]:  t = b_E / d,  b_R' = b_R - A t,  S y = b_R',  z_R = y,  z_E = (b_E - A^T y) / d
// The reduced system is solved with two dataflow sweeps (one kernel launch each) over the inverse panels (selective
// inversion): every supernode step is a matrix-vector product, cut into warp tasks that synchronise through
// per-supernode counters.
// Iterative refinement runs against the unperturbed K (SURVEY.md hard part 1).
#include "numeric.cuh"

#include <algorithm>
#include <mutex>
#include <string>
#include <cstdlib>

namespace b200
{

// ---- E-block elimination and back-substitution -------------------------------------------------
// b_R' = b_R - A D^-1 b_E in the permuted labels of the reduced system. Asval = A entries divided by the pivot of
// their column, Ak = K index of that column's variable (both resolved once: plan.hpp); four entries in flight.
__global__ void
k_pre(int m,
      const int* __restrict__ k_of_r,
      const int* __restrict__ pinv,
      const int* __restrict__ Aptr,
      const int* __restrict__ Ak,
      const double* __restrict__ Asval,
      const double* __restrict__ rhs,
      double* __restrict__ bR,
      int nflow,
      double* __restrict__ yf,
      double* __restrict__ x,
      int* __restrict__ flow)
{
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  // what the dataflow sweeps accumulate into starts at zero (counters and tickets: nflow ints)
  if (r < nflow)
  {
    flow[r] = 0;
  }
  if (r >= m)
  {
    return;
  }
  yf[r] = 0.0;
  x[r]  = 0.0;
  double acc  = rhs[k_of_r[r]];
  const int e = Aptr[r + 1];
  for (int q = Aptr[r]; q < e; q += 4)
  {
    int kk[4];
    double a[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      kk[u] = q + u < e ? Ak[q + u] : 0;
      a[u]  = q + u < e ? Asval[q + u] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      v[u] = q + u < e ? rhs[kk[u]] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      acc -= a[u] * v[u];
    }
  }
  bR[pinv[r]] = acc;
}

// z_R = y (original labels) and z_E = (b_E - A^T y) / d in one launch: thread i < m handles reduced row i, the others
// one eliminated variable each
__global__ void
k_post(int m,
       int nE,
       const int* __restrict__ k_of_r,
       const int* __restrict__ k_of_e,
       const int* __restrict__ pinv,
       const int* __restrict__ Aptr,
       const int* __restrict__ Ap, // permuted reduced row of every CSC entry
       const double* __restrict__ Aval,
       const double* __restrict__ dE,
       const double* __restrict__ rhs,
       const double* __restrict__ y,
       double* __restrict__ z)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m)
  {
    z[k_of_r[i]] = y[pinv[i]];
    return;
  }
  const int e = i - m;
  if (e >= nE)
  {
    return;
  }
  const int k   = k_of_e[e];
  const int end = Aptr[e + 1];
  double acc    = rhs[k];
  const double d = dE[e];
  for (int q = Aptr[e]; q < end; q += 4)
  {
    int pp[4];
    double a[4], v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      pp[u] = q + u < end ? Ap[q + u] : 0;
      a[u]  = q + u < end ? Aval[q + u] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      v[u] = q + u < end ? y[pp[u]] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      acc -= a[u] * v[u];
    }
  }
  z[k] = acc / d;
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

// a release is all the producer side of the counter hand-over needs: __threadfence() is fence.sc (MEMBAR.SC), and
// fence.acq_rel also invalidates the SM's L1 (CCTL.IVALL) for its acquire half
__device__ __forceinline__ void
fence_release()
{
  asm volatile("fence.release.gpu;" ::: "memory");
}

// All lanes poll the same word (one request per poll). While many producers are still missing the consumer is far
// from ready: it polls rarely and with relaxed loads (an acquire load invalidates the SM's whole L1, CCTL.IVALL, and
// thousands of warps hold tasks of the narrow top levels for most of the sweep: tight polls of all of them on the
// counter of a big supernode keep its L2 slice so busy that the producers' signals queue behind them). Once only
// FLOW_NEAR signals are missing: tight acquire polls, and the poll that succeeds is the acquire the vector loads
// need. (Back-to-back relaxed polls in that phase measured slightly slower.)
constexpr int FLOW_NEAR = 64;
__device__ __forceinline__ void
wait_counter(const int* cnt, int need, unsigned far_sleep, int near, unsigned per_signal)
{
  int c = ld_acquire(cnt);
  while (c < need)
  {
    if (need - c > near)
    {
      // the more signals are missing the longer the nap (4 ns per missing signal, far_sleep .. 8 far_sleep)
      __nanosleep(min(8u * far_sleep, max(far_sleep, per_signal * (unsigned)(need - c))));
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
static_assert(FLOW_SHARDS * FLOW_TICKET_PITCH == (FLOW_THREADS / 32) * 32, "sst_ticket_offset (numeric.cuh) assumes this layout");

struct FlowSched
{
  int ntasks;
  int far_sleep; // shortest nap of a consumer that is far from ready (ns)
  int near;      // tight polls once at most this many signals are missing
  int per_signal; // nap per missing signal (ns)
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

constexpr int FLOW_DEPTH     = 16; // panel entries in flight per lane = one batch of loads
constexpr int FLOW_MAX_DEPTH = FLOW_DEEP; // depth of a task: one batch in the narrow levels, two in the wide ones, a
                                          // rolling window over FLOW_DEEP entries in the bandwidth-bound ones (symbolic.cpp)

// A task of the bandwidth-bound levels (depth > 32, large 3D fronts): same contract as flow_task below, but the panel
// is streamed with a rolling window -- the load of entry u of the next batch is issued as soon as entry u of this
// batch has been consumed, so FLOW_DEPTH loads per lane stay in flight over the whole depth -- and the fixed cost of a
// task (record, vector, publication, ticket) is amortised over four times as much data. Not inlined: its register
// needs must not disturb the allocation of the latency-bound path.
template <bool FWD>
__device__ __noinline__ void
flow_task_deep(const SweepTask& T,
               int lane,
               const int* __restrict__ Ridx,
               const double* __restrict__ M,
               const double* __restrict__ Dinv,
               double* __restrict__ yacc,
               double* __restrict__ yf,
               double* __restrict__ x,
               const int* __restrict__ cnt,
               double* vsh,
               unsigned far_sleep,
               int near,
               unsigned per_signal)
{
  const int k = T.k, h = T.h;
  const int ld    = FWD ? h : k;
  const int o     = (FWD ? T.i0 : T.j0) + lane;
  const bool ov   = o < (FWD ? T.i1 : T.j1);
  const int d0    = FWD ? T.j0 : T.i0;
  const int nd    = (FWD ? T.j1 : T.i1) - d0; // > 32
  const double* P = M + T.Lptr + (long long)d0 * ld + (ov ? o : (FWD ? T.i0 : T.j0));
  double pre[FLOW_DEPTH];
#pragma unroll
  for (int u = 0; u < FLOW_DEPTH; ++u)
  {
    pre[u] = ov ? __ldcs(P + (long long)u * ld) : 0.0;
  }
  double* dst  = nullptr;
  double scale = 1.0;
  if (ov)
  {
    if (FWD)
    {
      if (o < k)
      {
        dst   = yf + T.first + o;
        scale = Dinv[T.first + o];
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
  // backward: the row indices behind this lane's share of the vector do not depend on the producers
  int ri[FLOW_DEEP / 32];
  if (!FWD)
  {
#pragma unroll
    for (int q = 0; q < FLOW_DEEP / 32; ++q)
    {
      const int d = d0 + lane + 32 * q;
      ri[q]       = (lane + 32 * q < nd && d >= k) ? Ridx[T.Rptr + d - k] : -1;
    }
  }
  if (T.wait_idx >= 0)
  {
    wait_counter(cnt + T.wait_idx, T.need, far_sleep, near, per_signal);
  }
  __syncwarp(); // the previous task's reads of vsh are done
#pragma unroll
  for (int q = 0; q < FLOW_DEEP / 32; ++q)
  {
    const int dd = lane + 32 * q;
    double v     = 0.0;
    if (dd < nd)
    {
      if (FWD)
      {
        v = __ldcg(yacc + T.first + d0 + dd);
      }
      else
      {
        v = ri[q] < 0 ? yf[T.first + d0 + dd] : __ldcg(x + ri[q]);
      }
    }
    vsh[dd] = v;
  }
  __syncwarp();
  double acc0 = 0.0, acc1 = 0.0;
  const double* vb = vsh;
  int rem          = nd;
  int ldv          = ld;
  do
  {
    rem -= FLOW_DEPTH;
    P += (long long)FLOW_DEPTH * ld;
    asm volatile("" : "+r"(ldv)); // one multiply-add per address instead of a hoisted table of sixteen 64-bit offsets
#pragma unroll
    for (int u = 0; u < FLOW_DEPTH; u += 2)
    {
      acc0 += pre[u] * vb[u];
      pre[u] = (ov && u < rem) ? __ldcs(P + (long long)u * ldv) : 0.0;
      acc1 += pre[u + 1] * vb[u + 1];
      pre[u + 1] = (ov && u + 1 < rem) ? __ldcs(P + (long long)(u + 1) * ldv) : 0.0;
    }
    vb += FLOW_DEPTH;
  } while (rem > 0);
  if (ov)
  {
    atomicAdd(dst, (acc0 + acc1) * scale);
  }
}

// One task up to (not including) the publication of its completion.
//   FWD: lanes are rows [i0, i1), depth is columns [j0, j1), panel element (r, j) at Mt[j * h + r].
//   BWD: lanes are columns [j0, j1), depth is rows [i0, i1), panel element (i, j) at Mr[i * k + j].
// vsh: FLOW_MAX_DEPTH doubles of shared memory private to the warp (broadcast of the vector).
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
          int near,
          unsigned per_signal,
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
    wait_counter(cnt + T.wait_idx, T.need, far_sleep, near, per_signal);
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
  vsh[lane] = vsrc ? (vplain ? *vsrc : __ldcg(vsrc)) : 0.0;
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
  if (nd > FLOW_DEPTH) // second batch of a deep task (wide levels only)
  {
#pragma unroll
    for (int u = 0; u < FLOW_DEPTH; ++u)
    {
      pre[u] = (ov && FLOW_DEPTH + u < nd) ? __ldcs(P + (long long)(FLOW_DEPTH + u) * ld) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < FLOW_DEPTH; u += 2)
    {
      acc0 += pre[u] * vsh[FLOW_DEPTH + u];
      acc1 += pre[u + 1] * vsh[FLOW_DEPTH + u + 1];
    }
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
  __shared__ double vsh_all[FLOW_THREADS / 32 * FLOW_MAX_DEPTH];
  __shared__ __align__(16) SweepTask slot_all[FLOW_THREADS / 32]; // the warp's task record, fetched with cp.async
  const int lane  = threadIdx.x & 31;
  double* vsh     = vsh_all + (threadIdx.x >> 5) * FLOW_MAX_DEPTH;
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
    __syncwarp(); // the record is in registers; the slot is overwritten by the next fetch of this warp only
    if (TRACE)
    {
      ph.fetch += clk_after((double)T.h) - cf;
    }
    if ((FWD ? T.j1 - T.j0 : T.i1 - T.i0) > 32) // bandwidth-bound level; reads the record from the slot (not traced)
    {
      flow_task_deep<FWD>(*slot, lane, Ridx, M, Dinv, yacc, yf, x, cnt, vsh, (unsigned)sched.far_sleep, sched.near, (unsigned)sched.per_signal);
    }
    else
    {
      flow_task<FWD, TRACE>(T, lane, Ridx, M, Dinv, yacc, yf, x, cnt, vsh, trace, (unsigned)sched.far_sleep, sched.near, (unsigned)sched.per_signal, ph);
    }
    if (TRACE)
    {
      cf = clk_after(0.0);
    }
    // publication. In the wide levels the draw of the next ticket is issued first, so that its round trip overlaps
    // with the fence (nothing can block between the draw and the signal). In the narrow levels consumers are
    // waiting for exactly this signal and the fence must not wait for the ticket atomic: signal first.
    const bool open      = tried < FLOW_SHARDS;
    const bool critical  = slot->pad1 != 0; // from the slot, not from the register copy: nothing of T stays live across
    const int signal_idx = slot->signal_idx; // the task (no spills around the call of the deep variant)
    int issued           = (open && !critical) ? flow_take_issue(ticket, lane, shard) : 0;
    if (signal_idx >= 0)
    {
      fence_release(); // every lane: its own atomics are visible device-wide before the counter moves
      __syncwarp();
      if (lane == 0)
      {
        atomicAdd(cnt + signal_idx, 1);
      }
    }
    if (open && critical)
    {
      issued = flow_take_issue(ticket, lane, shard);
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

// ---- small utilities -----------------------------------------------------------------------------------
__global__ void
k_scatter(int nnz, const int* __restrict__ idx, int first, const double* __restrict__ val, double* __restrict__ out, int offset)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz)
  {
    out[(idx ? idx[q] : first + q) + offset] = val[q];
  }
}

// values of tril(K) from the values of the constraint Jacobian (b200_fact_set_kkt): src < 0 = the constant 1
__global__ void
k_gather_kkt(int nnz, const int* __restrict__ src, const double* __restrict__ jval, double* __restrict__ val)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz)
  {
    const int sidx = src[q];
    val[q]         = sidx < 0 ? 1.0 : jval[sidx];
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

// ---- sparsification of a solution slice on the device (what sleqp_vec_set_from_raw does on the host, vec.c:72-104):
// keeps the entries with |v| > eps in ascending order. Three small kernels: per-chunk counts, scan of the counts,
// ordered write (ballot ranks inside a warp, shared-memory scan over the warps).
constexpr int CP_THREADS = 256, CP_PER = 4, CP_CHUNK = CP_THREADS * CP_PER;

__global__ void __launch_bounds__(CP_THREADS)
k_compact_count(int n, const double* __restrict__ x, double eps, int* __restrict__ chunk_cnt)
{
  __shared__ int wsum[CP_THREADS / 32];
  const int base = blockIdx.x * CP_CHUNK;
  int c          = 0;
#pragma unroll
  for (int r = 0; r < CP_PER; ++r)
  {
    const int i = base + r * CP_THREADS + threadIdx.x;
    c += i < n && fabs(x[i]) > eps;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0)
  {
    wsum[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    int t = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w)
    {
      t += wsum[w];
    }
    chunk_cnt[blockIdx.x] = t;
  }
}

// exclusive scan of the chunk counts in place (one CTA); chunk_cnt[nchunks] = total
__global__ void __launch_bounds__(1024)
k_compact_scan(int nchunks, int* __restrict__ chunk_cnt)
{
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0)
  {
    carry = 0;
  }
  __syncthreads();
  for (int b0 = 0; b0 < nchunks; b0 += 1024)
  {
    const int i = b0 + threadIdx.x;
    const int v = i < nchunks ? chunk_cnt[i] : 0;
    int incl    = v;
    for (int o = 1; o < 32; o <<= 1)
    {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o)
      {
        incl += t;
      }
    }
    if ((threadIdx.x & 31) == 31)
    {
      wsum[threadIdx.x >> 5] = incl;
    }
    __syncthreads();
    if (threadIdx.x < 32)
    {
      int w = wsum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1)
      {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o)
        {
          w += t;
        }
      }
      wsum[threadIdx.x] = w; // inclusive over the warps
    }
    __syncthreads();
    const int before = carry + (threadIdx.x >= 32 ? wsum[(threadIdx.x >> 5) - 1] : 0);
    if (i < nchunks)
    {
      chunk_cnt[i] = before + incl - v;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      carry += wsum[31];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0)
  {
    chunk_cnt[nchunks] = carry;
  }
}

__global__ void __launch_bounds__(CP_THREADS)
k_compact_write(int n, int nchunks, const double* __restrict__ x, double eps, const int* __restrict__ chunk_off, int* __restrict__ idx_out, double* __restrict__ val_out)
{
  __shared__ int wcnt[CP_PER][CP_THREADS / 32];
  const int base  = blockIdx.x * CP_CHUNK;
  const int total = chunk_off[nchunks];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[CP_PER];
  unsigned bal[CP_PER];
#pragma unroll
  for (int r = 0; r < CP_PER; ++r)
  {
    const int i = base + r * CP_THREADS + threadIdx.x;
    if (i < n && i >= total) // the slots behind the kept entries: the arrays are copied to the host at full length
    {
      idx_out[i] = 0;
      val_out[i] = 0.0;
    }
    v[r]        = i < n ? x[i] : 0.0;
    bal[r]      = __ballot_sync(0xffffffffu, i < n && fabs(v[r]) > eps);
    if (lane == 0)
    {
      wcnt[r][warp] = __popc(bal[r]);
    }
  }
  __syncthreads();
  int off = chunk_off[blockIdx.x];
#pragma unroll
  for (int r = 0; r < CP_PER; ++r)
  {
    int before = 0;
    for (int w = 0; w < CP_THREADS / 32; ++w)
    {
      before += w < warp ? wcnt[r][w] : 0;
    }
    if (bal[r] >> lane & 1u)
    {
      const int pos = off + before + __popc(bal[r] & ((1u << lane) - 1u));
      idx_out[pos]  = r * CP_THREADS + threadIdx.x + base;
      val_out[pos]  = v[r];
    }
    for (int w = 0; w < CP_THREADS / 32; ++w)
    {
      off += wcnt[r][w];
    }
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

namespace
{
int g_flow_sleep = 512;              // B200_FLOW_SLEEP: poll interval of far-away consumers (ns)
int g_flow_near = FLOW_NEAR, g_flow_per_signal = 4; // B200_FLOW_NEAR, B200_FLOW_PER_SIGNAL (experiments)
int g_flow_ctas  = FLOW_CTAS_PER_SM; // B200_FLOW_CTAS: fewer resident CTAs per SM (experiments)
} // namespace

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
    const int ns    = P.nsuper;
    const int nflow = sst_ticket_offset(ns) + 8; // counters, tickets of the dataflow sweeps, tickets of the sparse subtrees
    k_pre<<<nblocks(std::max(P.m, nflow), T), T, 0, stream>>>(P.m, dp.k_of_r.p, dp.pinv.p, dp.Acsr_ptr.p, dp.Acsr_k.p, nb.Acsr_sval, in, sb.y, nflow, sb.yf, sb.x,
                                                             sb.flow);
    lc.tick();
    mark(1);
    enqueue_sst_forward(dp, nb, sb, stream, lc); // sparse subtrees first (leaves of the supernodal tree)
    const int max_ctas      = sb.sms * g_flow_ctas;
    auto grid = [&](size_t ntasks) {
      // at least ~4 tasks per warp: small systems (batched multistart, many handles on one GPU) leave room for the
      // sweeps of other handles instead of parking spinning warps on every SM
      const long long ctas = ((long long)ntasks + 4 * FLOW_SHARDS - 1) / (4 * FLOW_SHARDS);
      return (unsigned)std::max<long long>(1, std::min<long long>(max_ctas, ctas));
    };
    const FlowSched fs{(int)P.ffl_tasks.size(), g_flow_sleep, g_flow_near, g_flow_per_signal};
    const FlowSched bs{(int)P.bfl_tasks.size(), g_flow_sleep, g_flow_near, g_flow_per_signal};
    int* const tickets_f = sb.flow + 2 * ns;
    int* const tickets_b = tickets_f + FLOW_SHARDS * FLOW_TICKET_PITCH;
    auto kf = sb.trace_fwd ? k_flow<true, true> : k_flow<true, false>;
    auto kb = sb.trace_bwd ? k_flow<false, true> : k_flow<false, false>;
    if (!P.ffl_tasks.empty()) // nothing dense left when sparse subtrees cover the whole tree
    {
      kf<<<sb.trace_fwd ? max_ctas : grid(P.ffl_tasks.size()), FLOW_THREADS, 0, stream>>>(dp.ffl_tasks.p, fs, dp.Ridx.p, nb.Mt, nb.Dinv, sb.y, sb.yf, sb.x, sb.flow, tickets_f,
                                                                  (const FlowTrace*)sb.trace_fwd);
      lc.tick();
    }
    mark(2);
    if (!P.bfl_tasks.empty())
    {
      kb<<<sb.trace_bwd ? max_ctas : grid(P.bfl_tasks.size()), FLOW_THREADS, 0, stream>>>(dp.bfl_tasks.p, bs, dp.Ridx.p, nb.Mr, nb.Dinv, sb.y, sb.yf, sb.x, sb.flow + ns, tickets_b,
                                                                   (const FlowTrace*)sb.trace_bwd);
      lc.tick();
    }
    enqueue_sst_backward(dp, nb, sb, stream, lc);
    mark(3);
  }
  if (P.m + P.nE > 0)
  {
    k_post<<<nblocks((long long)P.m + P.nE, T), T, 0, stream>>>(P.m, P.nE, dp.k_of_r.p, dp.k_of_e.p, dp.pinv.p, dp.Acsc_ptr.p, dp.Acsc_p.p, nb.Acsc_val, nb.dE, in, sb.x, out);
    lc.tick();
  }
  mark(4);
#ifdef B200_SST_TRACE_BUILD
  if (ev && std::getenv("B200_SST_TRACE"))
  {
    cudaStreamSynchronize(stream);
    dump_sst_trace();
  }
#endif
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
  // once per process, thread-safe (handles are created concurrently by independent solver threads): only the
  // experiment knobs live here; everything that depends on the device (SM count) is per handle (SolveBuffers::sms)
  static std::once_flag once;
  std::call_once(once, [] {
    if (const char* fc = std::getenv("B200_FLOW_CTAS"))
    {
      g_flow_ctas = std::min(FLOW_CTAS_PER_SM, std::max(1, std::atoi(fc)));
    }
    if (const char* fn = std::getenv("B200_FLOW_NEAR"))
    {
      g_flow_near = std::max(0, std::atoi(fn));
    }
    if (const char* fp = std::getenv("B200_FLOW_PER_SIGNAL"))
    {
      g_flow_per_signal = std::max(0, std::atoi(fp));
    }
    if (const char* fs = std::getenv("B200_FLOW_SLEEP"))
    {
      g_flow_sleep = std::max(32, std::atoi(fs));
    }
  });
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
  B200_CUDA(cudaMemsetAsync(nb.scal + 2, 0, 3 * sizeof(double), stream));
  const unsigned blocks = std::min<unsigned>(nblocks(P.N, 256), 1184);
  k_sumsq<<<blocks, 256, 0, stream>>>(P.N, sb.res, nb.scal + 2);
  lc.tick();
  k_sumsq<<<blocks, 256, 0, stream>>>(P.N, sb.rhs, nb.scal + 3);
  lc.tick();
  k_sumsq<<<blocks, 256, 0, stream>>>(P.N, sb.z, nb.scal + 4); // |x|^2: normwise backward error of the probe solve
  lc.tick();
  B200_CUDA(cudaGetLastError());
}

// scal[5], scal[6] = smallest and largest |v_i| (bit patterns of non-negative doubles)
void
enqueue_abs_range(const double* v, long long n, double* scal, cudaStream_t stream, LaunchCounter& lc)
{
  const unsigned long long init[2] = {0x7ff0000000000000ull, 0ull};
  B200_CUDA(cudaMemcpyAsync(scal + 5, init, sizeof(init), cudaMemcpyHostToDevice, stream));
  if (n > 0)
  {
    k_absrange<<<std::min<unsigned>(nblocks(n, 256), 1184), 256, 0, stream>>>((int)n, v, (unsigned long long*)(scal + 5), (unsigned long long*)(scal + 6));
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

void
enqueue_scatter_rhs(double* rhs, int n, int nnz, const int* d_idx, int first, const double* d_val, cudaStream_t stream, LaunchCounter& lc, int offset)
{
  B200_CUDA(cudaMemsetAsync(rhs, 0, sizeof(double) * (size_t)n, stream));
  if (nnz > 0)
  {
    k_scatter<<<nblocks(nnz, 256), 256, 0, stream>>>(nnz, d_idx, first, d_val, rhs, offset);
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

void
enqueue_gather_kkt(int nnz, const int* d_src, const double* d_jval, double* d_val, cudaStream_t stream)
{
  k_gather_kkt<<<nblocks(nnz, 256), 256, 0, stream>>>(nnz, d_src, d_jval, d_val);
  g_launches.fetch_add(1, std::memory_order_relaxed);
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
enqueue_compact(const double* x, int n, double eps, int* chunk_cnt, int* idx_out, double* val_out, cudaStream_t stream, LaunchCounter& lc)
{
  const int nchunks = (n + CP_CHUNK - 1) / CP_CHUNK;
  if (nchunks > 0)
  {
    k_compact_count<<<nchunks, CP_THREADS, 0, stream>>>(n, x, eps, chunk_cnt);
    lc.tick();
  }
  k_compact_scan<<<1, 1024, 0, stream>>>(nchunks, chunk_cnt);
  lc.tick();
  if (nchunks > 0)
  {
    k_compact_write<<<nchunks, CP_THREADS, 0, stream>>>(n, nchunks, x, eps, chunk_cnt, idx_out, val_out);
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

int
compact_chunks(int n)
{
  return (n + CP_CHUNK - 1) / CP_CHUNK;
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
