// numeric.cu -- device numeric factorization: S = G - A D^-1 A^T assembled straight into the
// supernodal panels, a multifrontal LDL^T driven by the stage schedule of the plan, then the
// selective inversion that turns every panel into Minv = [L11^-1; -L21 L11^-1] for the solves.
//
// Replaces the vendor calls umfpack_di_numeric (fact_umfpack.c:154) / cholmod_l_factorize
// (fact_cholmod.c:137). Kernels (all FP64):
//   k_assemble      one thread per entry of tril(S): gathers its product terms      (HBM-bound)
//   k_extend_add    child update matrix -> parent front (relative indices)         (HBM-bound)
//   k_panel         NB-wide panel step: LDL^T + inverse of the diagonal block in shared memory, then
//                   L21 = F21 L11^-T D^-1 as a DMMA product with the inverted block       (latency/HBM)
//   k_update        64x64 output tiles of C -= L_i D L_j^T on the FP64 tensor cores
//                   (mma.sync m8n8k4.f64 = DMMA), operands staged through shared memory
//                   (tcgen05 has no f64 kind, so DMMA is the FP64 tensor path on sm_100a)
//   k_inv_gemm      64x64 DMMA tiles of the triangular products of the selective inversion
#include "numeric.cuh"

#include <algorithm>
#include <mutex>

namespace b200
{

// ---------------------------------------------------------------------------------------------
__global__ void
k_assemble(long long nnzS,
           const long long* __restrict__ Sdest,
           const int* __restrict__ Sgsrc,
           const long long* __restrict__ Sterm_ptr,
           const int* __restrict__ Sa,
           const int* __restrict__ Sb,
           const int* __restrict__ Sd,
           const double* __restrict__ val,
           double* __restrict__ L)
{
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nnzS)
  {
    return;
  }
  const int g = Sgsrc[q];
  double v    = g >= 0 ? val[g] : 0.0;
  for (long long t = Sterm_ptr[q]; t < Sterm_ptr[q + 1]; ++t)
  {
    v -= val[Sa[t]] * val[Sb[t]] / val[Sd[t]];
  }
  L[Sdest[q]] = v;
}

// gathers contiguous value arrays used by the solve: dE, A (both layouts), G
__global__ void
k_gather(int n, const int* __restrict__ src, const double* __restrict__ val, double* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out[i] = val[src[i]];
  }
}

__global__ void
k_gather_scaled(int n, const int* __restrict__ src, const int* __restrict__ dsrc, const double* __restrict__ val, double* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out[i] = val[src[i]] / val[dsrc[i]];
  }
}

// scal[0] = max_j |S_jj|, scal[1] = 64 eps * that  (static pivot threshold)
__global__ void
k_diagmax(int m, const long long* __restrict__ Sdiag, const double* __restrict__ L, unsigned long long* __restrict__ scal_bits)
{
  double v = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x)
  {
    v = fmax(v, fabs(L[Sdiag[j]]));
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  }
  if ((threadIdx.x & 31) == 0 && v > 0.0)
  {
    atomicMax(scal_bits, (unsigned long long)__double_as_longlong(v)); // non-negative doubles order like integers
  }
}

// Static pivot threshold relative to the largest |S_jj|. Negative definite S (everything SLEQP builds): 64 ulps -- a
// pivot below that is rounding noise, i.e. dependent working-set rows. Quasi-definite S (candidates kept in the
// reduced system, plan.hpp n_demoted): a working-set row without an eliminated neighbour has an exactly zero pivot
// when the ordering places it before its variables; it is replaced by -sqrt(eps) |S|_max (the usual static-pivoting
// size) and iterative refinement against the unperturbed K removes the perturbation.
__global__ void
k_set_tau(double* scal, double rel)
{
  scal[1] = rel * scal[0];
}

// ---------------------------------------------------------------------------------------------
__global__ void
k_zero_updates(const int* __restrict__ zero_sn, const SnMeta* __restrict__ sn, double* __restrict__ U)
{
  const SnMeta s      = sn[zero_sn[blockIdx.x]];
  const long long cnt = (long long)s.r * s.r;
  double* u           = U + s.Uoff;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.y * blockDim.x)
  {
    u[i] = 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Extend-add of EA_COLS columns of a child's update matrix into its parent's front: one warp per
// column, lanes stride the rows with 8 independent loads in flight. Children of one parent run
// concurrently in the same launch, hence the (native FP64) atomics.
__global__ void __launch_bounds__(32 * EA_COLS)
k_extend_add(const EaTask* __restrict__ tasks,
             const SnMeta* __restrict__ sn,
             const int* __restrict__ rel,
             double* __restrict__ L,
             double* __restrict__ U)
{
  const EaTask t   = tasks[blockIdx.x];
  const SnMeta c   = sn[t.child];
  const SnMeta p   = sn[c.parent];
  const int* rl    = rel + c.Rptr;
  const double* Uc = U + c.Uoff;
  const int hp     = p.ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = t.jb * EA_COLS + warp;
  if (j >= c.r)
  {
    return;
  }
  const int pj      = rl[j];
  double* dst       = pj < p.k ? L + p.Lptr + (long long)pj * hp : U + p.Uoff + (long long)(pj - p.k) * p.r - p.k;
  const double* src = Uc + (long long)j * c.r;
  int i             = j + lane;
  for (; i + 32 * 7 < c.r; i += 32 * 8)
  {
    int ri[8];
    double uv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      ri[u] = rl[i + 32 * u];
      uv[u] = src[i + 32 * u];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      atomicAdd(dst + ri[u], uv[u]);
    }
  }
  for (; i < c.r; i += 32)
  {
    atomicAdd(dst + rl[i], src[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// DMMA: D(8x8) += A(8x4, row) * B(4x8, col). Lane l holds a = A[l/4][l%4], b = B[l%4][l/4],
// c0/c1 = C[l/4][2*(l%4) + 0/1].
__device__ __forceinline__ void
dmma(double& c0, double& c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------
// Panel step t of supernode T, columns [c0, c0+w) of the front (w <= NB), in two kernels:
//
// k_panel (one CTA of 256 threads per RB rows of a panel step): right-looking LDL^T of the w x w diagonal block with
//   static pivoting and, in the same sweep, the inverse of its unit lower factor (the elimination steps applied to
//   the identity). Thread (lane = row i, warp = column class c mod 8) owns 4 entries of the 32 x 32 work per pivot;
//   one barrier per pivot; shared-memory read-modify-writes are batched (all loads, then all stores). Column j stays
//   unscaled (f_ij) inside the loop: l_ij d_j l_cj = f_ij f_cj / d_j. Then L21 = F21 * L11^-T D^-1 as a DMMA product
//   with the inverted block: A fragments straight from global memory (loaded before the factorization starts), B
//   fragments Wm[k][n] = L11^-1[n][k] / d_n from shared memory. The triangular solve is a tensor-core GEMM.
//
// History (ncu, profiles/): a single fused kernel with the factorization unrolled on one warp was instruction-fetch
// bound (42 us per CTA), a 128-thread rolled version issue-bound (27 us); a 256-thread factorization kernel followed
// by a separate DMMA kernel cost two launches per step; every CTA factoring the block itself removes one of them.
constexpr int LDP = 36; // leading dimension of the B-fragment operand: (4 k + n) mod 16 is conflict-free

// One panel step of one supernode: CTA rb owns RB rows below the NB x NB diagonal block (8 warps x 16 rows).
// EVERY CTA of the step factors the diagonal block itself (shared memory, 256 threads, ~32 barriers) instead of
// waiting for a separate kernel to publish it: the redundant arithmetic is free, the kernel boundary it replaces
// cost ~10 us per step on the critical path of the factorization. The diagonal block of L is left as assembled
// (nobody reads it afterwards: pivots, inverse block and L21 carry everything); CTA 0 publishes the pivots and
// the inverse of the unit lower factor.
constexpr int PANEL_THR = 256;

__global__ void __launch_bounds__(PANEL_THR, 2)
k_panel(const PanelTask* __restrict__ tasks,
        const SnMeta* __restrict__ sn,
        double* __restrict__ L,
        double* __restrict__ Mt,
        double* __restrict__ D,
        double* __restrict__ Dinv,
        const double* __restrict__ scal,
        int* __restrict__ n_perturbed)
{
  __shared__ double A[NB][NB + 1];    // becomes the unit lower factor (strict lower part, unscaled columns)
  __shared__ double Ainv[NB][NB + 1]; // becomes its inverse
  __shared__ double dsh[NB], dinv[NB];
  __shared__ double Wm[NB][LDP]; // Wm[k][n] = L11^-1[n][k] / d_n
  const PanelTask t = tasks[blockIdx.x];
  const SnMeta s    = sn[t.sn];
  const int h       = s.k + s.r;
  const int ld      = s.ld;
  const int c0      = t.t * NB;
  const int w       = min(NB, s.k - c0);
  double* P         = L + s.Lptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWD = PANEL_THR / 32, EPT = NB / NWD; // entries per thread and pivot
  constexpr int MI  = RB / NWD / 8;                   // 8-row DMMA fragments per warp
  static_assert(RB == NWD * 8 * MI, "RB rows are split evenly over the warps");

  // A fragments of this warp's rows straight from global memory (in flight during the factorization)
  const int r0 = c0 + w + t.rb * RB + warp * (8 * MI);
  double af[MI][8];
  if (r0 < h)
  {
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int k4 = 0; k4 < 8; ++k4)
      {
        const int row = r0 + mi * 8 + (lane >> 2), col = k4 * 4 + (lane & 3);
        af[mi][k4]    = (row < h && col < w) ? P[(long long)(c0 + col) * ld + row] : 0.0;
      }
  }
#pragma unroll
  for (int u = 0; u < EPT; ++u)
  {
    const int c   = warp + NWD * u;
    A[lane][c]    = (lane < w && c <= lane) ? P[(long long)(c0 + c) * ld + c0 + lane] : 0.0;
    Ainv[lane][c] = lane == c ? 1.0 : 0.0;
  }
  const double tau = scal[1];
  int nper         = 0;
  __syncthreads();
  for (int j = 0; j < w; ++j)
  {
    double dj = A[j][j];
    if (!(fabs(dj) >= tau) || !isfinite(dj))
    {
      dj = tau > 0.0 ? -tau : -1e-300;
      ++nper;
    }
    const double rdj = 1.0 / dj;
    if (tid == 0)
    {
      dsh[j]  = dj;
      dinv[j] = rdj;
    }
    if (lane > j && lane < w)
    {
      const double lij = A[lane][j] * rdj;
      double cur[EPT], oth[EPT];
#pragma unroll
      for (int u = 0; u < EPT; ++u)
      {
        const int c = warp + NWD * u;
        cur[u]      = c > j ? A[lane][c] : Ainv[lane][c];
        oth[u]      = c > j ? A[c][j] : Ainv[j][c];
      }
#pragma unroll
      for (int u = 0; u < EPT; ++u)
      {
        const int c = warp + NWD * u;
        if (c > j)
        {
          if (c <= lane)
          {
            A[lane][c] = cur[u] - lij * oth[u];
          }
        }
        else
        {
          Ainv[lane][c] = cur[u] - lij * oth[u];
        }
      }
    }
    __syncthreads();
  }
  if (t.rb == 0)
  {
    if (tid == 0 && nper)
    {
      atomicAdd(n_perturbed, nper);
    }
    if (tid < w)
    {
      D[s.first + c0 + tid]    = dsh[tid];
      Dinv[s.first + c0 + tid] = dinv[tid];
    }
    double* M = Mt + s.Lptr;
#pragma unroll
    for (int u = 0; u < EPT; ++u)
    {
      const int c = warp + NWD * u;
      if (lane < w && c <= lane)
      {
        M[(long long)(c0 + c) * ld + c0 + lane] = Ainv[lane][c];
      }
    }
  }
  // L21 = F21 * L11^-T D^-1 as a 32x32x32 DMMA product per warp with the inverted block
  for (int idx = tid; idx < NB * NB; idx += PANEL_THR)
  {
    const int n = idx % NB, kk = idx / NB;
    Wm[kk][n]   = (n < w && kk <= n) ? Ainv[n][kk] * dinv[n] : 0.0;
  }
  __syncthreads();
  if (r0 >= h)
  {
    return;
  }
  double acc[MI][4][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
    {
      acc[mi][nj][0] = 0.0;
      acc[mi][nj][1] = 0.0;
    }
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4)
  {
    double bf[4];
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
    {
      bf[nj] = Wm[k4 * 4 + (lane & 3)][nj * 8 + (lane >> 2)];
    }
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int nj = 0; nj < 4; ++nj)
      {
        dmma(acc[mi][nj][0], acc[mi][nj][1], af[mi][k4], bf[nj]);
      }
  }
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int nj = 0; nj < 4; ++nj)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int row = r0 + mi * 8 + (lane >> 2), col = nj * 8 + 2 * (lane & 3) + e;
        if (row < h && col < w)
        {
          P[(long long)(c0 + col) * ld + row] = acc[mi][nj][e];
        }
      }
}

// ---------------------------------------------------------------------------------------------
// 64x64 output tiles on the FP64 tensor cores. Operands are staged through a 4-stage cp.async (LDGSTS)
// pipeline in shared memory: chunk c+3 is in flight while chunk c is multiplied, one barrier per chunk.
// (A 1-deep register prefetch left the DMMA pipe 22 % active with long-scoreboard stalls dominating:
// profiles/README.md.)
constexpr int KC     = 16;       // k-chunk per stage
constexpr int LDT    = TILE + 4; // padded leading dimension: (LDT mod 16) == 4 -> conflict-free fragments
constexpr int STAGES = 4;
constexpr size_t TILE_SMEM = sizeof(double) * (size_t)(STAGES * (2 * KC * LDT + KC));

__device__ __forceinline__ void
cp_async8(double* smem_dst, const double* gsrc, bool valid)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz     = valid ? 8 : 0; // src-size 0: nothing is read, the destination is zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void
cp_async_commit()
{
  asm volatile("cp.async.commit_group;\n" ::);
}
template <int N>
__device__ __forceinline__ void
cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct TileSmem
{
  double (*As)[KC][LDT];
  double (*Bs)[KC][LDT];
  double (*ds)[KC];
  __device__ explicit TileSmem(double* base)
  {
    As = reinterpret_cast<double(*)[KC][LDT]>(base);
    Bs = reinterpret_cast<double(*)[KC][LDT]>(base + STAGES * KC * LDT);
    ds = reinterpret_cast<double(*)[KC]>(base + 2 * STAGES * KC * LDT);
  }
};

// acc += As^T * (Bs .* scale) for one staged chunk: As[kk][i], Bs[kk][j]; warp (wy, wx) owns a 32x32 quadrant
template <bool SCALE>
__device__ __forceinline__ void
mma_chunk(const double (*As)[LDT], const double (*Bs)[LDT], const double* dsc, double (&acc)[4][4][2], int wy, int wx, int lane)
{
#pragma unroll
  for (int k4 = 0; k4 < KC / 4; ++k4)
  {
    const int kr   = k4 * 4 + (lane & 3);
    const double d = SCALE ? dsc[kr] : 1.0;
    double af[4], bf[4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
    {
      af[a] = As[kr][wy * 32 + a * 8 + (lane >> 2)];
      bf[a] = Bs[kr][wx * 32 + a * 8 + (lane >> 2)];
      if (SCALE)
      {
        bf[a] *= d;
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
      {
        dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
      }
  }
}

// C(tile) -= Lrow_i * diag(d) * Lrow_j^T over front columns [kb, ke). Operand rows come from the
// panel (column-major, leading dimension h): A-rows start at ra0, B-rows at rb0, at most na / nb
// valid. Output element (i, j) of the tile lives at C[i + j * ldc]; only gi >= gj are touched,
// where gi/gj are the global row/column ordinals used for the lower-triangle mask.
__device__ __forceinline__ void
tile_update(const double* __restrict__ P,
            int h,
            int kb,
            int ke,
            const double* __restrict__ d,
            int ra0,
            int na,
            int rb0,
            int nb,
            double* __restrict__ C,
            long long ldc,
            int gi0,
            int gj0,
            bool accumulate,
            TileSmem sm,
            int loa = 0, // first valid tile row / column (1 for the first tiles of a shifted Schur grid, else 0)
            int lob = 0)
{
  const int tid  = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wy = warp >> 1, wx = warp & 1;
  // a 32x32 warp tile strictly above the diagonal has nothing to do (diagonal tiles only)
  const bool active = (gi0 + wy * 32 + 31) >= (gj0 + wx * 32);
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
      acc[a][b][0] = 0.0;
      acc[a][b][1] = 0.0;
    }
  const int nchunks = (ke - kb + KC - 1) / KC;
  // element u of a thread: idx = tid + 128 u -> (kk = idx / TILE, ii = idx % TILE): consecutive threads copy
  // consecutive rows (coalesced 8-byte cp.async)
  auto issue = [&](int c, int st) {
    const int kc = kb + c * KC;
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const int idx = tid + 128 * u;
      const int kk = idx / TILE, ii = idx % TILE;
      const int col    = kc + kk;
      const bool okc   = col < ke;
      const double* pc = P + (long long)(okc ? col : kb) * h;
      const bool va = ii < na && ii >= loa, vb = ii < nb && ii >= lob;
      cp_async8(&sm.As[st][kk][ii], pc + ra0 + (va ? ii : loa), okc && va);
      cp_async8(&sm.Bs[st][kk][ii], pc + rb0 + (vb ? ii : lob), okc && vb);
    }
    if (tid < KC)
    {
      const int col = kc + tid;
      cp_async8(&sm.ds[st][tid], d + (col < ke ? col : kb), col < ke);
    }
  };
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c)
  {
    if (c < nchunks)
    {
      issue(c, c);
    }
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c)
  {
    cp_async_wait<STAGES - 2>(); // chunk c has landed (for this thread's copies)
    __syncthreads();             // ... for everybody's, and stage (c-1) % STAGES is free again
    if (c + STAGES - 1 < nchunks)
    {
      issue(c + STAGES - 1, (c + STAGES - 1) % STAGES);
    }
    cp_async_commit();
    if (active)
    {
      const int st = c % STAGES;
      mma_chunk<true>(sm.As[st], sm.Bs[st], sm.ds[st], acc, wy, wx, lane);
    }
  }
  if (!active)
  {
    return;
  }
  // epilogue in two phases (all loads, then all stores): interleaving them serialises 32 dependent
  // load -> store round trips per thread (ncu: 25 us for a single 64x64 tile with K = 32)
  double cv[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int i = wy * 32 + a * 8 + (lane >> 2);
        const int j = wx * 32 + b * 8 + 2 * (lane & 3) + e;
        const bool ok = i < na && j < nb && i >= loa && j >= lob && gi0 + i >= gj0 + j;
        cv[a][b][e]   = (ok && accumulate) ? C[i + (long long)j * ldc] : 0.0;
      }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int i = wy * 32 + a * 8 + (lane >> 2);
        const int j = wx * 32 + b * 8 + 2 * (lane & 3) + e;
        if (i < na && j < nb && i >= loa && j >= lob && gi0 + i >= gj0 + j)
        {
          C[i + (long long)j * ldc] = cv[a][b][e] - acc[a][b][e];
        }
      }
}

// ---------------------------------------------------------------------------------------------
// TMA variant of the tile update for large fronts: the operand tiles are fetched by the copy engine
// (cp.async.bulk.tensor, SASS UTMALDG) straight from the column-major panel through a per-supernode tensor map --
// one elected thread issues eight 16 x 16 boxes per k-chunk and arms an mbarrier with the byte count; nobody computes
// addresses or predicates per element, rows beyond the front are zero-filled by the hardware. The boxes land in shared
// memory in the 128-byte swizzle (dense, no padding: 16 KB per stage instead of 17.4 KB): element (kk, r) of a box
// sits at kk * 128 B + ((r / 2) ^ (kk % 8)) * 16 B + (r % 2) * 8 B. A DMMA fragment takes 8 rows x 4 consecutive kk; with
// the rows of a fragment chosen as {0,1,8,9,2,3,10,11} (+4 for the second fragment of a box) the 16 lanes of a half
// warp hit 16 different 8-byte slots of the 128-byte bank window: conflict-free without padding.
constexpr int TMA_BOX         = 16;                              // rows and columns of one box
constexpr int TMA_BOX_DOUBLES = TMA_BOX * KC;                    // 2 KB
constexpr int TMA_STAGE_DOUBLES = 2 * (TILE / TMA_BOX) * TMA_BOX_DOUBLES; // A and B boxes of one stage: 16 KB
static_assert(KC == 16 && TMA_BOX * sizeof(double) == 128, "one box row is one 128-byte swizzle span");
static_assert(sizeof(double) * (STAGES * (TMA_STAGE_DOUBLES + KC)) + 8 * STAGES + 1024 <= TILE_SMEM, "the TMA layout fits the same allocation");

__device__ __forceinline__ unsigned
smem_u32(const void* p)
{
  return (unsigned)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void
mbar_init(unsigned bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void
mbar_expect_tx(unsigned bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void
mbar_wait(unsigned bar, unsigned parity)
{
  asm volatile("{\n"
               ".reg .pred P1;\n"
               "LAB_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
               "@P1 bra DONE;\n"
               "bra LAB_WAIT;\n"
               "DONE:\n"
               "}\n" ::"r"(bar),
               "r"(parity)
               : "memory");
}

// one 16 x 16 box of the panel: rows [row, row + 16) x columns [col, col + 16) -> dst (1 KB aligned), completion on bar
__device__ __forceinline__ void
tma_load_box(unsigned dst, const void* tmap, int row, int col, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst), "l"(tmap), "r"(row), "r"(col),
               "r"(bar)
               : "memory");
}

// row inside a 16-row box of lane group g (0..7) for fragment f (0..1): {0,1,8,9,2,3,10,11} + 4 f
__device__ __forceinline__ int
frag_row(int f, int g)
{
  return (g & 1) + 8 * ((g >> 1) & 1) + 2 * (g >> 2) + 4 * f;
}

// offset (doubles) of element (kk, r) inside a swizzled box
__device__ __forceinline__ int
swz(int kk, int r)
{
  return kk * TMA_BOX + ((((r >> 1) ^ (kk & 7)) << 1) | (r & 1));
}

__device__ __forceinline__ void
tile_update_tma(const void* __restrict__ tmap,
                int kb,
                int ke,
                const double* __restrict__ d,
                int ra0,
                int na,
                int rb0,
                int nb,
                double* __restrict__ C,
                long long ldc,
                int gi0,
                int gj0,
                bool accumulate,
                double* smem_raw,
                int loa = 0,
                int lob = 0)
{
  // 1 KB alignment of the swizzle atoms
  double* base          = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* dsm           = base + STAGES * TMA_STAGE_DOUBLES;                       // [STAGES][KC]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(dsm + STAGES * KC); // [STAGES]
  const int tid  = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wy = warp >> 1, wx = warp & 1;
  const bool active = (gi0 + wy * 32 + 31) >= (gj0 + wx * 32);
  if (tid == 0)
  {
#pragma unroll
    for (int st = 0; st < STAGES; ++st)
    {
      mbar_init(smem_u32(bars + st), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    // the tensor map lives in global memory and was written by a host copy: the thread that hands it to the copy
    // engine acquires it through the tensormap proxy first (system scope)
    asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;\n" ::"l"(tmap) : "memory");
  }
  __syncthreads();
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
      acc[a][b][0] = 0.0;
      acc[a][b][1] = 0.0;
    }
  const int nchunks = (ke - kb + KC - 1) / KC;
  auto issue = [&](int c, int st) {
    const int kc = kb + c * KC;
    if (tid == 0)
    {
      // the stage was read through the generic proxy before the barrier the caller just passed
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      const unsigned bar = smem_u32(bars + st);
      mbar_expect_tx(bar, (unsigned)(TMA_STAGE_DOUBLES * sizeof(double)));
      const unsigned dst = smem_u32(base + st * TMA_STAGE_DOUBLES);
#pragma unroll
      for (int q = 0; q < TILE / TMA_BOX; ++q)
      {
        tma_load_box(dst + q * TMA_BOX_DOUBLES * 8, tmap, ra0 + q * TMA_BOX, kc, bar);
        tma_load_box(dst + (TILE / TMA_BOX + q) * TMA_BOX_DOUBLES * 8, tmap, rb0 + q * TMA_BOX, kc, bar);
      }
    }
    if (tid < KC)
    {
      const int col = kc + tid;
      cp_async8(dsm + st * KC + tid, d + (col < ke ? col : kb), col < ke); // zero beyond ke: those columns do not count
    }
  };
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c)
  {
    if (c < nchunks)
    {
      issue(c, c);
    }
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c)
  {
    const int st = c % STAGES;
    cp_async_wait<STAGES - 2>();
    mbar_wait(smem_u32(bars + st), (unsigned)((c / STAGES) & 1));
    __syncthreads(); // the pivots of everybody's copies; and stage (c - 1) % STAGES is free again
    if (c + STAGES - 1 < nchunks)
    {
      issue(c + STAGES - 1, (c + STAGES - 1) % STAGES);
    }
    cp_async_commit();
    if (active)
    {
      const double* As = base + st * TMA_STAGE_DOUBLES;
      const double* Bs = As + (TILE / TMA_BOX) * TMA_BOX_DOUBLES;
      const double* dsc = dsm + st * KC;
#pragma unroll
      for (int k4 = 0; k4 < KC / 4; ++k4)
      {
        const int kr   = k4 * 4 + (lane & 3);
        const double dk = dsc[kr];
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
        {
          const int r = frag_row(a & 1, lane >> 2);
          af[a]       = As[(wy * 2 + (a >> 1)) * TMA_BOX_DOUBLES + swz(kr, r)];
          bf[a]       = Bs[(wx * 2 + (a >> 1)) * TMA_BOX_DOUBLES + swz(kr, r)] * dk;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
          {
            dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
          }
      }
    }
  }
  if (!active)
  {
    return;
  }
  // epilogue (two phases, like tile_update); rows and columns follow the fragment permutation
  double cv[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int i = wy * 32 + (a >> 1) * 16 + frag_row(a & 1, lane >> 2);
        const int j = wx * 32 + (b >> 1) * 16 + frag_row(b & 1, 2 * (lane & 3) + e);
        const bool ok = i < na && j < nb && i >= loa && j >= lob && gi0 + i >= gj0 + j;
        cv[a][b][e]   = (ok && accumulate) ? C[i + (long long)j * ldc] : 0.0;
      }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int i = wy * 32 + (a >> 1) * 16 + frag_row(a & 1, lane >> 2);
        const int j = wx * 32 + (b >> 1) * 16 + frag_row(b & 1, 2 * (lane & 3) + e);
        if (i < na && j < nb && i >= loa && j >= lob && gi0 + i >= gj0 + j)
        {
          C[i + (long long)j * ldc] = cv[a][b][e] - acc[a][b][e];
        }
      }
}

__global__ void __launch_bounds__(128)
k_update(const Task5* __restrict__ tasks,
         const SnMeta* __restrict__ sn,
         double* __restrict__ L,
         double* __restrict__ U,
         const double* __restrict__ D,
         const PanelTensorMap* __restrict__ tmaps)
{
  extern __shared__ double tile_smem[];
  TileSmem sm(tile_smem);
  const Task5 t  = tasks[blockIdx.x];
  const SnMeta s = sn[t.sn];
  const int h    = s.k + s.r;
  const int ld   = s.ld;
  double* P      = L + s.Lptr;
  if (s.tmap >= 0) // large front: operands through the copy engine (uniform per CTA)
  {
    const void* tmap = tmaps + s.tmap;
    if (t.kind == UPD_INPANEL)
    {
      const int na = min(TILE, h - t.i0), nb = min(TILE, t.jend - t.j0);
      tile_update_tma(tmap, t.kb, t.ke, D + s.first, t.i0, na, t.j0, nb, P + t.i0 + (long long)t.j0 * ld, ld, t.i0, t.j0, true, tile_smem);
    }
    else
    {
      // shifted tile grid (even front rows): tile row ii is row u0 + ii of the update matrix, row -1 is masked
      const int sh = s.k & 1, u0 = t.i0 - sh, v0 = t.j0 - sh;
      const int na = min(TILE, s.r - u0), nb = min(TILE, s.r - v0);
      tile_update_tma(tmap, 0, s.k, D + s.first, s.k + u0, na, s.k + v0, nb, U + s.Uoff + u0 + (long long)v0 * s.r, s.r, u0, v0,
                      s.child_end > s.child_begin, tile_smem, u0 < 0, v0 < 0);
    }
    return;
  }
  if (t.kind == UPD_INPANEL)
  {
    // front rows [i0, i0+64) x front columns [j0, j0+64), columns < k, rows < h
    const int na = min(TILE, h - t.i0), nb = min(TILE, t.jend - t.j0);
    tile_update(P, ld, t.kb, t.ke, D + s.first, t.i0, na, t.j0, nb, P + t.i0 + (long long)t.j0 * ld, ld, t.i0, t.j0, true, sm);
    return;
  }
  // UPD_SCHUR: update rows/cols [i0, i0+64) x [j0, j0+64) of U (r x r), operands are panel rows k + ...
  {
    const int sh = s.k & 1, u0 = t.i0 - sh, v0 = t.j0 - sh; // shifted tile grid, see above
    const int na = min(TILE, s.r - u0), nb = min(TILE, s.r - v0);
    double* Um   = U + s.Uoff;
    const bool accumulate = s.child_end > s.child_begin;
    tile_update(P, ld, 0, s.k, D + s.first, s.k + u0, na, s.k + v0, nb, Um + u0 + (long long)v0 * s.r, s.r, u0, v0, accumulate, sm, u0 < 0, v0 < 0);
  }
}

// ---------------------------------------------------------------------------------------------
// out(tile) = sign * X[:, kb:ke] * Y[kb:ke, :], both operands column-major, non-transposed:
//   X element (i, q) at X[i + q * ldx] (na valid rows), Y element (q, j) at Y[q + j * ldy] (nb valid columns).
__global__ void __launch_bounds__(128)
k_inv_gemm(const InvTask* __restrict__ tasks,
           const SnMeta* __restrict__ sn,
           const double* __restrict__ L,
           double* __restrict__ Mt,
           double* __restrict__ tmp)
{
  extern __shared__ double tile_smem[];
  TileSmem sm(tile_smem);
  const InvTask t = tasks[blockIdx.x];
  const SnMeta s  = sn[t.sn];
  const int k = s.k, h = s.ld; // h: leading dimension of the panels
  const double* P = L + s.Lptr;
  double* M       = Mt + s.Lptr;
  double* Tm      = tmp + s.Tptr;

  const double *X, *Y;
  double* out;
  long long ldx, ldy, ldo;
  int na, nb;
  double sign;
  if (t.kind == INV_T1)
  {
    // Tmp[i, j] = sum_q L11[i, q] Ainv[q, j]; A columns end at ke
    X = P + t.i0, ldx = h;
    Y = M + (long long)t.j0 * h, ldy = h;
    out = Tm + t.i0 + (long long)t.j0 * k, ldo = k;
    na = min(TILE, k - t.i0), nb = min(TILE, t.ke - t.j0);
    sign = 1.0;
  }
  else if (t.kind == INV_T2)
  {
    // Minv[i, j] = -sum_q Cinv[i, q] Tmp[q, j]; C rows [kb, ...), A columns end at kb
    X = M + t.i0, ldx = h;
    Y = Tm + (long long)t.j0 * k, ldy = k;
    out = M + t.i0 + (long long)t.j0 * h, ldo = h;
    na = min(TILE, t.ke - t.i0), nb = min(TILE, t.kb - t.j0);
    sign = -1.0;
  }
  else
  {
    // Minv[k + i, j] = -sum_q L21[i, q] Linv[q, j]
    X = P + k + t.i0, ldx = h;
    Y = M + (long long)t.j0 * h, ldy = h;
    out = M + k + t.i0 + (long long)t.j0 * h, ldo = h;
    na = min(TILE, s.r - t.i0), nb = min(TILE, k - t.j0);
    sign = -1.0;
  }

  const int tid  = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wy = warp >> 1, wx = warp & 1;
  double acc[4][4][2];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
    {
      acc[a][b][0] = 0.0;
      acc[a][b][1] = 0.0;
    }
  const int nchunks = (t.ke - t.kb + KC - 1) / KC;
  auto issue = [&](int c, int st) {
    const int kc = t.kb + c * KC;
#pragma unroll
    for (int u = 0; u < 8; ++u)
    {
      const int idx = tid + 128 * u;
      {
        const int kk = idx / TILE, ii = idx % TILE; // consecutive threads -> consecutive rows of X
        const int q   = kc + kk;
        const bool ok = q < t.ke && ii < na;
        cp_async8(&sm.As[st][kk][ii], X + (ok ? ii + (long long)q * ldx : 0), ok);
      }
      {
        const int jj = idx / KC, kk = idx % KC; // consecutive threads -> consecutive rows of Y
        const int q   = kc + kk;
        const bool ok = q < t.ke && jj < nb;
        cp_async8(&sm.Bs[st][kk][jj], Y + (ok ? q + (long long)jj * ldy : 0), ok);
      }
    }
  };
#pragma unroll
  for (int c = 0; c < STAGES - 1; ++c)
  {
    if (c < nchunks)
    {
      issue(c, c);
    }
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c)
  {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    if (c + STAGES - 1 < nchunks)
    {
      issue(c + STAGES - 1, (c + STAGES - 1) % STAGES);
    }
    cp_async_commit();
    const int st = c % STAGES;
    mma_chunk<false>(sm.As[st], sm.Bs[st], nullptr, acc, wy, wx, lane);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e)
      {
        const int i = wy * 32 + a * 8 + (lane >> 2);
        const int j = wx * 32 + b * 8 + 2 * (lane & 3) + e;
        if (i < na && j < nb)
        {
          out[i + (long long)j * ldo] = sign * acc[a][b][e];
        }
      }
}

// ---------------------------------------------------------------------------------------------
// Row-major copy of the inverse panels: Mr[r * k + j] = Mt[j * h + r], 32x32 tiles through shared memory.
__global__ void __launch_bounds__(256)
k_transpose(const TrTask* __restrict__ tasks, const SnMeta* __restrict__ sn, const double* __restrict__ Mt, double* __restrict__ Mr)
{
  __shared__ double tile[32][33];
  const TrTask t = tasks[blockIdx.x];
  const SnMeta s = sn[t.sn];
  const int k = s.k, h = s.k + s.r, ld = s.ld;
  const double* src = Mt + s.Lptr;
  double* dst       = Mr + s.Lptr;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int jj = ty; jj < 32; jj += 8)
  {
    const int i = t.i0 + tx, j = t.j0 + jj;
    tile[jj][tx] = (i < h && j < k) ? src[(long long)j * ld + i] : 0.0;
  }
  __syncthreads();
  for (int ii = ty; ii < 32; ii += 8)
  {
    const int i = t.i0 + ii, j = t.j0 + tx;
    if (i < h && j < k)
    {
      dst[(long long)i * k + j] = tile[tx][ii];
    }
  }
}

// ---------------------------------------------------------------------------------------------
void
configure_numeric_kernels(int device)
{
  // cudaFuncSetAttribute is per device: every device a handle is created on is opted in once (the caller has made
  // `device` current). Handles are created concurrently by independent solver threads.
  static std::mutex mu;
  static std::vector<int> done;
  std::lock_guard<std::mutex> lock(mu);
  if (std::find(done.begin(), done.end(), device) != done.end())
  {
    return;
  }
  B200_CUDA(cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM));
  B200_CUDA(cudaFuncSetAttribute(k_inv_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM));
  done.push_back(device);
}

void
enqueue_numeric(const DevPlan& dp, const NumericBuffers& nb, cudaStream_t stream, LaunchCounter& lc, const NumericOverlap* ov)
{
  const Plan& P = *dp.plan;
  B200_CUDA(cudaMemsetAsync(nb.scal, 0, sizeof(double) * 4, stream));
  B200_CUDA(cudaMemsetAsync(nb.n_perturbed, 0, sizeof(int), stream));
  if (P.m > 0)
  {
    B200_CUDA(cudaMemsetAsync(nb.L, 0, sizeof(double) * (size_t)P.Lptr[P.nsuper], stream));
    B200_CUDA(cudaMemsetAsync(nb.Mt, 0, sizeof(double) * (size_t)P.Lptr[P.nsuper], stream));
    const int threads = 256;
    const unsigned blocks = (unsigned)((P.nnzS + threads - 1) / threads);
    k_assemble<<<blocks, threads, 0, stream>>>(P.nnzS, dp.Sdest.p, dp.Sgsrc.p, dp.Sterm_ptr.p, dp.Sterm_a.p, dp.Sterm_b.p, dp.Sterm_d.p, nb.val, nb.L);
    lc.tick("assemble");
    const unsigned rb = (unsigned)std::min<long long>((P.m + threads - 1) / threads, 1184);
    k_diagmax<<<rb, threads, 0, stream>>>(P.m, dp.Sdiag.p, nb.L, (unsigned long long*)nb.scal);
    lc.tick();
    k_set_tau<<<1, 1, 0, stream>>>(nb.scal, P.n_demoted > 0 ? STATIC_PIVOT_QUASI : STATIC_PIVOT_DEFINITE);
    lc.tick();
    enqueue_sst_factor(dp, nb, stream, lc); // the sparse subtrees are leaves: their update blocks are ready before stage 0
  }
  // Look-ahead over two streams. Main: zero-fill / extend-add of the supernodes that start, the panel step, then the
  // update tiles of the NEXT panel's columns. Side: the rest of the stage's update tiles (and the Schur complements),
  // next to the following panel step. rest(s) needs panel(s); the look-ahead tiles of stage s and anything that
  // assembles children need rest(s - 1) (same tiles / the children's Schur complements).
  const int nstages = (int)P.stages.size();
  for (int si = 0; si < nstages; ++si)
  {
    const Stage& st = P.stages[si];
    const bool assembles = st.zero_end > st.zero_begin || st.ea_end > st.ea_begin;
    if (ov && assembles && si >= 1)
    {
      B200_CUDA(cudaStreamWaitEvent(stream, ov->rest_done[(si - 1) % 3], 0));
    }
    if (st.zero_end > st.zero_begin)
    {
      dim3 grid((unsigned)(st.zero_end - st.zero_begin), 16);
      k_zero_updates<<<grid, 256, 0, stream>>>(dp.zero_sn.p + st.zero_begin, dp.sn.p, nb.U);
      lc.tick("zero");
    }
    if (st.ea_end > st.ea_begin)
    {
      k_extend_add<<<(unsigned)(st.ea_end - st.ea_begin), 32 * EA_COLS, 0, stream>>>(dp.ea_tasks.p + st.ea_begin, dp.sn.p, dp.rel.p, nb.L, nb.U);
      lc.tick("extend_add");
    }
    if (st.pan_end > st.pan_begin)
    {
      k_panel<<<(unsigned)(st.pan_end - st.pan_begin), PANEL_THR, 0, stream>>>(dp.pan_tasks.p + st.pan_begin, dp.sn.p, nb.L, nb.Mt, nb.D, nb.Dinv, nb.scal, nb.n_perturbed);
      lc.tick("panel");
    }
    cudaStream_t rest_stream = stream;
    if (ov)
    {
      B200_CUDA(cudaEventRecord(ov->panel_done[si % 2], stream));
      B200_CUDA(cudaStreamWaitEvent(ov->side, ov->panel_done[si % 2], 0));
      rest_stream = ov->side;
      if (si >= 1)
      {
        B200_CUDA(cudaStreamWaitEvent(stream, ov->rest_done[(si - 1) % 3], 0));
      }
    }
    if (st.upd_mid > st.upd_begin)
    {
      k_update<<<(unsigned)(st.upd_mid - st.upd_begin), 128, TILE_SMEM, stream>>>(dp.upd_tasks.p + st.upd_begin, dp.sn.p, nb.L, nb.U, nb.D, dp.tmaps.p);
      lc.tick("update");
    }
    if (st.upd_end > st.upd_mid)
    {
      k_update<<<(unsigned)(st.upd_end - st.upd_mid), 128, TILE_SMEM, rest_stream>>>(dp.upd_tasks.p + st.upd_mid, dp.sn.p, nb.L, nb.U, nb.D, dp.tmaps.p);
      lc.tick("update");
    }
    if (ov)
    {
      B200_CUDA(cudaEventRecord(ov->rest_done[si % 3], ov->side));
    }
  }
  if (ov && nstages > 0)
  {
    B200_CUDA(cudaStreamWaitEvent(stream, ov->rest_done[(nstages - 1) % 3], 0)); // join
  }
  // selective inversion
  for (size_t ph = 0; ph + 1 < P.inv_phase_ptr.size(); ++ph)
  {
    const int b = P.inv_phase_ptr[ph], e = P.inv_phase_ptr[ph + 1];
    if (e > b)
    {
      k_inv_gemm<<<(unsigned)(e - b), 128, TILE_SMEM, stream>>>(dp.inv_tasks.p + b, dp.sn.p, nb.L, nb.Mt, nb.tmp);
      lc.tick("inv_gemm");
    }
  }
  if (!P.tr_tasks.empty())
  {
    k_transpose<<<(unsigned)P.tr_tasks.size(), 256, 0, stream>>>(dp.tr_tasks.p, dp.sn.p, nb.Mt, nb.Mr);
    lc.tick("transpose");
  }
  // contiguous operator values for the solve
  const int threads = 256;
  if (P.nE > 0)
  {
    k_gather<<<(P.nE + threads - 1) / threads, threads, 0, stream>>>(P.nE, dp.dE_src.p, nb.val, nb.dE);
    lc.tick();
  }
  const int nnzA = (int)P.Acsc_src.size();
  if (nnzA > 0)
  {
    k_gather<<<(nnzA + threads - 1) / threads, threads, 0, stream>>>(nnzA, dp.Acsc_src.p, nb.val, nb.Acsc_val);
    lc.tick();
    k_gather<<<(nnzA + threads - 1) / threads, threads, 0, stream>>>(nnzA, dp.Acsr_src.p, nb.val, nb.Acsr_val);
    lc.tick();
    k_gather_scaled<<<(nnzA + threads - 1) / threads, threads, 0, stream>>>(nnzA, dp.Acsr_src.p, dp.Acsr_dsrc.p, nb.val, nb.Acsr_sval);
    lc.tick();
  }
  const int nnzG = (int)P.Gsym_src.size();
  if (nnzG > 0)
  {
    k_gather<<<(nnzG + threads - 1) / threads, threads, 0, stream>>>(nnzG, dp.Gsym_src.p, nb.val, nb.Gsym_val);
    lc.tick();
  }
  B200_CUDA(cudaGetLastError());
}

} // namespace b200
