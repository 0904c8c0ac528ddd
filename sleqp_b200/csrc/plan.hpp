// plan.hpp -- the cached symbolic analysis ("plan") shared by the host analysis and the
// device numeric/solve code. One Plan per sparsity pattern of tril(K); immutable once built.
//
// K = [D  A^T; A  G]  (reference: standard_aug_jac.c:135-237 builds it with D = I, G = 0)
//   E-nodes: columns with a non-zero diagonal and no off-diagonal coupling to other E-nodes
//            (the variables). They are eliminated in closed form (pivot d_e, L-column = A(:,e)/d_e).
//   R-nodes: the rest (working-set rows). The Schur system S = G - A D^-1 A^T on them is
//            factored by a supernodal multifrontal LDL^T.
// Full factorization order of K: [E-nodes in natural order | R-nodes in fill-reducing order],
// i.e. every constraint is eliminated after all variables it touches (SURVEY.md hard part 1).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b200
{

typedef int64_t i64;

constexpr int NB       = 32;  // panel width of the blocked dense LDL^T
constexpr int RB       = 128; // rows of L21 per CTA of a panel step (8 warps x 16 rows)
constexpr int TILE     = 64;  // output tile edge of the DMMA update kernel
constexpr int NBO      = 128; // outer block: columns beyond it are updated once per outer block with K = NBO
constexpr int EA_COLS  = 4;   // update-matrix columns per extend-add task (one warp each)
constexpr int FLOW_THREADS     = 256; // dataflow sweep kernels: persistent CTAs of 8 warps,
constexpr int FLOW_CTAS_PER_SM = 4;   // four per SM (64 registers per thread)
constexpr int FLOW_DEEP        = 128; // depth of a sweep task in the levels that are bandwidth-bound (symbolic.cpp, solve.cu)
constexpr int TMA_MIN_FRONT = 512; // fronts of at least this many rows fetch their update-tile operands through TMA (numeric.cu); measured: below ~500 rows the per-CTA barrier set-up costs more than the address arithmetic it saves (profiles/r02_tma_vs_cpasync.txt)
constexpr int LEAF_MAX = 32;  // etree subtrees up to this many columns become one dense supernode

// Leading dimension of a supernode's column-major panel (L and Mt): the front height rounded up to an even number of
// rows, so that every column starts on a 16-byte boundary -- what a TMA tensor map needs of its strides (numeric.cu).
// The row-major copy Mr of the same panel is dense h x k inside the same slot.
constexpr inline i64
panel_ld(i64 h)
{
  return (h + 1) & ~(i64)1;
}

// Sparse subtrees (SST). A complete subtree of the elimination tree whose columns hold only a few entries each (chains,
// banded systems: config 3) is kept as ONE supernode with its exact sparse structure instead of a dense panel: a dense
// 32-column block stores (and streams, in every sweep) ten times what a bidiagonal factor holds, and a tree of
// 32-column blocks is log32(m) hand-offs deep. One CTA factors / sweeps a whole subtree of up to SST_MAX_COLS columns in
// shared memory, level by level of its own elimination tree (columns of one level are independent).
constexpr int SST_MAX_COLS  = 1024; // columns of one sparse subtree
constexpr int SST_MIN_COLS  = 16;
constexpr int SST_MAX_TAIL  = 32;   // update rows (the subtree's contribution to its ancestors is a dense r x r block)
constexpr int SST_MAX_NNZ   = 6144; // entries of L in the subtree (values live in shared memory during the factorization)
constexpr int SST_MAX_AVG   = 8;    // mean entries per column: beyond that the dense supernodal path is the better one
constexpr int SST_THREADS   = 256;

// static pivot threshold relative to max |S_jj| (numeric.cu: k_set_tau)
constexpr double STATIC_PIVOT_DEFINITE = 64.0 * 2.220446049250313e-16;
constexpr double STATIC_PIVOT_QUASI    = 1.4901161193847656e-08;

constexpr size_t SST_SMEM_LIMIT = 200 * 1024; // what sst.cu opts its kernels into

struct SstMeta
{
  long long Lptr;  // values of the subtree in the panel buffer (compact, column by column, diagonal first)
  long long Uoff;  // its r x r update matrix
  int sn;          // supernode index
  int first, k, r;
  int Rptr;        // update rows in Ridx
  int signal;      // parent supernode (dense or sparse): its forward counter gets one signal when the subtree is swept, -1: root
  int blob;        // offset of the index blob in sst_blob (16-bit units, multiple of 8)
  int blob_len16;  // its length in 16-byte pieces
  int nslev;       // levels of segments
  int nseg;        // segments (single-child chains of the subtree's elimination tree)
  int nnz;         // entries
  int gen;         // generation: 0 = leaves of the supernodal tree, g = all children are subtrees of generations < g
  int o_segstart, o_seglen, o_colptr, o_rows; // parts of the blob after the level pointers (16-bit units from blob)
  int ea_begin, ea_end;                        // assembly of the children's update blocks: entries of sst_ea_src / _dst
  int col_ptr, row_ptr;                        // host copies: offsets into Plan::sst_colptr / sst_rows (32-bit)
  int nchild;                                  // child subtrees: signals the forward sweep waits for
  int parent_sst;                              // parent supernode if that is a sparse subtree (the backward sweep waits for its flag), else -1
  int o_rowptr, o_rcol, o_rpos;                // the off-diagonal entries by ROW (forward sweep: a column gathers): row pointers over
                                               // the k + r front rows, per entry its column and its position in the values
  int pad_;
};
static_assert(sizeof(SstMeta) == 120, "SstMeta layout");

// kinds of update tasks
enum
{
  UPD_INPANEL  = 0, // trailing update inside the supernode's own panel
  UPD_SCHUR    = 1, // U -= L21 D L21^T over the whole supernode width (last step only)
  UPD_DIAGCOPY = 2  // (unused since the diagonal block has its own kernel)
};

struct Task5
{
  int sn, jend, kind, i0, j0; // jend: first column the tile must not touch
  int kb, ke;                 // inner (front column) range
};

struct PanelTask
{
  int sn, t, rb, pad; // panel step t of a supernode: CTA rb owns RB rows of L21 below the diagonal block (there is
                      // always a CTA 0, it publishes the pivots and the inverse of the diagonal block)
};

struct EaTask
{
  int child, jb;
};

// Tile task of the selective-inversion phases (after the factorization proper):
//   INV_T1: Tmp[C x A]  =  L11[C, A] * Ainv            (Ainv lower triangular, from the inverse panel)
//   INV_T2: Minv[C x A] = -Cinv * Tmp[C x A]           (Cinv lower triangular)
//   INV_Z : Minv[tail, :] = -L21 * L11^-1
// i0/j0: first row / column of the 64x64 output tile in front coordinates, [kb, ke): inner range.
enum
{
  INV_T1 = 0,
  INV_T2 = 1,
  INV_Z  = 2
};

struct InvTask
{
  int sn, kind, i0, j0, kb, ke;
};

struct TrTask
{
  int sn, i0, j0; // 32x32 tile of a panel to transpose into the row-major copy
};

// Warp task of the dataflow sweeps (solve.cu: k_flow). One warp owns a block of a supernode's inverse panel,
// rows [i0, i1) x columns [j0, j1), one lane per output and at most 16 panel entries per lane:
//   forward : lanes = rows (i1 - i0 <= 32), depth = columns (j1 - j0 <= 16, or 32 in the wide levels), column-major panel Mt; partial sums
//             are added to yf (top block rows) or pushed to their final rows of the accumulator (tail rows);
//   backward: lanes = columns (j1 - j0 <= 32, j0 a multiple of 32), depth = rows (i1 - i0 <= 16 or 32, i0 a multiple
//             of 16), row-major copy Mr; partial sums are added to x.
// Dependencies are counters, one per supernode: a task waits until cnt[wait_idx] >= need (wait_idx < 0: no wait)
// and adds 1 to cnt[signal_idx] when its results are visible (signal_idx < 0: nobody waits for it).
//   forward : wait on the own supernode (its children signal it), signal the parent;
//   backward: wait on the parent (only tasks that touch tail rows), signal the own supernode.
// Tasks are stored in topological order and handed out by ticket counters in that order (solve.cu), so the lowest
// unfinished task is always running or about to be drawn: no deadlock however few warps are resident.
struct alignas(16) SweepTask
{
  long long Lptr; // panel offset (doubles)
  int Rptr;       // offset of the supernode's update rows in Ridx
  int first;      // first column (new labels)
  int k, h; // h: leading dimension of the column-major panel (panel_ld of the front height)
  int i0, i1;
  int j0, j1;
  int wait_idx, need;
  int signal_idx;
  int pad0, pad1, pad2; // pad0: level of the supernode (timeline traces); pad1: 1 in the narrow levels (fewer tasks than
                        // warps: latency-critical, the completion is signalled before the next ticket is drawn)
};
static_assert(sizeof(SweepTask) == 64, "SweepTask layout");

struct Stage
{
  int zero_begin, zero_end;
  int ea_begin, ea_end;
  int pan_begin, pan_end;
  // update tiles of the stage: [upd_begin, upd_mid) touch the columns of the NEXT panel step only ("look-ahead" part,
  // on the critical path), [upd_mid, upd_end) everything else (runs next to the next panel step on a second stream)
  int upd_begin, upd_mid, upd_end;
  int pad;
};

struct Plan
{
  int N = 0, nE = 0, m = 0;
  i64 nnzK = 0;
  i64 nnzK_input = 0; // entries of the caller's CSC (== nnzK for lower-triangular input)
  uint64_t pattern_hash = 0, pattern_hash2 = 0, perm_hash = 0; // the plan cache is keyed by (N, nnzK_input, both pattern hashes)

  // Plans built from (constraint Jacobian, working set) by b200_fact_set_kkt: the cache key is a hash of THOSE arrays
  // (pattern_hash / pattern_hash2 / key_aux = nnz of the Jacobian), and Ksrc tells where every value of tril(K)
  // comes from: -1 = the constant 1 (identity diagonal, row of an active bound), else an index into the Jacobian's
  // values (standard_aug_jac.c:135-237)
  bool kkt_keyed = false;
  i64 key_aux    = 0;
  std::vector<int> Ksrc;

  // node classification / index maps
  std::vector<int> e_of_k, r_of_k, k_of_e, k_of_r;
  std::vector<int> dE_src;                       // K-value index of d_e
  std::vector<int> Acsc_ptr, Acsc_row, Acsc_src; // by E column; rows = original R index
  std::vector<int> Acsr_ptr, Acsr_col, Acsr_src; // by R row (original index); cols = E index
  std::vector<int> Gsym_ptr, Gsym_col, Gsym_src; // symmetric R-R coupling by R row (original)
  // resolved index chains for the solve: K index of the column's variable and value index of its pivot (by CSR
  // entry), permuted reduced row (by CSC entry)
  std::vector<int> Acsr_k, Acsr_dsrc, Acsc_p;

  // ordering of the reduced system
  std::vector<int> perm, pinv; // new -> old R index, old -> new
  std::vector<int> parent, colcount;

  // supernodes
  int nsuper = 0;
  std::vector<int> sn_first; // nsuper+1
  std::vector<int> sn_of_col;
  std::vector<int> sn_parent, sn_level;
  std::vector<i64> Rptr; // nsuper+1, offsets into Ridx/rel
  std::vector<int> Ridx; // update rows (new labels), ascending
  std::vector<int> rel;  // position of each update row in the parent's front
  std::vector<i64> Lptr; // nsuper+1, panel offsets (doubles); panel is h x k column-major
  std::vector<i64> Wptr; // nsuper+1, prefix sum of front heights
  std::vector<int> child_ptr, child_idx;

  // sparse subtrees
  std::vector<char> sn_sparse; // nsuper: 1 if the supernode is a sparse subtree
  std::vector<SstMeta> sst;
  std::vector<int> sst_colptr, sst_rows; // 32-bit host copies (assembly map, emulation)
  std::vector<unsigned short> sst_blob;  // what the device reads
  std::vector<long long> sst_ea_src;     // assembly of child subtrees: offset of the entry in the update workspace ...
  std::vector<int> sst_ea_dst;           // ... and where it goes: >= 0 slot of the subtree's values, < 0: -1 - index in its update block
  std::vector<int> sst_gen_ptr;          // P.sst is sorted by generation: [gen_ptr[g], gen_ptr[g + 1])
  int n_demoted = 0; // indices with a non-zero diagonal kept in the reduced system (coupled to an E node, or a dense column)
  size_t sst_smem_bytes = 0; // dynamic shared memory of the sst kernels: what the largest subtree of this plan needs

  // assembly of S straight into the panels: S_e = val[gsrc] - sum_t val[a]*val[b]/val[d]
  i64 nnzS = 0;
  std::vector<i64> Sdest;
  std::vector<i64> Sdiag; // panel offset of S(j,j) for every column j
  std::vector<int> Sgsrc;
  std::vector<i64> Sterm_ptr;
  std::vector<int> Sterm_a, Sterm_b, Sterm_d;

  // update-matrix workspace (lifetime-packed)
  std::vector<i64> Uoff;
  i64 Utotal = 0;

  // numeric schedule
  std::vector<int> sn_base, sn_nt;
  std::vector<Stage> stages;
  std::vector<int> zero_sn;
  std::vector<EaTask> ea_tasks;
  std::vector<PanelTask> pan_tasks;
  std::vector<Task5> upd_tasks;
  int n_scratch_slots = 0;

  // selective inversion: phases of tile tasks; phase p covers inv_tasks[inv_phase_ptr[p], inv_phase_ptr[p+1])
  std::vector<InvTask> inv_tasks;
  std::vector<int> inv_phase_ptr;
  std::vector<TrTask> tr_tasks;
  std::vector<i64> Tptr; // nsuper+1, k x k scratch of the inversion (only supernodes wider than NB)

  // levels of the supernodal tree (height above the leaves)
  int nlevels = 0;
  std::vector<int> lvl_ptr, lvl_sn;
  // dataflow sweeps: tasks in topological (ticket) order
  std::vector<SweepTask> ffl_tasks, bfl_tasks;

  // statistics
  i64 nnzL = 0, nnzL_stored = 0;
  double flops = 0, flops_stored = 0;
  double flops_update = 0, flops_inv = 0; // useful flops of the k_update / k_inv_gemm tile tasks
  int max_front = 0;
  double ms_symbolic = 0;
};

// Builds the plan. Returns 0 or a B200_ERR_* code with a message in err.
int analyze(int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only, Plan& plan, std::string& err);

// cheap O(n) check of the column pointers; must hold before hash_pattern / analyze dereference through them
bool valid_csc_header(int n, int nnz, const int* colptr, const int* rowidx, const double* val);

// two independent 64-bit hashes of (n, pattern, which diagonals are non-zero): returns the first, *second gets the other
uint64_t hash_pattern(int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only, uint64_t* second = nullptr);

// Key of a KKT system given as (Jacobian pattern, working-set index maps): two independent hashes like hash_pattern.
uint64_t hash_kkt(int num_vars, int num_cons, int nnz_jac, const int* jac_cols, const int* jac_rows, const int* var_index, const int* cons_index, int ws_size,
                  uint64_t* second);

// tril([I A_W^T; A_W 0]) exactly as the reference's fill_aug_jac lays it out (standard_aug_jac.c:135-237, lower_only):
// per variable column the unit diagonal, the row of its active bound (if any), then its Jacobian entries in the rows of
// the active constraints; empty columns for the working set. src: see Plan::Ksrc. Returns false on malformed input.
bool build_kkt_lower(int num_vars, int num_cons, const int* jac_cols, const int* jac_rows, const int* var_index, const int* cons_index, int ws_size,
                     std::vector<int>& colptr, std::vector<int>& rowidx, std::vector<int>& src);

// Expands (perm, parent, colcount, supernodes) of the reduced system to the full order of K.
void full_structure(const Plan& p, int* perm, int* parent, int* colcount, int* n_super_total, int* super_first);

} // namespace b200
