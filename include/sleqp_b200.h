/*
 * sleqp_b200.h -- C-ABI of libsleqp_b200.so, the B200-native sparse KKT backend for SLEQP.
 *
 * Plain C types only (no C++/torch types): this header is what the C11 host glue under
 * `sleqp_b200/host/` (fact/fact_b200.c, aug_jac/b200_aug_jac.c, tr/tr_b200.c, sparse/mat_b200.c)
 * includes, exactly like fact_umfpack.c includes <umfpack.h>.
 *
 * Each entry point names the reference interface it stands behind (paths relative to the
 * reference tree, chrhansk/sleqp v1.0.2):
 *
 *   SleqpFactCallbacks.set_matrix   src/main/fact/fact_types.h:9-10,  fact.c:59-75,
 *                                   backend example fact_umfpack.c:119-183
 *   SleqpFactCallbacks.solve        fact_types.h:12,  fact.c:83-89,   fact_umfpack.c:207-233
 *   SleqpFactCallbacks.solution     fact_types.h:14-18, fact.c:91-102, fact_umfpack.c:245-262
 *   SleqpFactCallbacks.condition    fact_types.h:20-21, fact.c:104-118, fact_cholmod.c:197-209
 *   SleqpFactCallbacks.free         fact_types.h:23,  fact.c:128-141, fact_umfpack.c:264-284
 *   sleqp_mat_mult_vec              src/main/sparse/mat.c:282-310
 *   sleqp_mat_mult_vec_trans        src/main/sparse/mat.c:312-363
 *   SleqpAugJacCallbacks            src/main/aug_jac/aug_jac.h; set_iterate = fill_aug_jac + set_matrix
 *     (b200_fact_set_kkt,           (aug_jac/standard_aug_jac.c:135-293), the three solves :306-435
 *      b200_fact_solve_offset)
 *   SleqpTRCallbacks                src/main/tr/tr_types.h:9-30; the loop restated: tr/steihaug_solver.c:223-496,
 *     (b200_cg_*)                   tr_dual :187-221, Rayleigh bounds :150-183, boundary step tr/tr_util.c:9-50
 *
 * There is NO CPU fallback: every function that computes returns B200_ERR_CUDA when no
 * device is usable. Only the b200_symbolic_* entry points (host-side analysis, which the
 * north-star keeps on the host) work without a GPU.
 *
 * All functions return 0 (B200_OK) on success, non-zero on failure; b200_last_error()
 * returns a thread-local message for the last failing call on this thread (the glue turns
 * it into sleqp_raise(SLEQP_INTERNAL_ERROR, ...), pub_error.h:38-48).
 */
#ifndef SLEQP_B200_H
#define SLEQP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_VERSION_STRING "0.1.0"

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

enum
{
  B200_OK              = 0,
  B200_ERR_ARG         = 1, /* malformed input (dims, unsorted rows, ...) */
  B200_ERR_CUDA        = 2, /* no device / CUDA runtime failure */
  B200_ERR_SINGULAR    = 3, /* numerically singular KKT (Umfpack convention: error, fact_umfpack.c:66-82) */
  B200_ERR_UNSUPPORTED = 4, /* structure outside the supported class (the Schur pattern would exceed 4e9 entries) */
  B200_ERR_STATE       = 5  /* call protocol violated (solve before set_matrix, ...) */
};

typedef struct b200_fact b200_fact; /* opaque factorization handle (one per SleqpFact) */
typedef struct b200_mat b200_mat;   /* opaque device CSC matrix (one per SleqpMat used for SpMV) */

/* Statistics for parity checks and roofline arithmetic. */
typedef struct b200_stats
{
  int32_t n;               /* order of K */
  int32_t n_elim;          /* diagonal (variable) nodes eliminated in closed form */
  int32_t n_reduced;       /* nodes of the Schur system S (constraints) */
  int64_t nnz_K;           /* stored entries of tril(K) */
  int64_t nnz_S;           /* entries of tril(S) */
  int64_t nnz_L;           /* exact nnz(L_S) incl. diagonal = sum of column counts */
  int64_t nnz_L_stored;    /* entries of the dense supernodal panels (>= nnz_L) */
  int64_t n_row_idx;       /* sum over supernodes of update rows */
  int32_t n_supernodes;
  int32_t n_levels;        /* supernodal etree height (solve levels) */
  int32_t n_stages;        /* numeric factorization stages */
  int32_t max_front;       /* largest front order k+r */
  double flops_factor;     /* sum_j cc_j^2 (exact structure) */
  double flops_factor_stored; /* flops actually executed on the padded panels */
  int64_t update_ws_doubles;  /* high-water mark of the update-matrix workspace */
  uint64_t pattern_hash;
  uint64_t perm_hash;      /* FNV-1a over the full permutation of K */
  int32_t symbolic_cached; /* 1 if the last set_matrix reused a cached analysis */
  int32_t n_perturbed;     /* pivots replaced by the static-pivot threshold */
  int32_t refine_steps;    /* refinement steps each solve performs for this factor */
  double probe_residual;   /* ||K x - b|| / ||b|| of the probe solve after refinement */
  double ms_symbolic;      /* host wall time of the analysis (0 when cached) */
  double ms_numeric;       /* device time of the last numeric factorization (CUDA events) */
  double ms_solve;         /* device time of the last solve (CUDA events) */
  int32_t n_scratch_slots; /* diagonal-block scratch slots of the numeric schedule */
  int32_t reserved;
  uint64_t pattern_hash2;  /* second, independent hash of the pattern (the plan cache is keyed by both) */
  double flops_update;     /* useful flops of the DMMA update tiles (in-panel + Schur): roofline numerator of k_update */
  double flops_inv;        /* useful flops of the selective-inversion tiles: roofline numerator of k_inv_gemm */
  int64_t panel_doubles;   /* doubles of one copy of the supernodal panels as laid out in HBM (L; Mt and Mr are as large) */
  int64_t n_demoted;       /* indices with a non-zero diagonal kept in the reduced system instead of being eliminated in
                              closed form: coupled to an eliminated index (non-diagonal (1,1) block) or a dense column */
} b200_stats;

/* ---- factorization plugin (SleqpFactCallbacks) --------------------------------------- */

/* device = -1: use B200_DEVICE from the environment, else LOCAL_RANK, else 0. */
B200_API int b200_fact_create(b200_fact** handle, int device);

/* K given as CSC (colptr[n_cols+1], rowidx[nnz] strictly increasing per column, val[nnz]),
 * lower triangle incl. diagonal when lower_only != 0 (SLEQP_FACT_FLAGS_LOWER), otherwise the
 * full symmetric matrix (strictly upper entries are then ignored). Arrays are borrowed for the
 * duration of the call only. Looks up / builds the cached symbolic analysis, uploads the
 * values and runs the numeric LDL^T on the device. */
B200_API int b200_fact_set_matrix(b200_fact* handle,
                         int n_rows,
                         int n_cols,
                         int nnz,
                         const int* colptr,
                         const int* rowidx,
                         const double* val,
                         int lower_only);

/* The same factorization from what the augmented Jacobian is built of (standard_aug_jac.c:135-237, aug_jac_set_iterate
 * :239-293): the constraint Jacobian of the iterate (CSC, num_cons x num_vars, as sleqp_iterate_cons_jac returns it) and
 * the working set as index maps (var_index[j] / cons_index[i] = position in the working set or -1, exactly
 * sleqp_working_set_var_index / _cons_index, working_set.c:117-180). tril([I A_W^T; A_W 0]) is laid out like
 * fill_aug_jac does, but only when (Jacobian pattern, working set) is new; for a known key the host does nothing but
 * hash the arrays: the Jacobian values are copied to the device and the KKT values are gathered there. Arrays are
 * borrowed for the call. */
B200_API int b200_fact_set_kkt(b200_fact* handle,
                               int num_vars,
                               int num_cons,
                               int nnz_jac,
                               const int* jac_cols,
                               const int* jac_rows,
                               const double* jac_data,
                               const int* var_index,
                               const int* cons_index,
                               int working_set_size);

/* Solve K x = b for a sparse right-hand side (idx ascending, dim == n). The result stays in
 * device memory inside the handle (like `umfpack->solution`). */
B200_API int b200_fact_solve(b200_fact* handle, int nnz_rhs, const int* idx, const double* val, int dim);
/* Same with every index shifted by `offset`: the right-hand side [0; rhs] of the min-norm solve without touching the
 * caller's vector (the reference adds num_vars to rhs->indices and takes it off again, standard_aug_jac.c:328-345). */
B200_API int b200_fact_solve_offset(b200_fact* handle, int nnz_rhs, const int* idx, const double* val, int offset, int dim);

/* Copy x[begin:end) of the last solve into out_dense (host memory, end-begin doubles). */
B200_API int b200_fact_solution(b200_fact* handle, int begin, int end, double* out_dense);

/* Same, but returns a pointer into the handle's pinned staging buffer (valid until the next
 * call on this handle) -- saves one host memcpy in the glue. */
B200_API int b200_fact_solution_ptr(b200_fact* handle, int begin, int end, const double** out);

/* The slice as a sparse vector, exactly what sleqp_vec_set_from_raw (sparse/vec.c:72-104) would build from it:
 * entries with |v| > zero_eps, ascending. idx_out / val_out must hold end-begin entries (the contents behind the
 * first *nnz_out entries are unspecified). When both buffers are page-locked (b200_host_pin, cudaHostAlloc, ...)
 * the slice is sparsified on the device and copied straight into them; otherwise it goes through the handle's
 * pinned staging buffer and a host pass. */
B200_API int b200_fact_solution_sparse(b200_fact* handle, int begin, int end, double zero_eps, int* idx_out, double* val_out, int* nnz_out);

/* Device-resident variants used by the benchmark's "inputs already in HBM" leg and by a
 * device-resident CG: d_rhs and d_sol are device pointers to n doubles. */
B200_API int b200_fact_solve_device(b200_fact* handle, const double* d_rhs, double* d_sol);

/* Numeric refactorization with new values already in device memory (same pattern and value order
 * as the last set_matrix; no probe solve, no host synchronisation): the "inputs resident in HBM"
 * leg of the benchmark and the building block for device-side KKT assembly. */
B200_API int b200_fact_refactor_device(b200_fact* handle, const double* d_val);

/* Device time (ms, CUDA events, mean of `reps` eager runs) of the four phases of one unrefined
 * solve: ms_out[0] E-block elimination, [1] forward sweep, [2] backward sweep, [3] back-substitution.
 * For roofline arithmetic in bench.py. */
B200_API int b200_fact_profile_solve(b200_fact* handle, int reps, double* ms_out);

/* Device time (ms, CUDA events, one eager run) of the numeric factorization by kernel class:
 * ms_out[0] assemble, [1] zero, [2] extend_add, [3] panel, [4] update, [5] inv_gemm, [6] transpose, [7] rest. */
B200_API int b200_fact_profile_numeric(b200_fact* handle, double* ms_out);

/* rcond = min|d_i| / max|d_i| over the pivots of D (cf. cholmod_l_rcond, fact_cholmod.c:204). */
B200_API int b200_fact_rcond(b200_fact* handle, double* rcond);

B200_API int b200_fact_stats(b200_fact* handle, b200_stats* stats);

/* Structural outputs of the cached analysis, for parity tests. Every array is optional (NULL
 * = skip). perm[n]: position p of the factorization holds K index perm[p]; parent[n]:
 * elimination tree of P K P^T (-1 = root); colcount[n]: column counts of L incl. diagonal;
 * super_first[n_supernodes_total+1] over the full order (variables are 1x1 supernodes). */
B200_API int b200_fact_structure(b200_fact* handle, int* perm, int* parent, int* colcount, int* n_super_total, int* super_first);

/* Debug/parity: copy the dense pivots D (n doubles, factorization order) to the host. */
B200_API int b200_fact_pivots(b200_fact* handle, double* d_out);

/* CUDA stream the handle launches on (cudaStream_t as void*), for event timing by callers. */
B200_API void* b200_fact_stream(b200_fact* handle);
/* CUDA device ordinal the handle lives on (-1 for a null handle). */
B200_API int b200_fact_device(b200_fact* handle);
/* Device buffers the solve graph of the handle reads its right-hand side from and leaves its solution in (N doubles
 * each, valid after a successful set_matrix / set_kkt until the next one). b200_fact_solve_device called with exactly
 * these pointers skips its two device copies: a device-resident caller (the projected CG) keeps its residual there. */
B200_API int b200_fact_device_buffers(b200_fact* handle, double** rhs, double** solution);

B200_API int b200_fact_free(b200_fact** handle);

B200_API const char* b200_last_error(void);

/* ---- host-only symbolic analysis (no GPU needed) -------------------------------------- */

typedef struct b200_symbolic b200_symbolic;

B200_API int b200_symbolic_analyze(b200_symbolic** out,
                          int n,
                          int nnz,
                          const int* colptr,
                          const int* rowidx,
                          const double* val,
                          int lower_only);
/* Analysis of the KKT system b200_fact_set_kkt would build from (Jacobian, working set); the plan field "Ksrc" tells
 * where every value of tril(K) comes from (-1: the constant 1, else an index into jac_data). */
B200_API int b200_symbolic_analyze_kkt(b200_symbolic** out, int num_vars, int num_cons, int nnz_jac, const int* jac_cols, const int* jac_rows,
                                       const double* jac_data, const int* var_index, const int* cons_index, int working_set_size);
B200_API int b200_symbolic_stats(const b200_symbolic* s, b200_stats* stats);
B200_API int b200_symbolic_structure(const b200_symbolic* s, int* perm, int* parent, int* colcount, int* n_super_total, int* super_first);
/* Export of the numeric plan (reduced system) for the CPU emulation used in tests:
 * query sizes with all pointers NULL first. See sleqp_b200/fact.py for the field list. */
B200_API int b200_symbolic_export(const b200_symbolic* s, const char* field, void* out, int64_t* count);
B200_API int b200_symbolic_free(b200_symbolic** s);

/* ---- CSC SpMV / SpMV^T (sleqp_mat_mult_vec / sleqp_mat_mult_vec_trans) ---------------- */

B200_API int b200_mat_create(b200_mat** handle, int device);
/* Upload a CSC matrix (same layout as SleqpMat: cols[num_cols+1], rows[nnz], data[nnz]). A
 * CSR mirror for the gather-form y = A x is built on the host when the pattern changes. */
B200_API int b200_mat_set(b200_mat* handle, int num_rows, int num_cols, int nnz, const int* cols, const int* rows, const double* data);
/* result[num_rows] = A * x, x sparse (mat.c:282-310). */
B200_API int b200_mat_mult_vec(b200_mat* handle, int nnz_x, const int* idx, const double* val, double* result_dense);
/* result[num_cols] = A^T * v, v sparse; dense result, the glue drops |s| <= eps (mat.c:312-363). */
B200_API int b200_mat_mult_vec_trans(b200_mat* handle, int nnz_v, const int* idx, const double* val, double* result_dense);
/* The same product as a sparse vector, exactly what sleqp_mat_mult_vec_trans returns (mat.c:312-363: ascending column
 * indices, entries with |s| <= eps dropped): sparsified on the device, only the kept entries are copied back.
 * idx_out / val_out must hold num_cols entries. */
B200_API int b200_mat_mult_vec_trans_sparse(b200_mat* handle, int nnz_v, const int* idx, const double* val, double eps, int* idx_out, double* val_out,
                                            int* nnz_out);
/* Device-resident forms: d_x/d_y dense device vectors. */
B200_API int b200_mat_mult_vec_device(b200_mat* handle, const double* d_x, double* d_y);
B200_API int b200_mat_mult_vec_trans_device(b200_mat* handle, const double* d_v, double* d_y);
/* y = A x unless the device flag *d_skip is non-zero (then y is left alone): lets a device-resident loop that has
 * already met its exit condition run out without touching its state (the projected CG enqueues several iterations
 * ahead of the host). d_skip may be NULL. */
B200_API int b200_mat_mult_vec_device_if(b200_mat* handle, const double* d_x, double* d_y, const int* d_skip);
B200_API void* b200_mat_stream(b200_mat* handle);
/* Launch this matrix's products on another CUDA stream (e.g. b200_fact_stream of the factorization the
 * products alternate with in the projected-CG loop); the handle does not own that stream. */
B200_API int b200_mat_set_stream(b200_mat* handle, void* stream);
B200_API int b200_mat_free(b200_mat** handle);

/* ---- device-resident projected CG (SleqpTRSolver.solve, src/main/tr/tr_solver.c:54; algorithm of
 * src/main/tr/steihaug_solver.c:223-496) ------------------------------------------------------------------- */

typedef struct b200_cg b200_cg;

enum
{
  B200_CG_INTERIOR      = 0, /* |r.g| below tolerance: returns z (steihaug_solver.c:318-327) */
  B200_CG_BOUNDARY      = 1, /* step hit the trust region (steihaug_solver.c:419-441) */
  B200_CG_NEG_CURVATURE = 2, /* d^T H d <= 0 (steihaug_solver.c:349-402) */
  B200_CG_MAX_ITER      = 3  /* iteration cap: like the reference, the step is ZERO (steihaug_solver.c:302-305) */
};

/* Borrows a factorization handle (K = [I A_W^T; A_W 0] already factorized) and a matrix handle holding the
 * Hessian of the Lagrangian (n x n, full symmetric CSC); both must outlive the CG handle and live on one device.
 * hess may be NULL when the Hessian is matrix-free: set a host callback instead (below). */
B200_API int b200_cg_create(b200_cg** handle, b200_fact* fact, b200_mat* hess);
/* Matrix-free Hessian (the reference's user callback SLEQP_FUNC_HESS_PROD, pub_func.h:168, reached through
 * sleqp_problem_hess_prod, problem.c:632): out[n] = H * dir[n], dense host vectors, non-zero return = failure. Used when
 * the handle was created without a device matrix; costs one D2H + H2D of n doubles per CG iteration, everything else
 * (projection, recurrences, exit tests) stays on the device. */
typedef int (*b200_hess_prod_fn)(void* ctx, int n, const double* dir, double* out);
B200_API int b200_cg_set_hess_callback(b200_cg* handle, b200_hess_prod_fn fn, void* ctx);
/* min g^T p + 1/2 p^T H p  s.t.  A_W p = 0, |p| <= trust_radius. gradient: sparse host vector of dimension n;
 * rel_tol = stat_tol * 1e-2 in the reference (steihaug_solver.c:21,241); max_iter < 0: unlimited.
 * step_out: n doubles (host). */
B200_API int b200_cg_solve(b200_cg* handle, int n, int nnz_g, const int* g_idx, const double* g_val, double trust_radius, double rel_tol,
                           int max_iter, double* step_out, int* iterations, int* termination);
/* Same solve with the other outputs of SleqpTRCallbacks (tr/tr_types.h:9-30), each optional (NULL = skip):
 * tr_dual: dual of the trust-region constraint, computed on the boundary exit only like steihaug_tr_dual
 * (steihaug_solver.c:187-221, 432-437), NaN otherwise (the reference leaves SLEQP_NONE); min/max_rayleigh: bounds of
 * d^T H d / d^T d over the directions of this solve, both starting at 1 (steihaug_solver.c:150-183, 234-235). */
B200_API int b200_cg_solve_ex(b200_cg* handle, int n, int nnz_g, const int* g_idx, const double* g_val, double trust_radius, double rel_tol,
                              int max_iter, double* step_out, int* iterations, int* termination, double* tr_dual, double* min_rayleigh,
                              double* max_rayleigh);
/* Same solve with the step returned as a sparse vector with the contract of sleqp_vec_set_from_raw (vec.c:72-104):
 * entries with |v| <= zero_eps dropped, ascending indices; step_idx / step_val hold n entries. When both arrays (and
 * the gradient's) are page-locked (b200_host_pin) the step is sparsified on the device and DMA'd straight into them,
 * with no host pass over the n values -- what tr/tr_b200.c does with the arrays of the caller's SleqpVec. */
B200_API int b200_cg_solve_sparse(b200_cg* handle, int n, int nnz_g, const int* g_idx, const double* g_val, double trust_radius, double rel_tol,
                                  int max_iter, double zero_eps, int* step_idx, double* step_val, int* step_nnz, int* iterations,
                                  int* termination, double* tr_dual, double* min_rayleigh, double* max_rayleigh);
B200_API int b200_cg_free(b200_cg** handle);

/* ---- misc ------------------------------------------------------------------------------ */
/* Page-locks / releases a caller-owned host buffer (cudaHostRegister). Every entry point that takes a host
 * buffer lets the copy engine read or write it directly when it is page-locked and stages through the
 * handle's own pinned buffer otherwise; an integrator that keeps its vectors alive across calls (as the
 * aug_jac does with its rhs / sol caches, standard_aug_jac.c:306-435) pins them once. */
B200_API int b200_host_pin(void* ptr, size_t bytes);
B200_API int b200_host_unpin(void* ptr);
B200_API int b200_device_count(void);
/* Number of kernel launches issued by this library on this process so far (bench.py's
 * "gpu_launches" claim; graph replays count their kernel nodes). */
B200_API int64_t b200_launch_count(void);

#ifdef __cplusplus
}
#endif

#endif /* SLEQP_B200_H */
