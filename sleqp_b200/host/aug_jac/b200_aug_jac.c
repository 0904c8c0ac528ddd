/*
 * b200_aug_jac.c -- the augmented Jacobian system of SLEQP (aug_jac/aug_jac_types.h) with
 * assembly, factorization and solves on the B200.
 *
 * Meant to be dropped into the reference tree as src/main/aug_jac/b200_aug_jac.c next to
 * standard_aug_jac.c and created in trial_point.c:94-140 in place of the standard one (see
 * INTEGRATION.md). C11, no numerical code: everything runs in libsleqp_b200.so.
 *
 * Callbacks and their meaning follow standard_aug_jac.c:239-435 one to one:
 *   set_iterate        working set + constraint Jacobian -> factorization of [I A_W^T; A_W 0] (:239-293)
 *   solve_min_norm     K [x; y] = [0; rhs],  returns x            (:306-350)
 *   solve_lsq          K [x; y] = [rhs; 0],  returns y            (:352-394)
 *   project_nullspace  K [x; y] = [rhs; 0],  returns x            (:396-435)
 *   condition          estimate from the factorization, not exact (:295-304)
 */
#include "b200_aug_jac.h"

#include <assert.h>

#include <sleqp_b200.h>

#include "cmp.h"
#include "error.h"
#include "mem.h"
#include "problem.h"
#include "working_set.h"

#include "fact/fact_b200.h"

typedef struct
{
  SleqpProblem* problem;
  SleqpSettings* settings;

  b200_fact* handle;

  int working_set_size;
  bool has_factorization;

  // working set of the last factorization (linear problems: nothing to do while it stays the
  // same; all problems: the index maps below stay valid while it stays the same)
  SleqpWorkingSet* working_set;
  bool fixed_jacobian;
  bool maps_valid;

  // the working set as index maps (-1: not in the working set)
  int* var_index;
  int* cons_index;

  double condition;

  SleqpB200Pins pins;
} AugJacData;

#define B200_CALL(x)                                                           \
  do                                                                           \
  {                                                                            \
    const int b200_status = (x);                                               \
    if (b200_status != B200_OK)                                                \
    {                                                                          \
      sleqp_raise(SLEQP_INTERNAL_ERROR,                                        \
                  "B200 backend error %d: %s",                                 \
                  b200_status,                                                 \
                  b200_last_error());                                          \
    }                                                                          \
  } while (false)

static SLEQP_RETCODE
aug_jac_set_iterate(SleqpIterate* iterate, void* data)
{
  AugJacData* jacobian = (AugJacData*)data;

  SleqpProblem* problem        = jacobian->problem;
  SleqpWorkingSet* working_set = sleqp_iterate_working_set(iterate);

  // Do not recompute for linear problems & unchanged working set (standard_aug_jac.c:247-259).
  // The copy of the working set is kept for nonlinear problems too: while the working set stays
  // the same (the usual case once the active set has settled) the index maps handed to the
  // device library are still valid and the two passes over all variables and constraints that
  // rebuild them are skipped; the Jacobian values are of course sent every time.
  const bool same_working_set
    = jacobian->maps_valid
      && sleqp_working_set_eq(working_set, jacobian->working_set);

  if (same_working_set && jacobian->fixed_jacobian
      && jacobian->has_factorization)
  {
    sleqp_fact_b200_set_last_handle(jacobian->handle);
    return SLEQP_OKAY;
  }

  const int num_variables   = sleqp_problem_num_vars(problem);
  const int num_constraints = sleqp_problem_num_cons(problem);

  jacobian->condition = SLEQP_NONE;

  if (!same_working_set)
  {
    SLEQP_CALL(sleqp_working_set_copy(working_set, jacobian->working_set));

    jacobian->working_set_size = sleqp_working_set_size(working_set);

    for (int j = 0; j < num_variables; ++j)
    {
      jacobian->var_index[j] = sleqp_working_set_var_index(working_set, j);
    }

    for (int i = 0; i < num_constraints; ++i)
    {
      jacobian->cons_index[i] = sleqp_working_set_cons_index(working_set, i);
    }

    jacobian->maps_valid = true;
  }

  SleqpMat* cons_jac = sleqp_iterate_cons_jac(iterate);

  assert(sleqp_mat_num_cols(cons_jac) == num_variables);
  assert(sleqp_mat_num_rows(cons_jac) == num_constraints);

  B200_CALL(b200_fact_set_kkt(jacobian->handle,
                              num_variables,
                              num_constraints,
                              sleqp_mat_nnz(cons_jac),
                              sleqp_mat_cols(cons_jac),
                              sleqp_mat_rows(cons_jac),
                              sleqp_mat_data(cons_jac),
                              jacobian->var_index,
                              jacobian->cons_index,
                              jacobian->working_set_size));

  jacobian->has_factorization = true;

  // what a B200 trust-region solver on this thread projects with
  sleqp_fact_b200_set_last_handle(jacobian->handle);

  double rcond = 0.;

  B200_CALL(b200_fact_rcond(jacobian->handle, &rcond));

  jacobian->condition = 1. / rcond;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
aug_jac_condition(bool* exact, double* condition, void* data)
{
  AugJacData* jacobian = (AugJacData*)data;

  *exact     = false;
  *condition = jacobian->condition;

  return SLEQP_OKAY;
}

// the slice [begin, end) of the last solution as a sparse vector (what sleqp_fact_solution
// does through sleqp_vec_set_from_raw), sparsified on the device
static SLEQP_RETCODE
solution_slice(AugJacData* jacobian, SleqpVec* sol, int begin, int end)
{
  const double zero_eps
    = sleqp_settings_real_value(jacobian->settings, SLEQP_SETTINGS_REAL_ZERO_EPS);

  const int dim = end - begin;

  SLEQP_CALL(sleqp_vec_clear(sol));
  SLEQP_CALL(sleqp_vec_resize(sol, dim));
  SLEQP_CALL(sleqp_vec_reserve(sol, dim));

  sleqp_b200_pin_buffer(&jacobian->pins,
                        sol->data,
                        sizeof(double) * (size_t)sol->nnz_max);
  sleqp_b200_pin_buffer(&jacobian->pins,
                        sol->indices,
                        sizeof(int) * (size_t)sol->nnz_max);

  int nnz = 0;

  B200_CALL(b200_fact_solution_sparse(jacobian->handle,
                                      begin,
                                      end,
                                      zero_eps,
                                      sol->indices,
                                      sol->data,
                                      &nnz));

  sol->nnz = nnz;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
aug_jac_solve_min_norm(const SleqpVec* rhs, SleqpVec* sol, void* data)
{
  AugJacData* jacobian = (AugJacData*)data;

  const int num_variables = sleqp_problem_num_vars(jacobian->problem);
  const int total_size    = num_variables + jacobian->working_set_size;

  assert(sol->dim == num_variables);
  assert(rhs->dim == jacobian->working_set_size);

  sleqp_b200_pin_buffer(&jacobian->pins,
                        rhs->data,
                        sizeof(double) * (size_t)rhs->nnz_max);

  // [0; rhs]: the indices are shifted on the device, the caller's vector stays as it is
  B200_CALL(b200_fact_solve_offset(jacobian->handle,
                                   rhs->nnz,
                                   rhs->indices,
                                   rhs->data,
                                   num_variables,
                                   total_size));

  SLEQP_CALL(solution_slice(jacobian, sol, 0, num_variables));

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
aug_jac_solve_lsq(const SleqpVec* rhs, SleqpVec* sol, void* data)
{
  AugJacData* jacobian = (AugJacData*)data;

  const int num_variables = sleqp_problem_num_vars(jacobian->problem);
  const int total_size    = num_variables + jacobian->working_set_size;

  assert(rhs->dim == num_variables);
  assert(sol->dim == jacobian->working_set_size);

  sleqp_b200_pin_buffer(&jacobian->pins,
                        rhs->data,
                        sizeof(double) * (size_t)rhs->nnz_max);

  B200_CALL(b200_fact_solve(jacobian->handle,
                            rhs->nnz,
                            rhs->indices,
                            rhs->data,
                            total_size));

  SLEQP_CALL(solution_slice(jacobian, sol, num_variables, total_size));

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
aug_jac_project_nullspace(const SleqpVec* rhs, SleqpVec* sol, void* data)
{
  AugJacData* jacobian = (AugJacData*)data;

  const int num_variables = sleqp_problem_num_vars(jacobian->problem);
  const int total_size    = num_variables + jacobian->working_set_size;

  assert(rhs->dim == num_variables);
  assert(sol->dim == num_variables);

  sleqp_b200_pin_buffer(&jacobian->pins,
                        rhs->data,
                        sizeof(double) * (size_t)rhs->nnz_max);

  B200_CALL(b200_fact_solve(jacobian->handle,
                            rhs->nnz,
                            rhs->indices,
                            rhs->data,
                            total_size));

  SLEQP_CALL(solution_slice(jacobian, sol, 0, num_variables));

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
aug_jac_free(void* data)
{
  AugJacData* jacobian = (AugJacData*)data;

  sleqp_b200_unpin_all(&jacobian->pins);

  if (sleqp_fact_b200_last_handle() == jacobian->handle)
  {
    sleqp_fact_b200_set_last_handle(NULL);
  }

  B200_CALL(b200_fact_free(&jacobian->handle));

  sleqp_free(&jacobian->cons_index);
  sleqp_free(&jacobian->var_index);

  SLEQP_CALL(sleqp_working_set_release(&jacobian->working_set));

  SLEQP_CALL(sleqp_settings_release(&jacobian->settings));

  SLEQP_CALL(sleqp_problem_release(&jacobian->problem));

  sleqp_free(&jacobian);

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_b200_aug_jac_create(SleqpAugJac** star,
                          SleqpProblem* problem,
                          SleqpSettings* settings)
{
  AugJacData* jacobian = NULL;

  SLEQP_CALL(sleqp_malloc(&jacobian));

  *jacobian = (AugJacData){0};

  const int num_variables   = sleqp_problem_num_vars(problem);
  const int num_constraints = sleqp_problem_num_cons(problem);

  SLEQP_CALL(sleqp_problem_capture(problem));
  jacobian->problem = problem;

  SLEQP_CALL(sleqp_settings_capture(settings));
  jacobian->settings = settings;

  jacobian->condition = SLEQP_NONE;

  SLEQP_CALL(sleqp_alloc_array(&jacobian->var_index, num_variables));
  SLEQP_CALL(sleqp_alloc_array(&jacobian->cons_index, num_constraints));

  jacobian->fixed_jacobian = !(sleqp_problem_has_nonlinear_cons(problem));

  SLEQP_CALL(sleqp_working_set_create(&jacobian->working_set, problem));

  const int status = b200_fact_create(&jacobian->handle, -1);

  if (status != B200_OK)
  {
    sleqp_raise(SLEQP_INTERNAL_ERROR,
                "B200 backend error %d: %s",
                status,
                b200_last_error());
  }

  SleqpAugJacCallbacks callbacks
    = {.set_iterate       = aug_jac_set_iterate,
       .solve_min_norm    = aug_jac_solve_min_norm,
       .solve_lsq         = aug_jac_solve_lsq,
       .project_nullspace = aug_jac_project_nullspace,
       .condition         = aug_jac_condition,
       .free              = aug_jac_free};

  SLEQP_CALL(sleqp_aug_jac_create(star, problem, &callbacks, jacobian));

  return SLEQP_OKAY;
}
