/*
 * tr_b200.c -- SLEQP trust-region (EQP) solver "B200": the three SleqpTRCallbacks
 * (src/main/tr/tr_types.h:9-30) over the device-resident projected CG of libsleqp_b200.so.
 *
 * Meant to be dropped into the reference tree as src/main/tr/tr_b200.c next to
 * steihaug_solver.c and selected in newton.c:111-123. It is C11, uses
 * only reference-internal headers plus <sleqp_b200.h>, and contains no numerical code.
 *
 * Algorithm, exits, tolerance (stat_tol * 1e-2), iteration cap (zero step), Rayleigh bounds
 * and the dual of the trust region are those of tr/steihaug_solver.c:150-221,223-496; what
 * changes is that r, g, d, z and B d never leave the GPU: one solve is one sparse gradient
 * in, one step out, instead of one sparse-vector round trip through SleqpFact per iteration.
 */
#include "tr_b200.h"

#include <math.h>

#include <sleqp_b200.h>

#include "cmp.h"
#include "error.h"
#include "mem.h"
#include "problem.h"

#include "fact/fact_b200.h"

static const double tolerance_factor = 1e-2; // steihaug_solver.c:21

typedef struct
{
  SleqpProblem* problem;
  SleqpSettings* settings;

  int max_iter;

  double min_rayleigh;
  double max_rayleigh;

  b200_cg* cg;           // created lazily: needs the factorization's device handle
  b200_fact* cg_fact;    // ... the one it was created for
  bool cg_uses_matrix;

  b200_mat* hessian;     // device copy of the Hessian, when one has been provided
  bool hessian_set;

  // matrix-free Hessian products (sleqp_problem_hess_prod) of the callback mode
  const SleqpVec* multipliers;
  SleqpVec* direction;
  SleqpVec* product;
  bool callback_failed;


  int last_iterations;
  int last_exit;
  SleqpB200Pins pins; // page-locked arrays of the caller's gradient / step vectors
} TRData;

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

// SleqpTRSolver keeps its solver_data private (tr_solver.c:5-16): the setters below find the
// data of a solver in a small table instead
#define B200_MAX_SOLVERS 64

static struct
{
  SleqpTRSolver* solver;
  TRData* data;
} solver_table[B200_MAX_SOLVERS];

static TRData*
find_data(SleqpTRSolver* solver)
{
  for (int i = 0; i < B200_MAX_SOLVERS; ++i)
  {
    if (solver_table[i].solver == solver)
    {
      return solver_table[i].data;
    }
  }

  return NULL;
}

static int
hess_prod_callback(void* ctx, int n, const double* direction, double* product)
{
  TRData* data = (TRData*)ctx;

  const double zero_eps
    = sleqp_settings_real_value(data->settings, SLEQP_SETTINGS_REAL_ZERO_EPS);

  if (sleqp_vec_set_from_raw(data->direction, (double*)direction, n, zero_eps)
        != SLEQP_OKAY
      || sleqp_problem_hess_prod(data->problem,
                                 data->direction,
                                 data->multipliers,
                                 data->product)
           != SLEQP_OKAY
      || sleqp_vec_to_raw(data->product, product) != SLEQP_OKAY)
  {
    data->callback_failed = true;
    return 1;
  }

  return 0;
}

static SLEQP_RETCODE
b200_tr_solve(SleqpAugJac* jacobian,
              const SleqpVec* multipliers,
              const SleqpVec* gradient,
              SleqpVec* newton_step,
              double trust_radius,
              double* tr_dual,
              double time_limit,
              void* solver_data)
{
  TRData* data = (TRData*)solver_data;

  (void)jacobian;   // the projection uses its B200 factorization directly on the device
  (void)time_limit; // a solve is a bounded number of device iterations; checked by the caller in between

  data->min_rayleigh = 1.;
  data->max_rayleigh = 1.;

  *tr_dual = SLEQP_NONE;

  const double stat_eps
    = sleqp_settings_real_value(data->settings, SLEQP_SETTINGS_REAL_STAT_TOL);
  const double zero_eps
    = sleqp_settings_real_value(data->settings, SLEQP_SETTINGS_REAL_ZERO_EPS);

  const double rel_tol = stat_eps * tolerance_factor;

  const int num_vars = sleqp_problem_num_vars(data->problem);

  b200_fact* fact = sleqp_fact_b200_last_handle();

  if (!fact)
  {
    sleqp_raise(SLEQP_INTERNAL_ERROR,
                "B200 trust-region solver: no B200 factorization has been "
                "set on this thread (the augmented Jacobian must use the B200 "
                "backend)");
  }

  if (!data->cg || data->cg_fact != fact
      || data->cg_uses_matrix != data->hessian_set)
  {
    B200_CALL(b200_cg_free(&data->cg));

    B200_CALL(b200_cg_create(&data->cg,
                             fact,
                             data->hessian_set ? data->hessian : NULL));

    if (!data->hessian_set)
    {
      B200_CALL(
        b200_cg_set_hess_callback(data->cg, hess_prod_callback, (void*)data));
    }

    data->cg_fact        = fact;
    data->cg_uses_matrix = data->hessian_set;
  }

  data->multipliers     = multipliers;
  data->callback_failed = false;

  double dual = 0.;

  // The gradient is read from, and the step is written to, the arrays of the caller's
  // SleqpVec objects by the copy engine: registered once (SLEQP reuses these vectors across
  // iterations), re-registered if sleqp_vec_reserve moved them. The step comes back
  // sparsified on the device (contract of sleqp_vec_set_from_raw).
  SLEQP_CALL(sleqp_vec_clear(newton_step));
  SLEQP_CALL(sleqp_vec_reserve(newton_step, num_vars));

  sleqp_b200_pin_buffer(&data->pins, gradient->data, sizeof(double) * (size_t)gradient->nnz_max);
  sleqp_b200_pin_buffer(&data->pins, gradient->indices, sizeof(int) * (size_t)gradient->nnz_max);
  sleqp_b200_pin_buffer(&data->pins, newton_step->data, sizeof(double) * (size_t)newton_step->nnz_max);
  sleqp_b200_pin_buffer(&data->pins, newton_step->indices, sizeof(int) * (size_t)newton_step->nnz_max);

  int step_nnz = 0;

  const int status = b200_cg_solve_sparse(data->cg,
                                          num_vars,
                                          gradient->nnz,
                                          gradient->indices,
                                          gradient->data,
                                          trust_radius,
                                          rel_tol,
                                          data->max_iter, // SLEQP_NONE == -1: no cap
                                          zero_eps,
                                          newton_step->indices,
                                          newton_step->data,
                                          &step_nnz,
                                          &data->last_iterations,
                                          &data->last_exit,
                                          &dual,
                                          &data->min_rayleigh,
                                          &data->max_rayleigh);

  data->multipliers = NULL;

  if (data->callback_failed)
  {
    // the error of the failed reference call is still set
    return SLEQP_ERROR;
  }

  B200_CALL(status);

  newton_step->nnz = step_nnz;

  if (!isnan(dual))
  {
    *tr_dual = dual;
  }

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_tr_rayleigh(double* min_rayleigh, double* max_rayleigh, void* solver_data)
{
  TRData* data = (TRData*)solver_data;

  (*min_rayleigh) = data->min_rayleigh;
  (*max_rayleigh) = data->max_rayleigh;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_tr_free(void** star)
{
  TRData* data = (TRData*)(*star);

  if (!data)
  {
    return SLEQP_OKAY;
  }

  for (int i = 0; i < B200_MAX_SOLVERS; ++i)
  {
    if (solver_table[i].data == data)
    {
      solver_table[i].solver = NULL;
      solver_table[i].data   = NULL;
    }
  }

  B200_CALL(b200_cg_free(&data->cg));
  B200_CALL(b200_mat_free(&data->hessian));

  sleqp_b200_unpin_all(&data->pins);

  SLEQP_CALL(sleqp_vec_free(&data->product));
  SLEQP_CALL(sleqp_vec_free(&data->direction));

  SLEQP_CALL(sleqp_settings_release(&data->settings));
  SLEQP_CALL(sleqp_problem_release(&data->problem));

  sleqp_free(&data);

  *star = NULL;

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_b200_tr_solver_create(SleqpTRSolver** solver_star,
                            SleqpProblem* problem,
                            SleqpSettings* settings)
{
  TRData* data = NULL;

  const int num_vars = sleqp_problem_num_vars(problem);

  SLEQP_CALL(sleqp_malloc(&data));

  *data = (TRData){0};

  data->problem = problem;
  SLEQP_CALL(sleqp_problem_capture(data->problem));

  SLEQP_CALL(sleqp_settings_capture(settings));
  data->settings = settings;

  data->max_iter
    = sleqp_settings_int_value(settings,
                               SLEQP_SETTINGS_INT_MAX_NEWTON_ITERATIONS);

  data->min_rayleigh = 1.;
  data->max_rayleigh = 1.;

  SLEQP_CALL(sleqp_vec_create_empty(&data->direction, num_vars));
  SLEQP_CALL(sleqp_vec_create_empty(&data->product, num_vars));

  SleqpTRCallbacks callbacks = {.solve    = b200_tr_solve,
                                .rayleigh = b200_tr_rayleigh,
                                .free     = b200_tr_free};

  SLEQP_CALL(sleqp_tr_solver_create(solver_star, &callbacks, (void*)data));

  for (int i = 0; i < B200_MAX_SOLVERS; ++i)
  {
    if (!solver_table[i].solver)
    {
      solver_table[i].solver = *solver_star;
      solver_table[i].data   = data;
      break;
    }
  }

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_tr_b200_set_hessian(SleqpTRSolver* solver, const SleqpMat* hessian)
{
  TRData* data = find_data(solver);

  if (!data)
  {
    sleqp_raise(SLEQP_ILLEGAL_ARGUMENT, "not a B200 trust-region solver");
  }

  if (!hessian)
  {
    data->hessian_set = false;
    return SLEQP_OKAY;
  }

  const int num_vars = sleqp_problem_num_vars(data->problem);

  if (sleqp_mat_num_rows(hessian) != num_vars
      || sleqp_mat_num_cols(hessian) != num_vars)
  {
    sleqp_raise(SLEQP_ILLEGAL_ARGUMENT,
                "Hessian must be %d x %d",
                num_vars,
                num_vars);
  }

  if (!data->hessian)
  {
    B200_CALL(b200_mat_create(&data->hessian, -1));
  }

  B200_CALL(b200_mat_set(data->hessian,
                         num_vars,
                         num_vars,
                         sleqp_mat_nnz(hessian),
                         sleqp_mat_cols(hessian),
                         sleqp_mat_rows(hessian),
                         sleqp_mat_data(hessian)));

  data->hessian_set = true;

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_tr_b200_last_solve(SleqpTRSolver* solver, int* iterations, int* exit_code)
{
  TRData* data = find_data(solver);

  if (!data)
  {
    sleqp_raise(SLEQP_ILLEGAL_ARGUMENT, "not a B200 trust-region solver");
  }

  *iterations = data->last_iterations;
  *exit_code  = data->last_exit;

  return SLEQP_OKAY;
}
