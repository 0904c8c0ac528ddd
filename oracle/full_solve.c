/*
 * full_solve.c -- TEST INFRASTRUCTURE ONLY (oracle). A complete sleqp_solver_solve through UNMODIFIED reference code
 * (solver.c, problem_solver*.c, cauchy/standard_cauchy.c, eqp.c, newton.c, linesearch.c, ...), linked once with the
 * reference LAPACK factorization and once with the B200 backend (fact_b200.c + the integration patch that selects
 * b200_aug_jac.c / tr_b200.c), both over the LP backend sleqp_b200/host/lp/lpi_simplex.c. The two binaries must print
 * the same iterate sequence to 1e-8 (north_star: "the same SLEQP iterate sequence").
 *
 * Problems:
 *   hs71   the reference's own constrained fixture (src/test/constrained_fixture.c:17-273, restated): n = 4, m = 2, known
 *          optimum (1, 4.742999, 3.821151, 1.379408) to 1e-6 (src/test/constrained_test.c:84-103)
 *   chain  config 1 of BASELINE.json: chained Rosenbrock with c_k = x_2k x_2k+1 + x_2k+2 - 1 = 0, -2 <= x <= 2
 *          (the functions of eqp_harness.c)
 *
 * usage: full_solve <hs71|chain> [n] [max_iterations]
 * output: "iterate <k> <n> x..." per accepted iterate, then "status", "iterations", "solution <n> x...", "elapsed_ms".
 */
#define HARNESS_NO_MAIN
#include "eqp_harness.c"

#include <string.h>
#include <time.h>

#include "pub_log.h"
#include "pub_solver.h"

// ---- HS71 (constrained_fixture.c, restated) ------------------------------------------------------------------------
typedef struct
{
  double x[4];
} HS71;

static SLEQP_RETCODE
hs_set(SleqpFunc* func, SleqpVec* x, SLEQP_VALUE_REASON reason, bool* reject, void* fd)
{
  HS71* d = (HS71*)fd;
  SLEQP_CALL(sleqp_vec_to_raw(x, d->x));
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
hs_obj_val(SleqpFunc* func, double* v, void* fd)
{
  const double* x = ((HS71*)fd)->x;
  *v              = x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2];
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
hs_obj_grad(SleqpFunc* func, SleqpVec* g, void* fd)
{
  const double* x = ((HS71*)fd)->x;
  SLEQP_CALL(sleqp_vec_clear(g));
  SLEQP_CALL(sleqp_vec_reserve(g, 4));
  SLEQP_CALL(sleqp_vec_push(g, 0, (x[0] + x[1] + x[2]) * x[3] + x[0] * x[3]));
  SLEQP_CALL(sleqp_vec_push(g, 1, x[0] * x[3]));
  SLEQP_CALL(sleqp_vec_push(g, 2, x[0] * x[3] + 1));
  SLEQP_CALL(sleqp_vec_push(g, 3, (x[0] + x[1] + x[2]) * x[0]));
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
hs_cons_val(SleqpFunc* func, SleqpVec* c, void* fd)
{
  const double* x = ((HS71*)fd)->x;
  SLEQP_CALL(sleqp_vec_clear(c));
  SLEQP_CALL(sleqp_vec_reserve(c, 2));
  SLEQP_CALL(sleqp_vec_push(c, 0, x[0] * x[1] * x[2] * x[3]));
  SLEQP_CALL(sleqp_vec_push(c, 1, x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]));
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
hs_cons_jac(SleqpFunc* func, SleqpMat* J, void* fd)
{
  const double* x = ((HS71*)fd)->x;
  SLEQP_CALL(sleqp_mat_reserve(J, 8));
  for (int j = 0; j < 4; ++j)
  {
    double prod = 1.;
    for (int q = 0; q < 4; ++q)
    {
      prod *= q == j ? 1. : x[q];
    }
    SLEQP_CALL(sleqp_mat_push_col(J, j));
    SLEQP_CALL(sleqp_mat_push(J, 0, j, prod));
    SLEQP_CALL(sleqp_mat_push(J, 1, j, 2 * x[j]));
  }
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
hs_hess_prod(SleqpFunc* func, const SleqpVec* direction, const SleqpVec* cons_duals, SleqpVec* product, void* fd)
{
  const double* x = ((HS71*)fd)->x;
  double dir[4], du[2] = {0., 0.}, H[4][4] = {{0.}};
  SLEQP_CALL(sleqp_vec_to_raw(direction, dir));
  if (cons_duals)
  {
    SLEQP_CALL(sleqp_vec_to_raw(cons_duals, du));
  }
  // objective x0 x3 (x0 + x1 + x2) + x2
  H[0][0] = 2 * x[3];
  H[0][1] = H[1][0] = x[3];
  H[0][2] = H[2][0] = x[3];
  H[0][3] = H[3][0] = 2 * x[0] + x[1] + x[2];
  H[1][3] = H[3][1] = x[0];
  H[2][3] = H[3][2] = x[0];
  // du0 * hessian of x0 x1 x2 x3, du1 * hessian of sum x^2
  for (int a = 0; a < 4; ++a)
  {
    H[a][a] += 2 * du[1];
    for (int b = 0; b < 4; ++b)
    {
      if (a == b)
      {
        continue;
      }
      double prod = 1.;
      for (int q = 0; q < 4; ++q)
      {
        prod *= (q == a || q == b) ? 1. : x[q];
      }
      H[a][b] += du[0] * prod;
    }
  }
  double out[4];
  for (int a = 0; a < 4; ++a)
  {
    out[a] = 0.;
    for (int b = 0; b < 4; ++b)
    {
      out[a] += H[a][b] * dir[b];
    }
  }
  SLEQP_CALL(sleqp_vec_set_from_raw(product, out, 4, 0.));
  return SLEQP_OKAY;
}

// ---- driver ----------------------------------------------------------------------------------------------------------
static int accepted = 0;

static SLEQP_RETCODE
on_accepted(SleqpSolver* solver, SleqpIterate* iterate, void* data)
{
  char name[32];
  snprintf(name, sizeof(name), "iterate_%d", accepted++);
  dump(name, sleqp_iterate_primal(iterate));
  return SLEQP_OKAY;
}

int
main(int argc, char** argv)
{
  const char* kind    = argc > 1 ? argv[1] : "hs71";
  const bool hs       = kind[0] == 'h';
  const int n         = hs ? 4 : (argc > 2 ? atoi(argv[2]) : 100);
  const int max_iter  = argc > 3 ? atoi(argv[3]) : 200;
  const int m         = hs ? 2 : (n - 2) / 2;
  const double inf    = sleqp_infinity();

  if (getenv("FULL_SOLVE_DEBUG"))
  {
    sleqp_log_set_level(SLEQP_LOG_DEBUG);
  }

  HS71 hsdata;
  Data data = {n, m, (double*)calloc(n, sizeof(double)), 0, 0, 1e-2};

  SleqpFuncCallbacks hs_cb    = {.set_value = hs_set, .obj_val = hs_obj_val, .obj_grad = hs_obj_grad, .cons_val = hs_cons_val, .cons_jac = hs_cons_jac,
                                 .hess_prod = hs_hess_prod, .func_free = NULL};
  SleqpFuncCallbacks chain_cb = {.set_value = f_set, .obj_val = f_obj_val, .obj_grad = f_obj_grad, .cons_val = f_cons_val, .cons_jac = f_cons_jac,
                                 .hess_prod = f_hess_prod, .func_free = NULL};
  SleqpFunc* func;
  CHECK(sleqp_func_create(&func, hs ? &hs_cb : &chain_cb, n, m, hs ? (void*)&hsdata : (void*)&data));

  SleqpVec *var_lb, *var_ub, *cons_lb, *cons_ub, *x0;
  CHECK(sleqp_vec_create_full(&var_lb, n));
  CHECK(sleqp_vec_create_full(&var_ub, n));
  CHECK(sleqp_vec_create_full(&cons_lb, m));
  CHECK(sleqp_vec_create_full(&cons_ub, m));
  CHECK(sleqp_vec_create_full(&x0, n));
  if (hs)
  {
    const double start[4] = {1., 5., 5., 1.};
    for (int i = 0; i < 4; ++i)
    {
      CHECK(sleqp_vec_push(var_lb, i, 1.));
      CHECK(sleqp_vec_push(var_ub, i, 5.));
      CHECK(sleqp_vec_push(x0, i, start[i]));
    }
    CHECK(sleqp_vec_push(cons_lb, 0, 25.));
    CHECK(sleqp_vec_push(cons_lb, 1, 40.));
    CHECK(sleqp_vec_push(cons_ub, 0, inf));
    CHECK(sleqp_vec_push(cons_ub, 1, 40.));
  }
  else
  {
    // config 1: bounds -2 <= x <= 2, equality constraints, x0 ~ U(0.5, 1.5) (xorshift64, seed as in eqp_harness.c)
    unsigned long long state = 88172645463325252ull;
    for (int i = 0; i < n; ++i)
    {
      state ^= state << 13;
      state ^= state >> 7;
      state ^= state << 17;
      CHECK(sleqp_vec_push(var_lb, i, -2.));
      CHECK(sleqp_vec_push(var_ub, i, 2.));
      CHECK(sleqp_vec_push(x0, i, 0.5 + (double)(state >> 11) / 9007199254740992.0));
    }
  }

  if (getenv("FULL_SOLVE_PERTURB"))
  {
    // sensitivity probe: the start point moved by a relative 1e-15 (one or two units in the last place)
    for (int k = 0; k < x0->nnz; ++k)
    {
      x0->data[k] *= 1. + atof(getenv("FULL_SOLVE_PERTURB")) * ((k % 3) - 1);
    }
  }

  SleqpSettings* settings;
  CHECK(sleqp_settings_create(&settings));
  CHECK(sleqp_settings_set_enum_value(settings, SLEQP_SETTINGS_ENUM_TR_SOLVER, SLEQP_TR_SOLVER_CG)); // trlib is absent
  SleqpProblem* problem;
  CHECK(sleqp_problem_create_simple(&problem, func, var_lb, var_ub, cons_lb, cons_ub, settings));

  SleqpSolver* solver;
  CHECK(sleqp_solver_create(&solver, problem, x0, NULL));
  CHECK(sleqp_solver_add_callback(solver, SLEQP_SOLVER_EVENT_ACCEPTED_ITERATE, (void*)on_accepted, NULL));

  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  CHECK(sleqp_solver_solve(solver, max_iter, 600.));
  clock_gettime(CLOCK_MONOTONIC, &t1);

  SleqpIterate* iterate;
  CHECK(sleqp_solver_solution(solver, &iterate));
  printf("status 1 %d\n", (int)sleqp_solver_status(solver));
  printf("iterations 1 %d\n", sleqp_solver_iterations(solver));
  printf("objective 1 %.17g\n", sleqp_iterate_obj_val(iterate));
  dump("solution", sleqp_iterate_primal(iterate));
  dump("cons_dual", sleqp_iterate_cons_dual(iterate));
  printf("elapsed_ms 1 %.3f\n", 1e3 * (t1.tv_sec - t0.tv_sec) + 1e-6 * (t1.tv_nsec - t0.tv_nsec));

  CHECK(sleqp_solver_release(&solver));
  CHECK(sleqp_problem_release(&problem));
  CHECK(sleqp_settings_release(&settings));
  return 0;
}
