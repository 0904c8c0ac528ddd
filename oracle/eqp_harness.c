/*
 * eqp_harness.c -- TEST INFRASTRUCTURE ONLY (oracle). The EQP part of one SLEQP iteration driven
 * entirely by UNMODIFIED reference code, over whichever factorization backend is linked in:
 *
 *   problem / iterate / working set          src/main/problem.c, iterate.c, working_set.c
 *   sleqp_standard_aug_jac_create            src/main/aug_jac/standard_aug_jac.c:513-534
 *   sleqp_aug_jac_set_iterate                aug_jac.c:46 -> fill_aug_jac + sleqp_fact_set_matrix
 *   sleqp_aug_jac_solve_min_norm / _lsq / project_nullspace   standard_aug_jac.c:306-435
 *   Steihaug projected CG                    src/main/tr/steihaug_solver.c:223-496
 *
 * This is the SLEQP-iterate-sequence proxy of SURVEY.md section 8c (a full sleqp_solver_solve needs
 * an LP backend that is absent here): built once against libsleqp_ref_lapack.so (reference LAPACK
 * backend) and once against libsleqp_ref_b200.so (fact_b200.c), the two binaries must print the same
 * numbers to 1e-8 -- in particular the projected-CG path, sampled through a ladder of trust radii.
 *
 * Test problem (the configs 1/3 family): f = sum_i 100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2,
 * c_k = x_{2k} x_{2k+1} + x_{2k+2} - 1 = 0, -2 <= x <= 2. Working set set by hand like
 * src/test/constrained_newton_test.c:196-201: every constraint, plus every 10th even variable at its
 * upper bound.
 *
 * Second problem ("poisson", the configs 2/4/5 family): 2D Poisson control on a g x g grid, x = (y, u),
 * f = 1/2 |y - y_d|^2 + alpha/2 |u|^2, c = A y - u = 0 with the h^2-scaled 5-point Laplacian, bounds on u only;
 * working set: every constraint plus every 7th control at its upper bound.
 *
 * usage: eqp_harness <n> <number_of_trust_radii> [chain|poisson]     (poisson: <n> is the grid size g)
 * output: one "name count v0 v1 ..." line per quantity (%.17g).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cmp.h"
#include "func.h"
#include "iterate.h"
#include "mem.h"
#include "problem.h"
#include "util.h"
#include "working_set.h"

#include "aug_jac/standard_aug_jac.h"
#include "fact/fact.h"
#include "tr/steihaug_solver.h"
#include "tr/tr_solver.h"

#define CHECK(x)                                                               \
  do                                                                           \
  {                                                                            \
    if ((x) != SLEQP_OKAY)                                                     \
    {                                                                          \
      fprintf(stderr, "FAILED %s: %s\n", #x, sleqp_error_msg());               \
      exit(2);                                                                 \
    }                                                                          \
  } while (0)

typedef struct
{
  int n, m;
  double* x;
  int poisson; // 0: chained Rosenbrock, 1: 2D Poisson control
  int g;
  double alpha;
} Data;

static double
poisson_target(const Data* d, int i)
{
  const int ix = i % d->g, iy = i / d->g;
  return sin(M_PI * (ix + 1.) / (d->g + 1.)) * sin(M_PI * (iy + 1.) / (d->g + 1.));
}

static SLEQP_RETCODE
f_set(SleqpFunc* func, SleqpVec* x, SLEQP_VALUE_REASON reason, bool* reject, void* fd)
{
  Data* d = (Data*)fd;
  SLEQP_CALL(sleqp_vec_to_raw(x, d->x));
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
f_obj_val(SleqpFunc* func, double* v, void* fd)
{
  Data* d = (Data*)fd;
  double s = 0.;
  if (d->poisson)
  {
    const int q = d->m;
    for (int i = 0; i < q; ++i)
    {
      const double a = d->x[i] - poisson_target(d, i), u = d->x[q + i];
      s += 0.5 * a * a + 0.5 * d->alpha * u * u;
    }
    *v = s;
    return SLEQP_OKAY;
  }
  for (int i = 0; i + 1 < d->n; ++i)
  {
    const double a = d->x[i + 1] - d->x[i] * d->x[i], b = 1. - d->x[i];
    s += 100. * a * a + b * b;
  }
  *v = s;
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
f_obj_grad(SleqpFunc* func, SleqpVec* g, void* fd)
{
  Data* d = (Data*)fd;
  SLEQP_CALL(sleqp_vec_clear(g));
  SLEQP_CALL(sleqp_vec_reserve(g, d->n));
  for (int i = 0; i < d->n; ++i)
  {
    double v = 0.;
    if (d->poisson)
    {
      v = i < d->m ? d->x[i] - poisson_target(d, i) : d->alpha * d->x[i];
      SLEQP_CALL(sleqp_vec_push(g, i, v));
      continue;
    }
    if (i + 1 < d->n)
    {
      v += -400. * d->x[i] * (d->x[i + 1] - d->x[i] * d->x[i]) - 2. * (1. - d->x[i]);
    }
    if (i > 0)
    {
      v += 200. * (d->x[i] - d->x[i - 1] * d->x[i - 1]);
    }
    SLEQP_CALL(sleqp_vec_push(g, i, v));
  }
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
f_cons_val(SleqpFunc* func, SleqpVec* c, void* fd)
{
  Data* d = (Data*)fd;
  SLEQP_CALL(sleqp_vec_clear(c));
  SLEQP_CALL(sleqp_vec_reserve(c, d->m));
  for (int k = 0; k < d->m; ++k)
  {
    if (d->poisson)
    {
      const int g = d->g, ix = k % g, iy = k / g;
      double v = 4. * d->x[k] - d->x[d->m + k];
      v -= ix > 0 ? d->x[k - 1] : 0.;
      v -= ix + 1 < g ? d->x[k + 1] : 0.;
      v -= iy > 0 ? d->x[k - g] : 0.;
      v -= iy + 1 < g ? d->x[k + g] : 0.;
      SLEQP_CALL(sleqp_vec_push(c, k, v));
      continue;
    }
    SLEQP_CALL(sleqp_vec_push(c, k, d->x[2 * k] * d->x[2 * k + 1] + d->x[2 * k + 2] - 1.));
  }
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
f_cons_jac(SleqpFunc* func, SleqpMat* J, void* fd)
{
  Data* d = (Data*)fd;
  SLEQP_CALL(sleqp_mat_reserve(J, d->poisson ? 6 * d->m : 3 * d->m));
  for (int j = 0; j < d->n; ++j)
  {
    SLEQP_CALL(sleqp_mat_push_col(J, j));
    if (d->poisson)
    {
      const int g = d->g, q = d->m;
      if (j >= q)
      {
        SLEQP_CALL(sleqp_mat_push(J, j - q, j, -1.)); // -I block
        continue;
      }
      const int ix = j % g, iy = j / g; // column j of the symmetric Laplacian, rows ascending
      if (iy > 0)
      {
        SLEQP_CALL(sleqp_mat_push(J, j - g, j, -1.));
      }
      if (ix > 0)
      {
        SLEQP_CALL(sleqp_mat_push(J, j - 1, j, -1.));
      }
      SLEQP_CALL(sleqp_mat_push(J, j, j, 4.));
      if (ix + 1 < g)
      {
        SLEQP_CALL(sleqp_mat_push(J, j + 1, j, -1.));
      }
      if (iy + 1 < g)
      {
        SLEQP_CALL(sleqp_mat_push(J, j + g, j, -1.));
      }
      continue;
    }
    // column j: rows in ascending order
    if (j % 2 == 0)
    {
      const int kprev = j / 2 - 1, k = j / 2;
      if (kprev >= 0 && kprev < d->m)
      {
        SLEQP_CALL(sleqp_mat_push(J, kprev, j, 1.)); // x_{2k+2} of c_{kprev}
      }
      if (k < d->m)
      {
        SLEQP_CALL(sleqp_mat_push(J, k, j, d->x[j + 1])); // d/dx_{2k}
      }
    }
    else
    {
      const int k = (j - 1) / 2;
      if (k < d->m)
      {
        SLEQP_CALL(sleqp_mat_push(J, k, j, d->x[j - 1])); // d/dx_{2k+1}
      }
    }
  }
  return SLEQP_OKAY;
}

static SLEQP_RETCODE
f_hess_prod(SleqpFunc* func, const SleqpVec* dir, const SleqpVec* duals, SleqpVec* res, void* fd)
{
  Data* d    = (Data*)fd;
  const int n = d->n;
  double* v   = (double*)calloc(n, sizeof(double));
  double* lam = (double*)calloc(d->m > 0 ? d->m : 1, sizeof(double));
  double* out = (double*)calloc(n, sizeof(double));
  SLEQP_CALL(sleqp_vec_to_raw(dir, v));
  if (duals)
  {
    SLEQP_CALL(sleqp_vec_to_raw(duals, lam));
  }
  for (int i = 0; i < n && d->poisson; ++i)
  {
    out[i] = (i < d->m ? 1. : d->alpha) * v[i]; // constraints are linear: the Hessian of the Lagrangian is diagonal
  }
  for (int i = 0; i < n && !d->poisson; ++i)
  {
    double diag = 0.;
    if (i + 1 < n)
    {
      diag += 1200. * d->x[i] * d->x[i] - 400. * d->x[i + 1] + 2.;
      out[i] += -400. * d->x[i] * v[i + 1];
      out[i + 1] += -400. * d->x[i] * v[i];
    }
    if (i > 0)
    {
      diag += 200.;
    }
    out[i] += diag * v[i];
  }
  for (int k = 0; k < d->m && !d->poisson; ++k)
  {
    out[2 * k] += lam[k] * v[2 * k + 1];
    out[2 * k + 1] += lam[k] * v[2 * k];
  }
  SLEQP_CALL(sleqp_vec_set_from_raw(res, out, n, 0.));
  free(v);
  free(lam);
  free(out);
  return SLEQP_OKAY;
}

static void
dump(const char* name, const SleqpVec* v)
{
  double* raw = (double*)calloc(v->dim > 0 ? v->dim : 1, sizeof(double));
  CHECK(sleqp_vec_to_raw(v, raw));
  printf("%s %d", name, v->dim);
  for (int i = 0; i < v->dim; ++i)
  {
    printf(" %.17g", raw[i]);
  }
  printf("\n");
  free(raw);
}

int
main(int argc, char** argv)
{
  const int arg1    = argc > 1 ? atoi(argv[1]) : 100;
  const int max_it  = argc > 2 ? atoi(argv[2]) : 8;
  const int poisson = argc > 3 && argv[3][0] == 'p';
  const int n       = poisson ? 2 * arg1 * arg1 : arg1;
  const int m       = poisson ? arg1 * arg1 : (n - 2) / 2;
  Data data         = {n, m, (double*)calloc(n, sizeof(double)), poisson, arg1, 1e-2};

  SleqpFuncCallbacks callbacks = {.set_value = f_set,
                                  .obj_val   = f_obj_val,
                                  .obj_grad  = f_obj_grad,
                                  .cons_val  = f_cons_val,
                                  .cons_jac  = f_cons_jac,
                                  .hess_prod = f_hess_prod,
                                  .func_free = NULL};
  SleqpFunc* func;
  CHECK(sleqp_func_create(&func, &callbacks, n, m, &data));

  SleqpVec *var_lb, *var_ub, *cons_lb, *cons_ub, *x0;
  CHECK(sleqp_vec_create_full(&var_lb, n));
  CHECK(sleqp_vec_create_full(&var_ub, n));
  for (int i = 0; i < n; ++i) // poisson: the states y are free, the controls u are bounded
  {
    const double bound = (poisson && i < m) ? sleqp_infinity() : 2.;
    CHECK(sleqp_vec_push(var_lb, i, -bound));
    CHECK(sleqp_vec_push(var_ub, i, bound));
  }
  CHECK(sleqp_vec_create_full(&cons_lb, m));
  CHECK(sleqp_vec_create_full(&cons_ub, m));
  CHECK(sleqp_vec_create_full(&x0, n));
  unsigned long long state = 88172645463325252ull; // xorshift64: reproducible without libc rand
  for (int i = 0; i < n; ++i)
  {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    CHECK(sleqp_vec_push(x0, i, 0.5 + (double)(state >> 11) / 9007199254740992.0));
  }

  SleqpSettings* settings;
  CHECK(sleqp_settings_create(&settings));
  SleqpProblem* problem;
  CHECK(sleqp_problem_create_simple(&problem, func, var_lb, var_ub, cons_lb, cons_ub, settings));
  SleqpIterate* iterate;
  CHECK(sleqp_iterate_create(&iterate, problem, x0));
  CHECK(sleqp_set_and_evaluate(problem, iterate, SLEQP_VALUE_REASON_NONE, NULL));

  SleqpWorkingSet* ws = sleqp_iterate_working_set(iterate);
  CHECK(sleqp_working_set_reset(ws));
  int n_active_vars = 0;
  for (int j = poisson ? m : 0; j < n; j += poisson ? 7 : 20) // chain: every 10th even variable; poisson: every 7th control
  {
    CHECK(sleqp_working_set_add_var(ws, j, SLEQP_ACTIVE_UPPER));
    ++n_active_vars;
  }
  for (int k = 0; k < m; ++k)
  {
    CHECK(sleqp_working_set_add_cons(ws, k, SLEQP_ACTIVE_BOTH));
  }
  const int ws_size = n_active_vars + m;

  SleqpFact* fact;
  CHECK(sleqp_fact_create_default(&fact, settings));
  printf("backend 1 %d\n", (int)sleqp_fact_flags(fact));
  fprintf(stderr, "backend: %s\n", sleqp_fact_name(fact));
  SleqpAugJac* jac;
  CHECK(sleqp_standard_aug_jac_create(&jac, problem, settings, fact));
  CHECK(sleqp_aug_jac_set_iterate(jac, iterate));

  SleqpVec* grad = sleqp_iterate_obj_grad(iterate);
  SleqpVec *proj, *minnorm_rhs, *minnorm, *duals;
  CHECK(sleqp_vec_create_empty(&proj, n));
  CHECK(sleqp_vec_create_full(&minnorm_rhs, ws_size));
  CHECK(sleqp_vec_create_empty(&minnorm, n));
  CHECK(sleqp_vec_create_empty(&duals, ws_size));
  for (int i = 0; i < ws_size; ++i)
  {
    CHECK(sleqp_vec_push(minnorm_rhs, i, sin(0.37 * i) + 0.1));
  }

  CHECK(sleqp_aug_jac_project_nullspace(jac, grad, proj));
  dump("project_nullspace", proj);
  CHECK(sleqp_aug_jac_solve_min_norm(jac, minnorm_rhs, minnorm));
  dump("solve_min_norm", minnorm);
  CHECK(sleqp_aug_jac_solve_lsq(jac, grad, duals));
  dump("solve_lsq", duals);

  // The projected-CG path: the reference returns a zero step when the iteration cap is hit
  // (steihaug_solver.c:302-305 breaks before z is copied out), so intermediate iterates are sampled
  // through the trust radius instead: the step for radius Delta is the point where the piecewise
  // linear path z_0, z_1, ... leaves the ball (steihaug_solver.c:413-437), a function of every
  // iterate and direction before it. A ladder of radii samples the path segment by segment.
  SleqpVec* cons_dual = sleqp_iterate_cons_dual(iterate);
  // let the CG converge: with the default cap of 100 and stationarity tolerance 1e-6 the loop ends in the
  // zero step on this problem (steihaug_solver.c:302-305)
  CHECK(sleqp_settings_set_int_value(settings, SLEQP_SETTINGS_INT_MAX_NEWTON_ITERATIONS, 4 * n));
  CHECK(sleqp_settings_set_real_value(settings, SLEQP_SETTINGS_REAL_STAT_TOL, 1e-2));
  double full_norm = 0.;
  for (int it = -1; it < max_it; ++it)
  {
    // it = -1: unconstrained by the radius (the converged Newton step); then radii inside its norm
    const double radius = it < 0 ? 1e8 : full_norm * (0.6 + 0.4 * (it + 0.5) / max_it);
    SleqpTRSolver* tr;
    CHECK(sleqp_steihaug_solver_create(&tr, problem, settings));
    SleqpVec* step;
    CHECK(sleqp_vec_create_empty(&step, n));
    double tr_dual = 0.;
    CHECK(sleqp_tr_solver_solve(tr, jac, cons_dual, grad, step, radius, &tr_dual));
    char name[64];
    if (it < 0)
    {
      full_norm = sleqp_vec_norm(step);
      snprintf(name, sizeof(name), "cg_converged_step");
    }
    else
    {
      snprintf(name, sizeof(name), "cg_path_sample_%d", it);
    }
    dump(name, step);
    CHECK(sleqp_vec_free(&step));
    CHECK(sleqp_tr_solver_release(&tr));
  }

  CHECK(sleqp_aug_jac_release(&jac));
  CHECK(sleqp_fact_release(&fact));
  CHECK(sleqp_iterate_release(&iterate));
  CHECK(sleqp_problem_release(&problem));
  CHECK(sleqp_settings_release(&settings));
  return 0;
}
