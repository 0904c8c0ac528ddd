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
#ifdef HARNESS_B200_AUG_JAC
#include "aug_jac/b200_aug_jac.h"
#endif
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
  int poisson; // 0: chained Rosenbrock, 2 / 3: Poisson control on a g^2 / g^3 grid
  int g;
  double alpha;
} Data;

static double
poisson_target(const Data* d, int i)
{
  double v = 1.;
  for (int a = 0; a < d->poisson; ++a, i /= d->g)
  {
    v *= sin(M_PI * (i % d->g + 1.) / (d->g + 1.));
  }
  return v;
}

// grid neighbours of node k in ascending order (Dirichlet boundary: outside nodes are dropped); returns their number
static int
poisson_neighbours(const Data* d, int k, int* out)
{
  const int g = d->g, dim = d->poisson;
  int coord[3], stride[3], cnt = 0;
  for (int a = 0, q = k, st = 1; a < dim; ++a, q /= g, st *= g)
  {
    coord[a]  = q % g;
    stride[a] = st;
  }
  for (int a = dim - 1; a >= 0; --a)
  {
    if (coord[a] > 0)
    {
      out[cnt++] = k - stride[a];
    }
  }
  for (int a = 0; a < dim; ++a)
  {
    if (coord[a] + 1 < g)
    {
      out[cnt++] = k + stride[a];
    }
  }
  return cnt;
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
      int nb[6];
      const int cnt = poisson_neighbours(d, k, nb);
      double v      = 2. * d->poisson * d->x[k] - d->x[d->m + k];
      for (int q = 0; q < cnt; ++q)
      {
        v -= d->x[nb[q]];
      }
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
  SLEQP_CALL(sleqp_mat_reserve(J, d->poisson ? (2 * d->poisson + 2) * d->m : 3 * d->m));
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
      int nb[6]; // column j of the symmetric Laplacian, rows ascending
      const int cnt = poisson_neighbours(d, j, nb);
      (void)g;
      bool diag_done = false;
      for (int t = 0; t < cnt; ++t)
      {
        if (!diag_done && nb[t] > j)
        {
          SLEQP_CALL(sleqp_mat_push(J, j, j, 2. * d->poisson));
          diag_done = true;
        }
        SLEQP_CALL(sleqp_mat_push(J, nb[t], j, -1.));
      }
      if (!diag_done)
      {
        SLEQP_CALL(sleqp_mat_push(J, j, j, 2. * d->poisson));
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

// ---- problem set-up shared by the parity and the timing mode ----------------------------------------------
typedef struct
{
  Data data;
  SleqpFunc* func;
  SleqpSettings* settings;
  SleqpProblem* problem;
  SleqpIterate* iterate;
  SleqpFact* fact;
  SleqpAugJac* jac;
  int n, m, ws_size;
} Setup;

// kind: "chain" (size = number of variables), "poisson" / "poisson3" (size = grid edge g). active_every: every how
// many-th candidate variable sits at its upper bound (chain: even variables, poisson: controls); 0 = none.
static void
setup_problem(Setup* s, const char* kind, int size, int active_every)
{
  const int poisson = kind[0] == 'p' ? (kind[7] == '3' ? 3 : 2) : 0;
  int q             = 1;
  for (int a = 0; a < poisson; ++a)
  {
    q *= size;
  }
  const int n = poisson ? 2 * q : size;
  const int m = poisson ? q : (n - 2) / 2;
  s->n        = n;
  s->m        = m;
  s->data     = (Data){n, m, (double*)calloc(n, sizeof(double)), poisson, size, 1e-2};

  SleqpFuncCallbacks callbacks = {.set_value = f_set,
                                  .obj_val   = f_obj_val,
                                  .obj_grad  = f_obj_grad,
                                  .cons_val  = f_cons_val,
                                  .cons_jac  = f_cons_jac,
                                  .hess_prod = f_hess_prod,
                                  .func_free = NULL};
  CHECK(sleqp_func_create(&s->func, &callbacks, n, m, &s->data));

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

  CHECK(sleqp_settings_create(&s->settings));
  CHECK(sleqp_problem_create_simple(&s->problem, s->func, var_lb, var_ub, cons_lb, cons_ub, s->settings));
  CHECK(sleqp_iterate_create(&s->iterate, s->problem, x0));
  CHECK(sleqp_set_and_evaluate(s->problem, s->iterate, SLEQP_VALUE_REASON_NONE, NULL));

  SleqpWorkingSet* ws = sleqp_iterate_working_set(s->iterate);
  CHECK(sleqp_working_set_reset(ws));
  int n_active_vars = 0;
  if (active_every > 0)
  {
    const int step = poisson ? active_every : 2 * active_every;
    for (int j = poisson ? m : 0; j < n; j += step)
    {
      CHECK(sleqp_working_set_add_var(ws, j, SLEQP_ACTIVE_UPPER));
      ++n_active_vars;
    }
  }
  for (int k = 0; k < m; ++k)
  {
    CHECK(sleqp_working_set_add_cons(ws, k, SLEQP_ACTIVE_BOTH));
  }
  s->ws_size = n_active_vars + m;

  CHECK(sleqp_fact_create_default(&s->fact, s->settings));
#ifdef HARNESS_B200_AUG_JAC
  // the augmented Jacobian of the B200 backend: KKT assembly on the device (the SleqpFact above is only asked its name)
  if (!getenv("HARNESS_STANDARD_AUG_JAC"))
  {
    CHECK(sleqp_b200_aug_jac_create(&s->jac, s->problem, s->settings));
    return;
  }
#endif
  CHECK(sleqp_standard_aug_jac_create(&s->jac, s->problem, s->settings, s->fact));
}

// Hessian of the Lagrangian at the start iterate as a full symmetric CSC matrix (the constraint duals are zero there)
static SleqpMat*
hessian_matrix(const Data* d)
{
  const int n = d->n;
  SleqpMat* H;
  CHECK(sleqp_mat_create(&H, n, n, d->poisson ? n : 3 * n));
  for (int j = 0; j < n; ++j)
  {
    CHECK(sleqp_mat_push_col(H, j));
    if (d->poisson)
    {
      CHECK(sleqp_mat_push(H, j, j, j < d->m ? 1. : d->alpha));
      continue;
    }
    double diag = 0.;
    if (j + 1 < n)
    {
      diag += 1200. * d->x[j] * d->x[j] - 400. * d->x[j + 1] + 2.;
    }
    if (j > 0)
    {
      diag += 200.;
      CHECK(sleqp_mat_push(H, j - 1, j, -400. * d->x[j - 1]));
    }
    CHECK(sleqp_mat_push(H, j, j, diag));
    if (j + 1 < n)
    {
      CHECK(sleqp_mat_push(H, j + 1, j, -400. * d->x[j]));
    }
  }
  return H;
}

#ifdef HARNESS_B200_TR
#include "tr/tr_b200.h"
#endif

// the trust-region solver under test: the reference's Steihaug solver, or (-DHARNESS_B200_TR) the B200 one
static void
create_tr_solver(SleqpTRSolver** tr, Setup* s, SleqpMat* hessian)
{
#ifdef HARNESS_B200_TR
  CHECK(sleqp_b200_tr_solver_create(tr, s->problem, s->settings));
  if (hessian)
  {
    CHECK(sleqp_tr_b200_set_hessian(*tr, hessian));
  }
#else
  (void)hessian;
  CHECK(sleqp_steihaug_solver_create(tr, s->problem, s->settings));
#endif
}

#if defined(HARNESS_NO_MAIN)
// included by full_solve.c for the problem definitions only
#elif !defined(HARNESS_TIMING)

int
main(int argc, char** argv)
{
  const int arg1    = argc > 1 ? atoi(argv[1]) : 100;
  const int max_it  = argc > 2 ? atoi(argv[2]) : 8;
  const char* kind  = argc > 3 ? argv[3] : "chain";
  const bool matrix = argc > 4 && argv[4][0] == 'm'; // B200 TR solver: Hessian as a device matrix instead of the callback
  Setup S;
  setup_problem(&S, kind, arg1, kind[0] == 'p' ? 7 : 10);
  const int n = S.n, ws_size = S.ws_size;
  SleqpSettings* settings = S.settings;
  SleqpProblem* problem   = S.problem;
  SleqpIterate* iterate   = S.iterate;
  SleqpFact* fact         = S.fact;
  SleqpAugJac* jac        = S.jac;
  (void)problem;

  printf("backend 1 %d\n", (int)sleqp_fact_flags(fact));
  fprintf(stderr, "backend: %s\n", sleqp_fact_name(fact));
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
  SleqpMat* hessian = matrix ? hessian_matrix(&S.data) : NULL;
  double full_norm  = 0.;
  for (int it = -1; it < max_it; ++it)
  {
    // it = -1: unconstrained by the radius (the converged Newton step); then radii inside its norm
    const double radius = it < 0 ? 1e8 : full_norm * (0.6 + 0.4 * (it + 0.5) / max_it);
    SleqpTRSolver* tr;
    create_tr_solver(&tr, &S, hessian);
    SleqpVec* step;
    CHECK(sleqp_vec_create_empty(&step, n));
    double tr_dual = 0.;
    CHECK(sleqp_tr_solver_solve(tr, jac, cons_dual, grad, step, radius, &tr_dual));
    double ray_min = 0., ray_max = 0.;
    CHECK(sleqp_tr_solver_current_rayleigh(tr, &ray_min, &ray_max));
    char name[64];
    if (it < 0)
    {
      full_norm = sleqp_vec_norm(step);
      snprintf(name, sizeof(name), "cg_converged_step");
      printf("tr_info_converged 3 %.17g %.17g %.17g\n", tr_dual, ray_min, ray_max);
    }
    else
    {
      snprintf(name, sizeof(name), "cg_path_sample_%d", it);
      // dual of the trust region (boundary exit) and the Rayleigh bounds of the directions so far
      printf("tr_info_%d 3 %.17g %.17g %.17g\n", it, tr_dual, ray_min, ray_max);
    }
    dump(name, step);
    CHECK(sleqp_vec_free(&step));
    CHECK(sleqp_tr_solver_release(&tr));
  }

  CHECK(sleqp_aug_jac_release(&jac));
  CHECK(sleqp_fact_release(&fact));
  CHECK(sleqp_iterate_release(&iterate));
  CHECK(sleqp_problem_release(&S.problem));
  CHECK(sleqp_settings_release(&settings));
  return 0;
}

#else // HARNESS_TIMING: the EQP part of one SQP iteration, timed end to end through reference code

#include <time.h>

#include "sparse/mat.h"

#ifdef HARNESS_B200_TR
#include "sparse/mat_b200.h"
#endif

static double
now_ms(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return 1e3 * t.tv_sec + 1e-6 * t.tv_nsec;
}

/*
 * usage: eqp_step_b200 <chain|poisson|poisson3> <size> <cg_iters> <steps> [steihaug] [hostspmv]
 *
 * One step = what SLEQP does between the Cauchy step and the trial point, every call a reference
 * function on host data:
 *   sleqp_aug_jac_set_iterate    host fill_aug_jac (standard_aug_jac.c:135-237) + SleqpFact.set_matrix + condition
 *   sleqp_aug_jac_solve_min_norm, sleqp_aug_jac_solve_lsq
 *   J^T v (sparse multipliers, newton.c:377) and J x (direction.c:66) through the shipped SpMV glue sparse/mat_b200.c
 *   (mirror refreshed once per step, like the Jacobian of a new iterate), or with "hostspmv" the reference's own
 *   sleqp_mat_mult_vec_trans / sleqp_mat_mult_vec -- whose transposed product is quadratic (mat.c:329-331)
 *   sleqp_tr_solver_solve capped at <cg_iters> iterations: the B200 TR solver (device CG, Hessian as a matrix), or with
 *   the 5th argument the reference's own Steihaug solver over the B200 factorization (one boundary crossing per iteration)
 * Prints one JSON line.
 */
int
main(int argc, char** argv)
{
  const char* kind   = argc > 1 ? argv[1] : "chain";
  const int size     = argc > 2 ? atoi(argv[2]) : 100000;
  const int cg_iters = argc > 3 ? atoi(argv[3]) : 30;
  const int steps    = argc > 4 ? atoi(argv[4]) : 5;
  const bool steihaug = argc > 5 && argv[5][0] == 's';
  const bool hostspmv = (argc > 5 && argv[5][0] == 'h') || (argc > 6 && argv[6][0] == 'h');
  Setup S;
  // working sets like sleqp_b200/problems.py: chain = every constraint, poisson = every constraint + 10 % of the controls
  setup_problem(&S, kind, size, kind[0] == 'p' ? 10 : 0);
  const int n = S.n, m = S.m, ws_size = S.ws_size;
  CHECK(sleqp_settings_set_int_value(S.settings, SLEQP_SETTINGS_INT_MAX_NEWTON_ITERATIONS, cg_iters));
  CHECK(sleqp_settings_set_real_value(S.settings, SLEQP_SETTINGS_REAL_STAT_TOL, 1e-12)); // run to the cap

  SleqpVec* grad      = sleqp_iterate_obj_grad(S.iterate);
  SleqpVec* cons_dual = sleqp_iterate_cons_dual(S.iterate);
  SleqpMat* J         = sleqp_iterate_cons_jac(S.iterate);
  SleqpVec *minnorm_rhs, *minnorm, *duals, *step, *viol, *jtv, *xfull;
  CHECK(sleqp_vec_create_full(&minnorm_rhs, ws_size));
  CHECK(sleqp_vec_create_empty(&minnorm, n));
  CHECK(sleqp_vec_create_empty(&duals, ws_size));
  CHECK(sleqp_vec_create_empty(&step, n));
  CHECK(sleqp_vec_create_empty(&viol, m));
  CHECK(sleqp_vec_create_empty(&jtv, n));
  CHECK(sleqp_vec_create_full(&xfull, n));
  for (int i = 0; i < ws_size; ++i)
  {
    CHECK(sleqp_vec_push(minnorm_rhs, i, sin(0.37 * i) + 0.1));
  }
  CHECK(sleqp_vec_reserve(viol, m / 100 + 1));
  for (int k = 0; k < m; k += 100)
  {
    CHECK(sleqp_vec_push(viol, k, cos(0.11 * k)));
  }
  for (int i = 0; i < n; ++i)
  {
    CHECK(sleqp_vec_push(xfull, i, sin(0.05 * i) + 0.2));
  }
  double* jx = (double*)calloc(m > 0 ? m : 1, sizeof(double));

  SleqpMat* hessian = steihaug ? NULL : hessian_matrix(&S.data);
  SleqpTRSolver* tr;
#ifdef HARNESS_B200_TR
  if (steihaug)
  {
    CHECK(sleqp_steihaug_solver_create(&tr, S.problem, S.settings));
  }
  else
  {
    create_tr_solver(&tr, &S, hessian);
  }
#else
  create_tr_solver(&tr, &S, hessian);
#endif

#ifdef HARNESS_B200_TR
  SleqpMatB200* mirror = NULL;
  if (!hostspmv)
  {
    CHECK(sleqp_mat_b200_create(&mirror));
  }
#endif

  double t_set = 0., t_solves = 0., t_spmv = 0., t_tr = 0., t_total = 0.;
  double tr_dual = 0.;
  const int warm = 2;
  for (int it = -warm; it < steps; ++it)
  {
    const double t0 = now_ms();
    CHECK(sleqp_aug_jac_set_iterate(S.jac, S.iterate));
    const double t1 = now_ms();
    CHECK(sleqp_aug_jac_solve_min_norm(S.jac, minnorm_rhs, minnorm));
    CHECK(sleqp_aug_jac_solve_lsq(S.jac, grad, duals));
    const double t2 = now_ms();
#ifdef HARNESS_B200_TR
    if (mirror)
    {
      CHECK(sleqp_mat_b200_update(mirror, J)); // the Jacobian of this iterate goes to the device once
      CHECK(sleqp_mat_b200_mult_vec_trans(mirror, viol, 0., jtv));
      CHECK(sleqp_mat_b200_mult_vec(mirror, xfull, jx));
    }
    else
#endif
    {
      CHECK(sleqp_mat_mult_vec_trans(J, viol, 0., jtv));
      CHECK(sleqp_mat_mult_vec(J, xfull, jx));
    }
    const double t3 = now_ms();
    CHECK(sleqp_tr_solver_solve(tr, S.jac, cons_dual, grad, step, 1e8, &tr_dual));
    const double t4 = now_ms();
    if (it >= 0)
    {
      t_set += t1 - t0;
      t_solves += t2 - t1;
      t_spmv += t3 - t2;
      t_tr += t4 - t3;
      t_total += t4 - t0;
    }
  }
  int cg_done = cg_iters, cg_exit = -1;
#ifdef HARNESS_B200_TR
  if (!steihaug)
  {
    CHECK(sleqp_tr_b200_last_solve(tr, &cg_done, &cg_exit));
  }
#endif
  const SleqpMat* K = NULL;
  (void)K;
  const double ms = t_total / steps;
  // bytes over PCIe per step (counted from the vectors that cross): K / Jacobian values in; per aug_jac solve a sparse
  // right-hand side in (8 B values; the indices of a contiguous run stay on the host) and a sparsified slice out (12 B
  // per entry, copied at full length); the TR solve: gradient in (8 B, contiguous), sparse step out (12 B) -- or, with
  // the reference's Steihaug loop, one projection (12 n in, 12 n out) per iteration; the device mirror of the Jacobian:
  // its values in, x in (8 B), the multipliers in (12 B, sparse), J^T v out (12 B, sparsified), J x out (8 B)
  const double nnzK = (double)n + 3. * m; // order of magnitude only for chain; exact value is printed by bench.py's C-ABI leg
  (void)nnzK;
  const long long h2d = 8LL * sleqp_mat_nnz(J) + 8LL * n /* diag */ + 8LL * ws_size + 8LL * n
                        + (steihaug ? 12LL * n * (cg_iters + 1) : 8LL * n + (hessian ? 0 : 8LL * n * cg_iters));
  const long long d2h = 12LL * n + 12LL * ws_size + (steihaug ? 12LL * n * (cg_iters + 1) : 12LL * n + (hessian ? 0 : 8LL * n * cg_iters));
  printf("{\"driver\": \"eqp_step (reference aug_jac/TR code over the B200 glue)\", \"problem\": \"%s\", \"size\": %d, \"n\": %d, \"m\": %d, "
         "\"ws_size\": %d, \"N\": %d, \"backend\": \"%s\", \"aug_jac\": \"%s\", \"tr_solver\": \"%s\", \"cg_iters_cap\": %d, \"cg_iterations\": %d, \"cg_exit\": %d, "
         "\"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.6f, \"iters_per_s\": %.6f, \"set_iterate_ms\": %.6f, \"two_solves_ms\": %.6f, "
         "\"solve_ms\": %.6f, \"spmv\": \"%s\", \"jac_products_ms\": %.6f, \"tr_solve_ms\": %.6f, \"h2d_bytes_per_step\": %lld, \"d2h_bytes_per_step\": %lld}\n",
         kind, size, n, m, ws_size, n + ws_size, sleqp_fact_name(S.fact),
#ifdef HARNESS_B200_AUG_JAC
         getenv("HARNESS_STANDARD_AUG_JAC") ? "reference standard_aug_jac.c (host fill_aug_jac)" : "aug_jac/b200_aug_jac.c (KKT assembled on the device)",
#else
         "reference standard_aug_jac.c (host fill_aug_jac)",
#endif
         steihaug ? "reference steihaug_solver.c over SleqpFact B200" : "tr_b200.c (device CG, Hessian as a device matrix)", cg_iters, cg_done, cg_exit,
         steps, warm, ms, 1e3 / ms, t_set / steps, t_solves / steps, t_solves / steps / 2.,
         hostspmv ? "reference sleqp_mat_mult_vec(_trans) on the host" : "sparse/mat_b200.c (device mirror, refreshed every step)", t_spmv / steps, t_tr / steps,
         h2d + (hostspmv ? 0 : 8LL * sleqp_mat_nnz(J) + 8LL * n + 12LL * (m / 100 + 1)), d2h + (hostspmv ? 0 : 12LL * n + 8LL * m));
  return 0;
}

#endif
