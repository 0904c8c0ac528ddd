/*
 * lpi_simplex.c -- SLEQP LP backend "Simplex": the sixteen SleqpLPiCallbacks
 * (src/main/lp/lpi_types.h:34-125) over a self-contained dense bounded-variable primal
 * simplex method.
 *
 * none of the reference's LP backends (HiGHS, Gurobi, SoPlex;
 * lp/lpi_highs.c, lpi_gurobi.c, lpi_soplex.cc) exists in this environment, so a full
 * sleqp_solver_solve could not be linked. This file is the missing piece: small, host-only,
 * made for the LPs SLEQP actually poses (standard_cauchy.c:155-190: num_vars + 2 num_cons
 * columns, num_cons rows, every column bounded below, a feasible slack basis handed over by
 * create_and_set_slack_basis, :70-131). It is meant for problems of up to a few thousand
 * columns -- config 1 of BASELINE.json and the reference's own test problems -- not for the
 * large configurations, where the KKT path is benchmarked without an LP.
 *
 * Drop-in like its siblings: defines sleqp_lpi_create_default (lpi_highs.c:744-755). C11, no
 * dependencies beyond the reference's own headers.
 *
 * Method. Rows are turned into equations A x - r = 0 with "logical" variables r bounded by the
 * row bounds, so every variable (structural or logical) is either basic or sits at a bound
 * (SLEQP_BASESTAT_LOWER / UPPER, ZERO for free non-basic ones, lpi_types.h:16-22). Primal
 * simplex with an explicit dense inverse of the basis (product-form updates, refactorised
 * regularly), Dantzig pricing with a switch to Bland's rule after a run of degenerate
 * pivots, bound flips, and a composite phase 1 that minimises the sum of the infeasibilities
 * of the basic variables when a warm-start basis is not primal feasible.
 * Duals follow the convention of lpi_highs.c:583-612: vars_dual = reduced costs c - A^T y,
 * cons_dual = y, non-negative for a row at its lower bound (minimisation).
 */
#include "lpi_simplex.h"

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "cmp.h"
#include "defs.h"
#include "error.h"
#include "log.h"
#include "mem.h"

#define SIMPLEX_NAME "Simplex"
#define SIMPLEX_VERSION "0.1"

static const double feas_tol   = 1e-9;
static const double opt_tol    = 1e-9;
static const double pivot_tol  = 1e-9;
static const int refactor_freq = 40;

typedef struct
{
  int num_cols;
  int num_rows;

  // problem data: A column-major (num_rows x num_cols), bounds of the structural
  // variables followed by those of the logical ones (row activities)
  double* A;
  double* cost;
  double* lb;
  double* ub;

  // basis
  SLEQP_BASESTAT* stat; // num_cols + num_rows
  int* basic;           // num_rows: variable in every basis position
  int* position;        // num_cols + num_rows: basis position or -1
  double* Binv;         // num_rows x num_rows, row-major
  double* x;            // values of all variables
  bool basis_valid;

  // saved bases
  int num_bases;
  SLEQP_BASESTAT** saved;

  // results
  SLEQP_LP_STATUS status;
  double* y;      // row duals
  double* dj;     // reduced costs of the structural variables
  double objective;
  int iterations;

  // work
  double* work;   // num_rows
  double* column; // num_rows
  double* cB;     // num_rows
  double* scratch; // num_rows x num_rows (refactorisation)
} Simplex;

static inline bool
is_neg_inf(double v)
{
  return sleqp_is_infinite(-v);
}

static inline bool
is_pos_inf(double v)
{
  return sleqp_is_infinite(v);
}

// column of variable j of [A, -I] into out
static void
get_column(const Simplex* lp, int j, double* out)
{
  const int m = lp->num_rows;

  if (j < lp->num_cols)
  {
    memcpy(out, lp->A + (size_t)j * m, sizeof(double) * m);
  }
  else
  {
    memset(out, 0, sizeof(double) * m);
    out[j - lp->num_cols] = -1.;
  }
}

// a status that is consistent with the bounds of the variable (SLEQP may hand over LOWER
// for a variable whose lower bound has become infinite in the meantime)
static SLEQP_BASESTAT
sanitize(const Simplex* lp, int j, SLEQP_BASESTAT stat)
{
  if (stat == SLEQP_BASESTAT_BASIC)
  {
    return stat;
  }

  const bool has_lb = !is_neg_inf(lp->lb[j]);
  const bool has_ub = !is_pos_inf(lp->ub[j]);

  if (stat == SLEQP_BASESTAT_LOWER && has_lb)
  {
    return stat;
  }
  if (stat == SLEQP_BASESTAT_UPPER && has_ub)
  {
    return stat;
  }
  if (has_lb)
  {
    return SLEQP_BASESTAT_LOWER;
  }
  if (has_ub)
  {
    return SLEQP_BASESTAT_UPPER;
  }
  return SLEQP_BASESTAT_ZERO;
}

static double
nonbasic_value(const Simplex* lp, int j)
{
  switch (lp->stat[j])
  {
  case SLEQP_BASESTAT_LOWER:
    return lp->lb[j];
  case SLEQP_BASESTAT_UPPER:
    return lp->ub[j];
  default:
    return 0.;
  }
}

static void
slack_basis(Simplex* lp)
{
  const int n = lp->num_cols, m = lp->num_rows;

  for (int j = 0; j < n; ++j)
  {
    lp->stat[j] = SLEQP_BASESTAT_LOWER;
  }
  for (int i = 0; i < m; ++i)
  {
    lp->stat[n + i] = SLEQP_BASESTAT_BASIC;
  }
}

// Inverse of the basis matrix by Gauss-Jordan elimination with partial pivoting. Returns
// false if the basis is (numerically) singular.
static bool
invert_basis(Simplex* lp)
{
  const int m = lp->num_rows;
  double* B   = lp->scratch;
  double* inv = lp->Binv;

  for (int p = 0; p < m; ++p)
  {
    get_column(lp, lp->basic[p], lp->column);
    for (int i = 0; i < m; ++i)
    {
      B[(size_t)i * m + p] = lp->column[i];
    }
  }

  memset(inv, 0, sizeof(double) * (size_t)m * m);
  for (int i = 0; i < m; ++i)
  {
    inv[(size_t)i * m + i] = 1.;
  }

  for (int c = 0; c < m; ++c)
  {
    int piv     = c;
    double best = fabs(B[(size_t)c * m + c]);
    for (int i = c + 1; i < m; ++i)
    {
      const double v = fabs(B[(size_t)i * m + c]);
      if (v > best)
      {
        best = v;
        piv  = i;
      }
    }
    if (best < 1e-12)
    {
      return false;
    }
    if (piv != c)
    {
      for (int q = 0; q < m; ++q)
      {
        double t               = B[(size_t)c * m + q];
        B[(size_t)c * m + q]   = B[(size_t)piv * m + q];
        B[(size_t)piv * m + q] = t;
        t                      = inv[(size_t)c * m + q];
        inv[(size_t)c * m + q]   = inv[(size_t)piv * m + q];
        inv[(size_t)piv * m + q] = t;
      }
    }
    const double d = 1. / B[(size_t)c * m + c];
    for (int q = 0; q < m; ++q)
    {
      B[(size_t)c * m + q] *= d;
      inv[(size_t)c * m + q] *= d;
    }
    for (int i = 0; i < m; ++i)
    {
      if (i == c)
      {
        continue;
      }
      const double f = B[(size_t)i * m + c];
      if (f == 0.)
      {
        continue;
      }
      for (int q = 0; q < m; ++q)
      {
        B[(size_t)i * m + q] -= f * B[(size_t)c * m + q];
        inv[(size_t)i * m + q] -= f * inv[(size_t)c * m + q];
      }
    }
  }

  return true;
}

// basic list / positions from the statuses; repairs a basis with the wrong number of basic
// variables or a singular one by falling back to the slack basis
static void
install_basis(Simplex* lp)
{
  const int n = lp->num_cols, m = lp->num_rows, N = n + m;

  int count = 0;
  for (int j = 0; j < N; ++j)
  {
    lp->stat[j] = sanitize(lp, j, lp->stat[j]);
    count += lp->stat[j] == SLEQP_BASESTAT_BASIC;
  }

  for (int attempt = 0; attempt < 2; ++attempt)
  {
    if (count != m || attempt == 1)
    {
      slack_basis(lp);
      for (int j = 0; j < N; ++j)
      {
        lp->stat[j] = sanitize(lp, j, lp->stat[j]);
      }
    }

    int p = 0;
    for (int j = 0; j < N; ++j)
    {
      lp->position[j] = -1;
      if (lp->stat[j] == SLEQP_BASESTAT_BASIC)
      {
        lp->position[j] = p;
        lp->basic[p++]  = j;
      }
    }

    if (m == 0 || invert_basis(lp))
    {
      break;
    }
    count = -1; // singular: second attempt with the slack basis
  }

  lp->basis_valid = true;
}

// x_N from the statuses, x_B = -Binv N x_N
static void
compute_primal(Simplex* lp)
{
  const int n = lp->num_cols, m = lp->num_rows, N = n + m;
  double* rhs = lp->work;

  memset(rhs, 0, sizeof(double) * m);

  for (int j = 0; j < N; ++j)
  {
    if (lp->stat[j] == SLEQP_BASESTAT_BASIC)
    {
      continue;
    }
    const double v = nonbasic_value(lp, j);
    lp->x[j]       = v;
    if (v == 0.)
    {
      continue;
    }
    if (j < n)
    {
      const double* col = lp->A + (size_t)j * m;
      for (int i = 0; i < m; ++i)
      {
        rhs[i] -= col[i] * v;
      }
    }
    else
    {
      rhs[j - n] += v;
    }
  }

  for (int p = 0; p < m; ++p)
  {
    const double* row = lp->Binv + (size_t)p * m;
    double s          = 0.;
    for (int i = 0; i < m; ++i)
    {
      s += row[i] * rhs[i];
    }
    lp->x[lp->basic[p]] = s;
  }
}

static SLEQP_RETCODE
simplex_solve(void* lp_data, int num_cols, int num_rows, double time_limit)
{
  Simplex* lp = (Simplex*)lp_data;

  (void)time_limit;

  const int n = num_cols, m = num_rows, N = n + m;

  assert(n == lp->num_cols && m == lp->num_rows);

  lp->status = SLEQP_LP_STATUS_UNKNOWN;

  for (int j = 0; j < N; ++j)
  {
    if (lp->lb[j] > lp->ub[j] + feas_tol)
    {
      lp->status = SLEQP_LP_STATUS_INF;
      return SLEQP_OKAY;
    }
  }

  install_basis(lp);
  compute_primal(lp);

  const int max_iter = 50 * (N + 10);
  int degenerate_run = 0;
  int since_refactor = 0;

  for (int iter = 0;; ++iter)
  {
    if (iter >= max_iter)
    {
      sleqp_raise(SLEQP_INTERNAL_ERROR,
                  "Simplex LP backend: iteration limit (%d) reached",
                  max_iter);
    }

    // phase: costs of the basic variables
    bool phase1 = false;
    for (int p = 0; p < m; ++p)
    {
      const int j = lp->basic[p];
      if (lp->x[j] < lp->lb[j] - feas_tol)
      {
        lp->cB[p] = -1.;
        phase1    = true;
      }
      else if (lp->x[j] > lp->ub[j] + feas_tol)
      {
        lp->cB[p] = 1.;
        phase1    = true;
      }
      else
      {
        lp->cB[p] = 0.;
      }
    }
    if (!phase1)
    {
      for (int p = 0; p < m; ++p)
      {
        const int j = lp->basic[p];
        lp->cB[p]   = j < n ? lp->cost[j] : 0.;
      }
    }

    // y = Binv^T cB
    for (int i = 0; i < m; ++i)
    {
      lp->y[i] = 0.;
    }
    for (int p = 0; p < m; ++p)
    {
      const double c = lp->cB[p];
      if (c == 0.)
      {
        continue;
      }
      const double* row = lp->Binv + (size_t)p * m;
      for (int i = 0; i < m; ++i)
      {
        lp->y[i] += c * row[i];
      }
    }

    // pricing
    const bool bland = degenerate_run > 2 * (m + 5);
    int enter        = -1;
    double best      = 0.;
    int direction    = 0;
    for (int j = 0; j < N; ++j)
    {
      if (lp->stat[j] == SLEQP_BASESTAT_BASIC || lp->lb[j] == lp->ub[j])
      {
        continue;
      }
      double d;
      if (j < n)
      {
        const double* col = lp->A + (size_t)j * m;
        double s          = 0.;
        for (int i = 0; i < m; ++i)
        {
          s += col[i] * lp->y[i];
        }
        d = (phase1 ? 0. : lp->cost[j]) - s;
      }
      else
      {
        d = lp->y[j - n];
      }
      int dir = 0;
      if (d < -opt_tol && lp->stat[j] != SLEQP_BASESTAT_UPPER)
      {
        dir = 1;
      }
      else if (d > opt_tol && lp->stat[j] != SLEQP_BASESTAT_LOWER)
      {
        dir = -1;
      }
      if (!dir)
      {
        continue;
      }
      if (bland)
      {
        enter     = j;
        direction = dir;
        break;
      }
      if (fabs(d) > best)
      {
        best      = fabs(d);
        enter     = j;
        direction = dir;
      }
    }

    if (enter == -1)
    {
      if (phase1)
      {
        lp->status = SLEQP_LP_STATUS_INF;
        return SLEQP_OKAY;
      }
      break; // optimal
    }

    // w = Binv a_enter
    get_column(lp, enter, lp->column);
    double* w = lp->work;
    for (int p = 0; p < m; ++p)
    {
      const double* row = lp->Binv + (size_t)p * m;
      double s          = 0.;
      for (int i = 0; i < m; ++i)
      {
        s += row[i] * lp->column[i];
      }
      w[p] = s;
    }

    // ratio test: x_enter moves by t * direction, x_B by -t * direction * w
    double t_max = lp->ub[enter] - lp->lb[enter]; // bound flip (inf for a free or half-bounded variable)
    if (!(t_max < sleqp_infinity()))
    {
      t_max = INFINITY;
    }
    int leave_pos   = -1;
    bool leave_upper = false;
    double leave_piv = 0.;
    for (int p = 0; p < m; ++p)
    {
      const double delta = -direction * w[p]; // change of the basic variable per unit step
      if (fabs(delta) <= pivot_tol)
      {
        continue;
      }
      const int j    = lp->basic[p];
      const double v = lp->x[j];
      double t       = INFINITY;
      bool to_upper  = false;
      if (delta > 0.)
      {
        // increases: blocks at its upper bound, or -- infeasible below -- when it reaches the lower one
        if (v < lp->lb[j] - feas_tol)
        {
          t = (lp->lb[j] - v) / delta;
        }
        else if (!is_pos_inf(lp->ub[j]) && v <= lp->ub[j] + feas_tol)
        {
          t        = fmax(0., lp->ub[j] - v) / delta;
          to_upper = true;
        }
      }
      else
      {
        if (v > lp->ub[j] + feas_tol)
        {
          t        = (lp->ub[j] - v) / delta;
          to_upper = true;
        }
        else if (!is_neg_inf(lp->lb[j]) && v >= lp->lb[j] - feas_tol)
        {
          t = fmax(0., v - lp->lb[j]) / (-delta);
        }
      }
      // ties: the larger pivot element (stability); under Bland's rule the smallest variable index
      if (t < t_max - 1e-12
          || (leave_pos != -1 && fabs(t - t_max) <= 1e-12
              && (bland ? j < lp->basic[leave_pos] : fabs(delta) > leave_piv)))
      {
        t_max       = t;
        leave_pos   = p;
        leave_upper = to_upper;
        leave_piv   = fabs(delta);
      }
    }

    if (!(t_max < INFINITY))
    {
      if (phase1)
      {
        // cannot happen for a sum of infeasibilities (bounded below): numerical trouble
        sleqp_raise(SLEQP_INTERNAL_ERROR,
                    "Simplex LP backend: unbounded phase 1");
      }
      lp->status = SLEQP_LP_STATUS_UNBOUNDED;
      return SLEQP_OKAY;
    }

    degenerate_run = t_max <= 1e-12 ? degenerate_run + 1 : 0;

    // move
    const double step = t_max * direction;
    for (int p = 0; p < m; ++p)
    {
      lp->x[lp->basic[p]] -= step * w[p];
    }
    lp->x[enter] += step;

    if (leave_pos == -1)
    {
      // bound flip of the entering variable
      lp->stat[enter] = direction > 0 ? SLEQP_BASESTAT_UPPER : SLEQP_BASESTAT_LOWER;
      lp->x[enter]    = nonbasic_value(lp, enter);
      continue;
    }

    // pivot: the leaving variable goes to the bound it has reached
    const int leave = lp->basic[leave_pos];
    lp->stat[leave] = leave_upper ? SLEQP_BASESTAT_UPPER : SLEQP_BASESTAT_LOWER;
    lp->stat[leave] = sanitize(lp, leave, lp->stat[leave]);
    lp->x[leave]    = nonbasic_value(lp, leave);
    lp->position[leave] = -1;

    lp->stat[enter]      = SLEQP_BASESTAT_BASIC;
    lp->basic[leave_pos] = enter;
    lp->position[enter]  = leave_pos;

    if (++since_refactor >= refactor_freq)
    {
      since_refactor = 0;
      if (!invert_basis(lp))
      {
        sleqp_raise(SLEQP_INTERNAL_ERROR,
                    "Simplex LP backend: singular basis after a pivot");
      }
      compute_primal(lp);
    }
    else
    {
      // product-form update of the inverse: row leave_pos scaled by 1 / w, eliminated from the others
      double* prow       = lp->Binv + (size_t)leave_pos * m;
      const double pivot = w[leave_pos];
      for (int i = 0; i < m; ++i)
      {
        prow[i] /= pivot;
      }
      for (int p = 0; p < m; ++p)
      {
        if (p == leave_pos || w[p] == 0.)
        {
          continue;
        }
        double* row    = lp->Binv + (size_t)p * m;
        const double f = w[p];
        for (int i = 0; i < m; ++i)
        {
          row[i] -= f * prow[i];
        }
      }
    }

    lp->iterations = iter + 1;
  }

  // optimal: clean values from a fresh factorisation, duals and reduced costs
  if (m > 0 && !invert_basis(lp))
  {
    sleqp_raise(SLEQP_INTERNAL_ERROR,
                "Simplex LP backend: singular optimal basis");
  }
  compute_primal(lp);

  for (int p = 0; p < m; ++p)
  {
    const int j = lp->basic[p];
    lp->cB[p]   = j < n ? lp->cost[j] : 0.;
  }
  for (int i = 0; i < m; ++i)
  {
    lp->y[i] = 0.;
  }
  for (int p = 0; p < m; ++p)
  {
    const double c = lp->cB[p];
    if (c == 0.)
    {
      continue;
    }
    const double* row = lp->Binv + (size_t)p * m;
    for (int i = 0; i < m; ++i)
    {
      lp->y[i] += c * row[i];
    }
  }

  lp->objective = 0.;
  for (int j = 0; j < n; ++j)
  {
    const double* col = lp->A + (size_t)j * m;
    double s          = 0.;
    for (int i = 0; i < m; ++i)
    {
      s += col[i] * lp->y[i];
    }
    lp->dj[j] = lp->stat[j] == SLEQP_BASESTAT_BASIC ? 0. : lp->cost[j] - s;
    lp->objective += lp->cost[j] * lp->x[j];
  }

  lp->status = SLEQP_LP_STATUS_OPTIMAL;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_create_problem(void** star,
                       int num_cols,
                       int num_rows,
                       SleqpSettings* settings)
{
  (void)settings;

  Simplex* lp = NULL;

  SLEQP_CALL(sleqp_malloc(&lp));

  *lp = (Simplex){0};

  *star = lp;

  lp->num_cols = num_cols;
  lp->num_rows = num_rows;

  const size_t N = (size_t)num_cols + num_rows;
  const size_t m = num_rows > 0 ? num_rows : 1;

  if ((double)num_cols * (double)num_rows > 5e7 || (double)m * (double)m > 5e7)
  {
    sleqp_raise(SLEQP_INTERNAL_ERROR,
                "Simplex LP backend: %d x %d is beyond this dense implementation",
                num_rows,
                num_cols);
  }

  SLEQP_CALL(sleqp_alloc_array(&lp->A, (size_t)num_cols * m + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->cost, N + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->lb, N + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->ub, N + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->stat, N + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->basic, m));
  SLEQP_CALL(sleqp_alloc_array(&lp->position, N + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->Binv, m * m));
  SLEQP_CALL(sleqp_alloc_array(&lp->scratch, m * m));
  SLEQP_CALL(sleqp_alloc_array(&lp->x, N + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->y, m));
  SLEQP_CALL(sleqp_alloc_array(&lp->dj, (size_t)num_cols + 1));
  SLEQP_CALL(sleqp_alloc_array(&lp->work, m));
  SLEQP_CALL(sleqp_alloc_array(&lp->column, m));
  SLEQP_CALL(sleqp_alloc_array(&lp->cB, m));

  memset(lp->A, 0, sizeof(double) * ((size_t)num_cols * m + 1));
  memset(lp->cost, 0, sizeof(double) * (N + 1));

  const double inf = sleqp_infinity();

  for (size_t j = 0; j < N; ++j)
  {
    lp->lb[j] = j < (size_t)num_cols ? 0. : -inf;
    lp->ub[j] = inf;
    lp->x[j]  = 0.;
  }

  slack_basis(lp);

  lp->status = SLEQP_LP_STATUS_UNKNOWN;

  return SLEQP_OKAY;
}

static SLEQP_LP_STATUS
simplex_status(void* lp_data)
{
  return ((Simplex*)lp_data)->status;
}

static SLEQP_RETCODE
simplex_set_bounds(void* lp_data,
                   int num_cols,
                   int num_rows,
                   double* cons_lb,
                   double* cons_ub,
                   double* vars_lb,
                   double* vars_ub)
{
  Simplex* lp = (Simplex*)lp_data;

  for (int j = 0; j < num_cols; ++j)
  {
    lp->lb[j] = vars_lb[j];
    lp->ub[j] = vars_ub[j];
  }
  for (int i = 0; i < num_rows; ++i)
  {
    lp->lb[num_cols + i] = cons_lb[i];
    lp->ub[num_cols + i] = cons_ub[i];
  }

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_set_coeffs(void* lp_data,
                   int num_cols,
                   int num_rows,
                   SleqpMat* coeff_matrix)
{
  Simplex* lp = (Simplex*)lp_data;

  assert(sleqp_mat_num_rows(coeff_matrix) == num_rows);
  assert(sleqp_mat_num_cols(coeff_matrix) == num_cols);

  const int* cols    = sleqp_mat_cols(coeff_matrix);
  const int* rows    = sleqp_mat_rows(coeff_matrix);
  const double* data = sleqp_mat_data(coeff_matrix);

  memset(lp->A, 0, sizeof(double) * (size_t)num_cols * num_rows);

  for (int j = 0; j < num_cols; ++j)
  {
    for (int k = cols[j]; k < cols[j + 1]; ++k)
    {
      lp->A[(size_t)j * num_rows + rows[k]] = data[k];
    }
  }

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_set_objective(void* lp_data,
                      int num_cols,
                      int num_rows,
                      double* objective)
{
  Simplex* lp = (Simplex*)lp_data;

  (void)num_rows;

  for (int j = 0; j < num_cols; ++j)
  {
    lp->cost[j] = objective[j];
  }

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
reserve_bases(Simplex* lp, int size)
{
  if (size <= lp->num_bases)
  {
    return SLEQP_OKAY;
  }

  SLEQP_CALL(sleqp_realloc(&lp->saved, size));

  const int N = lp->num_cols + lp->num_rows;

  for (int b = lp->num_bases; b < size; ++b)
  {
    SLEQP_CALL(sleqp_alloc_array(&lp->saved[b], N + 1));
    for (int j = 0; j < N; ++j)
    {
      lp->saved[b][j] = j < lp->num_cols ? SLEQP_BASESTAT_LOWER : SLEQP_BASESTAT_BASIC;
    }
  }

  lp->num_bases = size;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_set_basis(void* lp_data,
                  int index,
                  const SLEQP_BASESTAT* col_stats,
                  const SLEQP_BASESTAT* row_stats)
{
  Simplex* lp = (Simplex*)lp_data;

  SLEQP_CALL(reserve_bases(lp, index + 1));

  memcpy(lp->saved[index], col_stats, sizeof(SLEQP_BASESTAT) * lp->num_cols);
  memcpy(lp->saved[index] + lp->num_cols,
         row_stats,
         sizeof(SLEQP_BASESTAT) * lp->num_rows);

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_save_basis(void* lp_data, int index)
{
  Simplex* lp = (Simplex*)lp_data;

  SLEQP_CALL(reserve_bases(lp, index + 1));

  memcpy(lp->saved[index],
         lp->stat,
         sizeof(SLEQP_BASESTAT) * (lp->num_cols + lp->num_rows));

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_restore_basis(void* lp_data, int index)
{
  Simplex* lp = (Simplex*)lp_data;

  assert(index >= 0 && index < lp->num_bases);

  memcpy(lp->stat,
         lp->saved[index],
         sizeof(SLEQP_BASESTAT) * (lp->num_cols + lp->num_rows));

  lp->basis_valid = false;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_primal_sol(void* lp_data,
                   int num_cols,
                   int num_rows,
                   double* objective_value,
                   double* solution_values)
{
  Simplex* lp = (Simplex*)lp_data;

  (void)num_rows;

  if (objective_value)
  {
    *objective_value = lp->objective;
  }

  if (solution_values)
  {
    memcpy(solution_values, lp->x, sizeof(double) * num_cols);
  }

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_dual_sol(void* lp_data,
                 int num_cols,
                 int num_rows,
                 double* vars_dual,
                 double* cons_dual)
{
  Simplex* lp = (Simplex*)lp_data;

  if (vars_dual)
  {
    memcpy(vars_dual, lp->dj, sizeof(double) * num_cols);
  }

  if (cons_dual)
  {
    memcpy(cons_dual, lp->y, sizeof(double) * num_rows);
  }

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_vars_stats(void* lp_data,
                   int num_cols,
                   int num_rows,
                   SLEQP_BASESTAT* variable_stats)
{
  Simplex* lp = (Simplex*)lp_data;

  (void)num_rows;

  memcpy(variable_stats, lp->stat, sizeof(SLEQP_BASESTAT) * num_cols);

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_cons_stats(void* lp_data,
                   int num_cols,
                   int num_rows,
                   SLEQP_BASESTAT* constraint_stats)
{
  Simplex* lp = (Simplex*)lp_data;

  memcpy(constraint_stats,
         lp->stat + num_cols,
         sizeof(SLEQP_BASESTAT) * num_rows);

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_basis_cond(void* lp_data, bool* exact, double* condition)
{
  (void)lp_data;

  // like lpi_highs.c:656-663
  *exact     = false;
  *condition = SLEQP_NONE;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_write(void* lp_data, const char* filename)
{
  Simplex* lp = (Simplex*)lp_data;

  FILE* out = fopen(filename, "w");

  if (!out)
  {
    sleqp_raise(SLEQP_INTERNAL_ERROR, "Cannot open %s", filename);
  }

  // CPLEX LP format (what SLEQP's debugging dumps are read with)
  const int n = lp->num_cols, m = lp->num_rows;

  fprintf(out, "Minimize\n obj:");
  for (int j = 0; j < n; ++j)
  {
    if (lp->cost[j] != 0.)
    {
      fprintf(out, " %+.17g x%d", lp->cost[j], j);
    }
  }
  fprintf(out, "\nSubject To\n");
  for (int i = 0; i < m; ++i)
  {
    for (int side = 0; side < 2; ++side)
    {
      const double bound = side ? lp->ub[n + i] : lp->lb[n + i];
      if (sleqp_is_infinite(fabs(bound)))
      {
        continue;
      }
      fprintf(out, " r%d_%c:", i, side ? 'u' : 'l');
      for (int j = 0; j < n; ++j)
      {
        const double a = lp->A[(size_t)j * m + i];
        if (a != 0.)
        {
          fprintf(out, " %+.17g x%d", a, j);
        }
      }
      fprintf(out, " %s %.17g\n", side ? "<=" : ">=", bound);
    }
  }
  fprintf(out, "Bounds\n");
  for (int j = 0; j < n; ++j)
  {
    fprintf(out, " ");
    if (is_neg_inf(lp->lb[j]))
    {
      fprintf(out, "-inf");
    }
    else
    {
      fprintf(out, "%.17g", lp->lb[j]);
    }
    fprintf(out, " <= x%d <= ", j);
    if (is_pos_inf(lp->ub[j]))
    {
      fprintf(out, "+inf\n");
    }
    else
    {
      fprintf(out, "%.17g\n", lp->ub[j]);
    }
  }
  fprintf(out, "End\n");
  fclose(out);

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
simplex_free(void** star)
{
  Simplex* lp = (Simplex*)(*star);

  if (!lp)
  {
    return SLEQP_OKAY;
  }

  for (int b = 0; b < lp->num_bases; ++b)
  {
    sleqp_free(&lp->saved[b]);
  }
  sleqp_free(&lp->saved);

  sleqp_free(&lp->cB);
  sleqp_free(&lp->column);
  sleqp_free(&lp->work);
  sleqp_free(&lp->dj);
  sleqp_free(&lp->y);
  sleqp_free(&lp->x);
  sleqp_free(&lp->scratch);
  sleqp_free(&lp->Binv);
  sleqp_free(&lp->position);
  sleqp_free(&lp->basic);
  sleqp_free(&lp->stat);
  sleqp_free(&lp->ub);
  sleqp_free(&lp->lb);
  sleqp_free(&lp->cost);
  sleqp_free(&lp->A);

  sleqp_free(star);

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_lpi_simplex_create(SleqpLPi** lp_star,
                         int num_cols,
                         int num_rows,
                         SleqpSettings* settings)
{
  SleqpLPiCallbacks callbacks = {.create_problem = simplex_create_problem,
                                 .solve          = simplex_solve,
                                 .status         = simplex_status,
                                 .set_bounds     = simplex_set_bounds,
                                 .set_coeffs     = simplex_set_coeffs,
                                 .set_obj        = simplex_set_objective,
                                 .set_basis      = simplex_set_basis,
                                 .save_basis     = simplex_save_basis,
                                 .restore_basis  = simplex_restore_basis,
                                 .primal_sol     = simplex_primal_sol,
                                 .dual_sol       = simplex_dual_sol,
                                 .vars_stats     = simplex_vars_stats,
                                 .cons_stats     = simplex_cons_stats,
                                 .basis_cond     = simplex_basis_cond,
                                 .write          = simplex_write,
                                 .free_problem   = simplex_free};

  return sleqp_lpi_create(lp_star,
                          SIMPLEX_NAME,
                          SIMPLEX_VERSION,
                          num_cols,
                          num_rows,
                          settings,
                          &callbacks);
}

SLEQP_RETCODE
sleqp_lpi_create_default(SleqpLPi** lp_interface,
                         int num_variables,
                         int num_constraints,
                         SleqpSettings* settings)
{
  SLEQP_CALL(sleqp_lpi_simplex_create(lp_interface,
                                      num_variables,
                                      num_constraints,
                                      settings));

  return SLEQP_OKAY;
}
