#ifndef SLEQP_TR_B200_H
#define SLEQP_TR_B200_H

/**
 * @file tr_b200.h
 * @brief B200 (CUDA, sm_100a) trust-region (EQP) solver: Steihaug's projected CG with
 * all vectors resident on the device.
 *
 * Counterpart of tr/steihaug_solver.h / tr/trlib_solver.h in the reference tree.
 **/

#include "tr/tr_solver.h"

#include "sparse/pub_mat.h"

/**
 * Creates a B200 trust-region solver. It registers the three SleqpTRCallbacks
 * (solve / rayleigh / free, tr/tr_types.h:9-30) exactly like
 * sleqp_steihaug_solver_create (tr/steihaug_solver.c:498-540) and runs the same
 * algorithm; projections onto the null space of the working-set rows use the B200
 * factorization of the augmented Jacobian system directly on the device
 * (sleqp_fact_b200_last_handle), so the given SleqpAugJac must be the standard
 * augmented Jacobian over the B200 factorization.
 *
 * Hessian products go through sleqp_problem_hess_prod (matrix-free, problem.c:632)
 * unless a Hessian matrix has been provided with sleqp_tr_b200_set_hessian.
 **/
SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_b200_tr_solver_create(SleqpTRSolver** solver_star,
                            SleqpProblem* problem,
                            SleqpSettings* settings);

/**
 * Provides the Hessian of the Lagrangian at the current iterate as a sparse
 * matrix (num_vars x num_vars, full symmetric, CSC). The values are copied to the
 * device; products then never leave it. Pass NULL to return to the
 * callback. Front ends that have the Hessian as a matrix (AMPL, the synthetic
 * configurations of bench.py) call this once per iterate, next to
 * sleqp_aug_jac_set_iterate.
 **/
SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_tr_b200_set_hessian(SleqpTRSolver* solver, const SleqpMat* hessian);

/** Number of CG iterations and exit (B200_CG_*) of the last solve. **/
SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_tr_b200_last_solve(SleqpTRSolver* solver, int* iterations, int* exit_code);

#endif /* SLEQP_TR_B200_H */
