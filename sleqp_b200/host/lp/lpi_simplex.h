#ifndef SLEQP_LPI_SIMPLEX_H
#define SLEQP_LPI_SIMPLEX_H

/**
 * @file lpi_simplex.h
 * @brief A self-contained LP backend (dense bounded-variable primal simplex) for
 * environments without HiGHS / Gurobi / SoPlex.
 *
 * Counterpart of lp/lpi_highs.h in the reference tree.
 **/

#include "lp/lpi.h"

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_lpi_simplex_create(SleqpLPi** lp_star,
                         int num_variables,
                         int num_constraints,
                         SleqpSettings* settings);

#endif /* SLEQP_LPI_SIMPLEX_H */
