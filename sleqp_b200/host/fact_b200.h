#ifndef SLEQP_FACT_B200_H
#define SLEQP_FACT_B200_H

/**
 * @file fact_b200.h
 * @brief B200 (CUDA, sm_100a) sparse LDL^T factorization backend.
 *
 * Counterpart of fact_umfpack.h / fact_cholmod.h in the reference tree.
 **/

#include "fact.h"

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_fact_b200_create(SleqpFact** star, SleqpSettings* settings);

#endif /* SLEQP_FACT_B200_H */
