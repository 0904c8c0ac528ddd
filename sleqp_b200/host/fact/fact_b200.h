#ifndef SLEQP_FACT_B200_H
#define SLEQP_FACT_B200_H

/**
 * @file fact_b200.h
 * @brief B200 (CUDA, sm_100a) sparse LDL^T factorization backend.
 *
 * Counterpart of fact_umfpack.h / fact_cholmod.h in the reference tree.
 **/

#include "fact.h"

struct b200_fact;

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_fact_b200_create(SleqpFact** star, SleqpSettings* settings);

/**
 * Device handle of the B200 factorization that last ran set_matrix on the calling
 * thread (NULL if none): what a B200 trust-region solver (tr/tr_b200.h) projects with
 * when it was created without an explicit handle.
 **/
struct b200_fact*
sleqp_fact_b200_last_handle(void);

#endif /* SLEQP_FACT_B200_H */
