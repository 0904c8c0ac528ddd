#ifndef SLEQP_FACT_B200_H
#define SLEQP_FACT_B200_H

/**
 * @file fact_b200.h
 * @brief B200 (CUDA, sm_100a) sparse LDL^T factorization backend.
 *
 * Counterpart of fact_umfpack.h / fact_cholmod.h in the reference tree.
 **/

#include <stddef.h>

#include "fact.h"

struct b200_fact;

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_fact_b200_create(SleqpFact** star, SleqpSettings* settings);

/**
 * Page-locked caller buffers (shared by the factorization and the augmented Jacobian of
 * this backend): the arrays of the few SleqpVec objects SLEQP reuses for right-hand sides
 * and solutions are registered with the CUDA driver once, so that the device library DMAs
 * straight from / into them. Re-registers what sleqp_vec_reserve has moved.
 **/
#define SLEQP_B200_MAX_PINNED 16

typedef struct
{
  char* ptr;
  size_t bytes;
} SleqpB200Pinned;

typedef struct
{
  SleqpB200Pinned pinned[SLEQP_B200_MAX_PINNED];
  int num_pinned;
  int next_evict;
} SleqpB200Pins;

void
sleqp_b200_pin_buffer(SleqpB200Pins* pins, const void* buffer, size_t bytes);

void
sleqp_b200_unpin_all(SleqpB200Pins* pins);

/**
 * Device handle of the B200 factorization that last ran set_matrix on the calling
 * thread (NULL if none): what a B200 trust-region solver (tr/tr_b200.h) projects with
 * when it was created without an explicit handle.
 **/
struct b200_fact*
sleqp_fact_b200_last_handle(void);

void
sleqp_fact_b200_set_last_handle(struct b200_fact* handle);

#endif /* SLEQP_FACT_B200_H */
