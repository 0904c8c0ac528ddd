#ifndef SLEQP_MAT_B200_H
#define SLEQP_MAT_B200_H

/**
 * @file mat_b200.h
 * @brief Device-resident mirror of a SleqpMat for the products sleqp_mat_mult_vec /
 * sleqp_mat_mult_vec_trans (sparse/mat.c:282-363) on the B200.
 *
 * Counterpart of the two functions in sparse/mat.h for a matrix whose values stay
 * on the device between products: the constraint Jacobian of an iterate is refreshed
 * once per accepted step (iterate.c:87, sleqp_set_and_evaluate) and then multiplied
 * many times (direction.c:66,113; working_step.c:341; newton.c:377; util.c:62).
 **/

#include "sparse/mat.h"
#include "sparse/vec.h"

typedef struct SleqpMatB200 SleqpMatB200;

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_mat_b200_create(SleqpMatB200** star);

/** Copies pattern and values of the matrix to the device (call when the matrix changed). **/
SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_mat_b200_update(SleqpMatB200* mirror, const SleqpMat* matrix);

/** result = matrix * vector, same contract as sleqp_mat_mult_vec (mat.c:282-310). **/
SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_mat_b200_mult_vec(SleqpMatB200* mirror,
                        const SleqpVec* vector,
                        double* result);

/** result = matrix^T * vector, entries with |s| <= eps dropped, same contract as
 *  sleqp_mat_mult_vec_trans (mat.c:312-363). **/
SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_mat_b200_mult_vec_trans(SleqpMatB200* mirror,
                              const SleqpVec* vector,
                              double eps,
                              SleqpVec* result);

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_mat_b200_free(SleqpMatB200** star);

#endif /* SLEQP_MAT_B200_H */
