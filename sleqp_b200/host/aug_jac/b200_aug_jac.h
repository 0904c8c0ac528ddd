#ifndef SLEQP_B200_AUG_JAC_H
#define SLEQP_B200_AUG_JAC_H

/**
 * @file b200_aug_jac.h
 * @brief Augmented Jacobian system [I A_W^T; A_W 0] assembled and factorized on the B200.
 *
 * Counterpart of aug_jac/standard_aug_jac.h: same callbacks (aug_jac/aug_jac_types.h),
 * same results; what changes is that set_iterate does not build the augmented matrix on
 * the host. The constraint Jacobian and the working set (as index maps) go to the device
 * library, which lays tril(K) out like fill_aug_jac (standard_aug_jac.c:135-237) only when
 * (Jacobian pattern, working set) is new and otherwise gathers the values on the device.
 **/

#include "aug_jac/aug_jac.h"

SLEQP_WARNUNUSED
SLEQP_RETCODE
sleqp_b200_aug_jac_create(SleqpAugJac** star,
                          SleqpProblem* problem,
                          SleqpSettings* settings);

#endif /* SLEQP_B200_AUG_JAC_H */
