/*
 * mat_b200.c -- sleqp_mat_mult_vec / sleqp_mat_mult_vec_trans (src/main/sparse/mat.c:282-363)
 * on the B200, for a matrix whose values stay on the device between products.
 *
 * Meant to be dropped into the reference tree as src/main/sparse/mat_b200.c; the call
 * sites that would use it are listed in INTEGRATION.md (direction.c:66,113,
 * working_step.c:341, newton.c:377, util.c:62). C11, no numerical code: the products run
 * in libsleqp_b200.so (b200_mat_*), there is no CPU fallback.
 */
#include "mat_b200.h"

#include <assert.h>

#include <sleqp_b200.h>

#include "cmp.h"
#include "error.h"
#include "mem.h"

#include "fact/fact_b200.h"

struct SleqpMatB200
{
  b200_mat* handle;
  int num_rows;
  int num_cols;
  SleqpB200Pins pins; // page-locked caller arrays (matrix values, vectors, results): DMA without staging copies
};

#define B200_CALL(x)                                                           \
  do                                                                           \
  {                                                                            \
    const int b200_status = (x);                                               \
    if (b200_status != B200_OK)                                                \
    {                                                                          \
      sleqp_raise(SLEQP_INTERNAL_ERROR,                                        \
                  "B200 backend error %d: %s",                                 \
                  b200_status,                                                 \
                  b200_last_error());                                          \
    }                                                                          \
  } while (false)

SLEQP_RETCODE
sleqp_mat_b200_create(SleqpMatB200** star)
{
  SLEQP_CALL(sleqp_malloc(star));

  SleqpMatB200* mirror = *star;

  *mirror = (SleqpMatB200){0};

  const int status = b200_mat_create(&mirror->handle, -1);

  if (status != B200_OK)
  {
    sleqp_free(star);
    sleqp_raise(SLEQP_INTERNAL_ERROR,
                "B200 backend error %d: %s",
                status,
                b200_last_error());
  }

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_mat_b200_update(SleqpMatB200* mirror, const SleqpMat* matrix)
{
  mirror->num_rows = sleqp_mat_num_rows(matrix);
  mirror->num_cols = sleqp_mat_num_cols(matrix);

  sleqp_b200_pin_buffer(&mirror->pins,
                        sleqp_mat_data(matrix),
                        sizeof(double) * (size_t)sleqp_mat_nnz(matrix));

  B200_CALL(b200_mat_set(mirror->handle,
                         mirror->num_rows,
                         mirror->num_cols,
                         sleqp_mat_nnz(matrix),
                         sleqp_mat_cols(matrix),
                         sleqp_mat_rows(matrix),
                         sleqp_mat_data(matrix)));

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_mat_b200_mult_vec(SleqpMatB200* mirror,
                        const SleqpVec* vector,
                        double* result)
{
  assert(mirror->num_cols == vector->dim);

  sleqp_b200_pin_buffer(&mirror->pins,
                        vector->data,
                        sizeof(double) * (size_t)vector->nnz_max);
  sleqp_b200_pin_buffer(&mirror->pins,
                        result,
                        sizeof(double) * (size_t)mirror->num_rows);

  B200_CALL(b200_mat_mult_vec(mirror->handle,
                              vector->nnz,
                              vector->indices,
                              vector->data,
                              result));

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_mat_b200_mult_vec_trans(SleqpMatB200* mirror,
                              const SleqpVec* vector,
                              double eps,
                              SleqpVec* result)
{
  assert(mirror->num_rows == vector->dim);
  assert(mirror->num_cols == result->dim);

  // sparsified on the device like mat.c:355-358 (|s| <= eps dropped): only the kept entries come back
  SLEQP_CALL(sleqp_vec_clear(result));
  SLEQP_CALL(sleqp_vec_reserve(result, mirror->num_cols));

  int nnz = 0;

  B200_CALL(b200_mat_mult_vec_trans_sparse(mirror->handle,
                                           vector->nnz,
                                           vector->indices,
                                           vector->data,
                                           eps,
                                           result->indices,
                                           result->data,
                                           &nnz));

  result->nnz = nnz;

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_mat_b200_free(SleqpMatB200** star)
{
  SleqpMatB200* mirror = *star;

  if (!mirror)
  {
    return SLEQP_OKAY;
  }

  sleqp_b200_unpin_all(&mirror->pins);

  B200_CALL(b200_mat_free(&mirror->handle));

  sleqp_free(star);

  return SLEQP_OKAY;
}
