/*
 * fact_b200.c -- SLEQP factorization backend "B200": the five SleqpFactCallbacks
 * (src/main/fact/fact_types.h:25-32) over the C-ABI of libsleqp_b200.so.
 *
 * This file is meant to be dropped into the reference tree as src/main/fact/fact_b200.c and
 * compiled into sleqp_objects like its siblings (selection is link-time: every backend defines
 * sleqp_fact_create_default, e.g. fact_umfpack.c:325-331; see INTEGRATION.md for the CMake
 * entries). It is C11, uses only reference-internal headers plus <sleqp_b200.h>, and contains
 * no numerical code: the device library does everything and there is no CPU fallback.
 *
 * Flags: SLEQP_FACT_FLAGS_LOWER only (fact.h:9-14) -- the standard augmented Jacobian then
 * passes tril([I A_W^T; A_W 0]) (standard_aug_jac.c:267-271) and, without PSD, AUTO keeps the
 * indefinite standard form (trial_point.c:94-109).
 */
#include "fact_b200.h"

#include <assert.h>

#include <sleqp_b200.h>

#include "defs.h"
#include "error.h"
#include "mem.h"

typedef struct
{
  b200_fact* handle;
  int num_rows;
} B200Data;

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

static SLEQP_RETCODE
b200_set_matrix(void* fact_data, SleqpMat* matrix)
{
  B200Data* data = (B200Data*)fact_data;

  const int num_cols = sleqp_mat_num_cols(matrix);
  const int num_rows = sleqp_mat_num_rows(matrix);

  assert(num_cols == num_rows);

  // The matrix is borrowed (it is the aug_jac's own buffer, overwritten at the next
  // set_iterate, standard_aug_jac.c:143): the device library copies what it needs.
  B200_CALL(b200_fact_set_matrix(data->handle,
                                 num_rows,
                                 num_cols,
                                 sleqp_mat_nnz(matrix),
                                 sleqp_mat_cols(matrix),
                                 sleqp_mat_rows(matrix),
                                 sleqp_mat_data(matrix),
                                 /* lower_only = */ 1));

  data->num_rows = num_rows;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_solve(void* fact_data, const SleqpVec* rhs)
{
  B200Data* data = (B200Data*)fact_data;

  assert(rhs->dim == data->num_rows);

  // sparse right-hand side goes over as-is; the scatter into the zeroed dense vector
  // (set_cache / reset_cache, fact_umfpack.c:185-205) happens on the device
  B200_CALL(
    b200_fact_solve(data->handle, rhs->nnz, rhs->indices, rhs->data, rhs->dim));

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_solution(void* fact_data,
              SleqpVec* sol,
              int begin,
              int end,
              double zero_eps)
{
  B200Data* data = (B200Data*)fact_data;

  assert(begin <= end);

  const double* values = NULL;

  B200_CALL(b200_fact_solution_ptr(data->handle, begin, end, &values));

  SLEQP_CALL(sleqp_vec_set_from_raw(sol, values, end - begin, zero_eps));

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_condition(void* fact_data, double* condition)
{
  B200Data* data = (B200Data*)fact_data;

  double rcond = 0.;

  B200_CALL(b200_fact_rcond(data->handle, &rcond));

  // same convention as fact_umfpack.c:240 / fact_cholmod.c:206
  *condition = 1. / rcond;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_free(void** star)
{
  B200Data* data = (B200Data*)(*star);

  if (!data)
  {
    return SLEQP_OKAY;
  }

  B200_CALL(b200_fact_free(&data->handle));

  sleqp_free(&data);

  *star = NULL;

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_fact_b200_create(SleqpFact** star, SleqpSettings* settings)
{
  SleqpFactCallbacks callbacks = {.set_matrix = b200_set_matrix,
                                  .solve      = b200_solve,
                                  .solution   = b200_solution,
                                  .condition  = b200_condition,
                                  .free       = b200_free};

  B200Data* data = NULL;

  SLEQP_CALL(sleqp_malloc(&data));

  *data = (B200Data){0};

  // device -1: B200_DEVICE / LOCAL_RANK / 0; one handle (own CUDA stream) per SleqpFact, so
  // independent solver instances on different threads never share mutable state
  // (src/test/thread_test.c:92-110)
  const int status = b200_fact_create(&data->handle, -1);

  if (status != B200_OK)
  {
    sleqp_free(&data);
    sleqp_raise(SLEQP_INTERNAL_ERROR,
                "B200 backend error %d: %s",
                status,
                b200_last_error());
  }

  SLEQP_CALL(sleqp_fact_create(star,
                               SLEQP_FACT_B200_NAME,
                               SLEQP_FACT_B200_VERSION,
                               settings,
                               &callbacks,
                               SLEQP_FACT_FLAGS_LOWER,
                               (void*)data));

  return SLEQP_OKAY;
}

SLEQP_RETCODE
sleqp_fact_create_default(SleqpFact** star, SleqpSettings* settings)
{
  SLEQP_CALL(sleqp_fact_b200_create(star, settings));

  return SLEQP_OKAY;
}
