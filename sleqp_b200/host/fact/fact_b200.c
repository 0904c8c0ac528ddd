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

// Page-locked caller buffers. The aug_jac reuses a handful of SleqpVec objects for its right-hand sides and
// solutions (standard_aug_jac.c:306-435): their data / indices arrays are registered with the CUDA driver once, so
// that the device library DMAs straight from / into them (b200_host_pin; a pageable buffer costs a staging copy and,
// for solutions, a host pass over the slice). A vector that has been reallocated since (sleqp_vec_reserve) shows up
// as a new pointer: whatever overlapped the new range is released first.
typedef struct
{
  b200_fact* handle;
  int num_rows;
  SleqpB200Pins pins;
} B200Data;

// handle of the factorization that last ran set_matrix on this thread (one SleqpFact per solver thread,
// src/test/thread_test.c:92-110): what a B200 trust-region solver created without an explicit handle projects with
static _Thread_local b200_fact* last_handle = NULL;

b200_fact*
sleqp_fact_b200_last_handle(void)
{
  return last_handle;
}

void
sleqp_fact_b200_set_last_handle(b200_fact* handle)
{
  last_handle = handle;
}

void
sleqp_b200_pin_buffer(SleqpB200Pins* data, const void* buffer, size_t bytes)
{
  char* ptr = (char*)buffer;

  if (!ptr || bytes < (1 << 16)) // small vectors: the staging copy is cheaper than a registration
  {
    return;
  }

  for (int i = 0; i < data->num_pinned; ++i)
  {
    if (data->pinned[i].ptr == ptr && data->pinned[i].bytes >= bytes)
    {
      return;
    }
  }

  // release whatever overlaps the new range (stale registrations of reallocated vectors)
  for (int i = 0; i < data->num_pinned;)
  {
    SleqpB200Pinned* reg = data->pinned + i;

    if (reg->ptr < ptr + bytes && ptr < reg->ptr + reg->bytes)
    {
      b200_host_unpin(reg->ptr);
      *reg = data->pinned[--data->num_pinned];
    }
    else
    {
      ++i;
    }
  }

  if (data->num_pinned == SLEQP_B200_MAX_PINNED)
  {
    SleqpB200Pinned* reg = data->pinned + (data->next_evict++ % SLEQP_B200_MAX_PINNED);
    b200_host_unpin(reg->ptr);
    *reg = data->pinned[--data->num_pinned];
  }

  // failure is not an error: the library falls back to its own staging buffer for pageable memory
  if (b200_host_pin(ptr, bytes) == B200_OK)
  {
    data->pinned[data->num_pinned++] = (SleqpB200Pinned){ptr, bytes};
  }
}

void
sleqp_b200_unpin_all(SleqpB200Pins* data)
{
  for (int i = 0; i < data->num_pinned; ++i)
  {
    b200_host_unpin(data->pinned[i].ptr);
  }

  data->num_pinned = 0;
}

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

  last_handle = data->handle;

  return SLEQP_OKAY;
}

static SLEQP_RETCODE
b200_solve(void* fact_data, const SleqpVec* rhs)
{
  B200Data* data = (B200Data*)fact_data;

  assert(rhs->dim == data->num_rows);

  // sparse right-hand side goes over as-is; the scatter into the zeroed dense vector
  // (set_cache / reset_cache, fact_umfpack.c:185-205) happens on the device
  sleqp_b200_pin_buffer(&data->pins, rhs->data, sizeof(double) * (size_t)rhs->nnz_max);

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

  // What sleqp_vec_set_from_raw (vec.c:72-104) would build from the dense slice, built on the device instead:
  // the entries with |v| > zero_eps in ascending order are written straight into the (page-locked) arrays of `sol`.
  // The arrays are reserved at full length because the device copies them at full length (one synchronisation
  // serves the count and the data).
  const int dim = end - begin;

  SLEQP_CALL(sleqp_vec_clear(sol));
  SLEQP_CALL(sleqp_vec_resize(sol, dim));
  SLEQP_CALL(sleqp_vec_reserve(sol, dim));

  sleqp_b200_pin_buffer(&data->pins, sol->data, sizeof(double) * (size_t)sol->nnz_max);
  sleqp_b200_pin_buffer(&data->pins, sol->indices, sizeof(int) * (size_t)sol->nnz_max);

  int nnz = 0;

  B200_CALL(b200_fact_solution_sparse(data->handle,
                                      begin,
                                      end,
                                      zero_eps,
                                      sol->indices,
                                      sol->data,
                                      &nnz));

  sol->nnz = nnz;

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

  sleqp_b200_unpin_all(&data->pins);

  if (last_handle == data->handle)
  {
    last_handle = NULL;
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
