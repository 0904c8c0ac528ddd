// common.hpp -- shared declarations of the host side of libsleqp_b200.so
#pragma once
#include "../../include/sleqp_b200.h"
#include "plan.hpp"

#include <memory>
#include <string>

namespace b200
{

extern thread_local std::string g_last_error;

int set_error(int code, const std::string& msg);

// Pattern-keyed cache lookup (or fresh analysis). `cached` reports a hit.
int get_plan(int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only, std::shared_ptr<const Plan>& out, bool& cached);

// Same for a KKT system given as (constraint Jacobian, working-set index maps) -- b200_fact_set_kkt: the key is a hash of
// those arrays; on a miss tril(K) is laid out like the reference's fill_aug_jac and analysed.
int get_plan_kkt(int num_vars, int num_cons, int nnz_jac, const int* jac_cols, const int* jac_rows, const double* jac_data, const int* var_index,
                 const int* cons_index, int ws_size, std::shared_ptr<const Plan>& out, bool& cached);

void fill_stats_from_plan(const Plan& P, b200_stats* s);

} // namespace b200
