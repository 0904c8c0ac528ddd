// capi_symbolic.cpp -- host-only C-ABI entry points (symbolic analysis, plan cache, errors).
#include "common.hpp"

#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>

namespace b200
{

thread_local std::string g_last_error;

int
set_error(int code, const std::string& msg)
{
  g_last_error = msg;
  return code;
}

// Process-wide pattern-keyed cache of analyses (north-star: "done once per sparsity pattern and
// cached across SQP iterations"). Plans are immutable, so handles on different threads/devices
// share them read-only.
namespace
{
std::mutex g_cache_mutex;
std::list<std::shared_ptr<const Plan>> g_cache; // most recently used first
constexpr size_t CACHE_CAPACITY = 16;
} // namespace

// The analysis depends on a few environment knobs (measurements and tests: B200_SST, B200_GROUP_CAP,
// B200_FLOW_DEEP_TASKS): they are part of the cache key, so a plan built under other settings is never handed out.
static uint64_t
plan_variant_bits()
{
  uint64_t v = 0;
  for (const char* name : {"B200_SST", "B200_GROUP_CAP", "B200_FLOW_DEEP_TASKS"})
  {
    const char* e = std::getenv(name);
    v             = v * 0x100000001B3ull + 0x9E37;
    for (; e && *e; ++e)
    {
      v = (v ^ (unsigned char)*e) * 0x100000001B3ull;
    }
  }
  return v;
}

int
get_plan(int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only, std::shared_ptr<const Plan>& out, bool& cached)
{
  if (!valid_csc_header(n, nnz, colptr, rowidx, val))
  {
    return set_error(B200_ERR_ARG, "malformed CSC header (null arrays, colptr not monotone or out of range)");
  }
  uint64_t h2      = 0;
  const uint64_t h = hash_pattern(n, nnz, colptr, rowidx, val, lower_only, &h2);
  h2 ^= plan_variant_bits();
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
    {
      if (!(*it)->kkt_keyed && (*it)->pattern_hash == h && (*it)->pattern_hash2 == h2 && (*it)->N == n && (*it)->nnzK_input == nnz)
      {
        out = *it;
        g_cache.splice(g_cache.begin(), g_cache, it);
        cached = true;
        return B200_OK;
      }
    }
  }
  auto plan = std::make_shared<Plan>();
  std::string err;
  int rc = analyze(n, nnz, colptr, rowidx, val, lower_only, *plan, err);
  if (rc != B200_OK)
  {
    return set_error(rc, err);
  }
  plan->pattern_hash2 = h2; // incl. the variant bits
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_cache.push_front(plan);
    while (g_cache.size() > CACHE_CAPACITY)
    {
      g_cache.pop_back();
    }
  }
  out    = plan;
  cached = false;
  return B200_OK;
}

int
get_plan_kkt(int num_vars, int num_cons, int nnz_jac, const int* jac_cols, const int* jac_rows, const double* jac_data, const int* var_index, const int* cons_index,
             int ws_size, std::shared_ptr<const Plan>& out, bool& cached)
{
  if (num_vars < 0 || num_cons < 0 || nnz_jac < 0 || ws_size < 0 || !jac_cols || !var_index || (num_cons > 0 && !cons_index) || (nnz_jac > 0 && (!jac_rows || !jac_data))
      || jac_cols[0] != 0 || jac_cols[num_vars] != nnz_jac)
  {
    return set_error(B200_ERR_ARG, "malformed Jacobian / working-set arrays");
  }
  uint64_t h2      = 0;
  const uint64_t h = hash_kkt(num_vars, num_cons, nnz_jac, jac_cols, jac_rows, var_index, cons_index, ws_size, &h2);
  h2 ^= plan_variant_bits();
  const int N      = num_vars + ws_size;
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (auto it = g_cache.begin(); it != g_cache.end(); ++it)
    {
      if ((*it)->kkt_keyed && (*it)->pattern_hash == h && (*it)->pattern_hash2 == h2 && (*it)->N == N && (*it)->key_aux == nnz_jac)
      {
        out = *it;
        g_cache.splice(g_cache.begin(), g_cache, it);
        cached = true;
        return B200_OK;
      }
    }
  }
  std::vector<int> colptr, rowidx, src;
  if (!build_kkt_lower(num_vars, num_cons, jac_cols, jac_rows, var_index, cons_index, ws_size, colptr, rowidx, src))
  {
    return set_error(B200_ERR_ARG, "malformed Jacobian / working-set arrays (indices out of range or rows not increasing)");
  }
  std::vector<double> val(src.size());
  for (size_t q = 0; q < src.size(); ++q)
  {
    val[q] = src[q] < 0 ? 1.0 : jac_data[src[q]];
  }
  auto plan = std::make_shared<Plan>();
  std::string err;
  int rc = analyze(N, (int)rowidx.size(), colptr.data(), rowidx.data(), val.data(), /*lower_only=*/1, *plan, err);
  if (rc != B200_OK)
  {
    return set_error(rc, err);
  }
  plan->kkt_keyed     = true;
  plan->key_aux       = nnz_jac;
  plan->pattern_hash  = h;
  plan->pattern_hash2 = h2;
  plan->Ksrc          = std::move(src);
  {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    g_cache.push_front(plan);
    while (g_cache.size() > CACHE_CAPACITY)
    {
      g_cache.pop_back();
    }
  }
  out    = plan;
  cached = false;
  return B200_OK;
}

void
fill_stats_from_plan(const Plan& P, b200_stats* s)
{
  std::memset(s, 0, sizeof(*s));
  s->n                   = P.N;
  s->n_elim              = P.nE;
  s->n_reduced           = P.m;
  s->nnz_K               = P.nnzK;
  s->nnz_S               = P.nnzS;
  s->nnz_L               = P.nnzL;
  s->nnz_L_stored        = P.nnzL_stored;
  s->n_row_idx           = (int64_t)P.Ridx.size();
  s->n_supernodes        = P.nsuper;
  s->n_levels            = P.nlevels;
  s->n_stages            = (int)P.stages.size();
  s->max_front           = P.max_front;
  s->flops_factor        = P.flops;
  s->flops_factor_stored = P.flops_stored;
  s->update_ws_doubles   = P.Utotal;
  s->pattern_hash        = P.pattern_hash;
  s->pattern_hash2       = P.pattern_hash2;
  s->flops_update        = P.flops_update;
  s->flops_inv           = P.flops_inv;
  s->panel_doubles       = P.Lptr.empty() ? 0 : P.Lptr[P.nsuper];
  s->n_demoted           = P.n_demoted;
  s->perm_hash           = P.perm_hash;
  s->ms_symbolic         = P.ms_symbolic;
  s->n_scratch_slots     = P.n_scratch_slots;
}

} // namespace b200

using namespace b200;

struct b200_symbolic
{
  std::shared_ptr<const Plan> plan;
  bool cached;
};

namespace
{
template <typename T>
int
export_vec(const std::vector<T>& v, void* out, int64_t* count)
{
  if (count)
  {
    *count = (int64_t)v.size();
  }
  if (out && !v.empty())
  {
    std::memcpy(out, v.data(), sizeof(T) * v.size());
  }
  return B200_OK;
}
} // namespace

extern "C" {

const char*
b200_last_error(void)
{
  return g_last_error.c_str();
}

int
b200_symbolic_analyze(b200_symbolic** out, int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only)
{
  if (!out)
  {
    return set_error(B200_ERR_ARG, "null output handle");
  }
  *out = nullptr;
  // Always analyse afresh here (this entry point exists for tests and timing of the analysis);
  // b200_fact_set_matrix goes through the cache.
  auto plan = std::make_shared<Plan>();
  std::string err;
  int rc = analyze(n, nnz, colptr, rowidx, val, lower_only, *plan, err);
  if (rc != B200_OK)
  {
    return set_error(rc, err);
  }
  *out = new b200_symbolic{plan, false};
  return B200_OK;
}

int
b200_symbolic_analyze_kkt(b200_symbolic** out,
                          int num_vars,
                          int num_cons,
                          int nnz_jac,
                          const int* jac_cols,
                          const int* jac_rows,
                          const double* jac_data,
                          const int* var_index,
                          const int* cons_index,
                          int working_set_size)
{
  if (!out)
  {
    return set_error(B200_ERR_ARG, "null output handle");
  }
  *out = nullptr;
  if (num_vars < 0 || num_cons < 0 || nnz_jac < 0 || working_set_size < 0 || !jac_cols || !var_index || (num_cons > 0 && !cons_index)
      || (nnz_jac > 0 && (!jac_rows || !jac_data)) || jac_cols[0] != 0 || jac_cols[num_vars] != nnz_jac)
  {
    return set_error(B200_ERR_ARG, "malformed Jacobian / working-set arrays");
  }
  std::vector<int> colptr, rowidx, src;
  if (!build_kkt_lower(num_vars, num_cons, jac_cols, jac_rows, var_index, cons_index, working_set_size, colptr, rowidx, src))
  {
    return set_error(B200_ERR_ARG, "malformed Jacobian / working-set arrays (indices out of range or rows not increasing)");
  }
  std::vector<double> val(src.size());
  for (size_t q = 0; q < src.size(); ++q)
  {
    val[q] = src[q] < 0 ? 1.0 : jac_data[src[q]];
  }
  auto plan = std::make_shared<Plan>();
  std::string err;
  int rc = analyze(num_vars + working_set_size, (int)rowidx.size(), colptr.data(), rowidx.data(), val.data(), 1, *plan, err);
  if (rc != B200_OK)
  {
    return set_error(rc, err);
  }
  plan->kkt_keyed = true;
  plan->key_aux   = nnz_jac;
  plan->Ksrc      = std::move(src);
  *out            = new b200_symbolic{plan, false};
  return B200_OK;
}

int
b200_symbolic_stats(const b200_symbolic* s, b200_stats* stats)
{
  if (!s || !stats)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  fill_stats_from_plan(*s->plan, stats);
  return B200_OK;
}

int
b200_symbolic_structure(const b200_symbolic* s, int* perm, int* parent, int* colcount, int* n_super_total, int* super_first)
{
  if (!s)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  full_structure(*s->plan, perm, parent, colcount, n_super_total, super_first);
  return B200_OK;
}


// Field list (element type): see sleqp_b200/fact.py PLAN_FIELDS.
int
b200_symbolic_export(const b200_symbolic* s, const char* field, void* out, int64_t* count)
{
  if (!s || !field)
  {
    return set_error(B200_ERR_ARG, "null argument");
  }
  const Plan& P = *s->plan;
  const std::string f(field);
#define FIELD(name)                                                                                                     \
  if (f == #name)                                                                                                       \
  {                                                                                                                     \
    return export_vec(P.name, out, count);                                                                              \
  }
  FIELD(e_of_k)
  FIELD(r_of_k)
  FIELD(k_of_e)
  FIELD(k_of_r)
  FIELD(dE_src)
  FIELD(Acsc_ptr)
  FIELD(Acsc_row)
  FIELD(Acsc_src)
  FIELD(Acsr_ptr)
  FIELD(Acsr_col)
  FIELD(Acsr_src)
  FIELD(Gsym_ptr)
  FIELD(Gsym_col)
  FIELD(Gsym_src)
  FIELD(perm)
  FIELD(pinv)
  FIELD(parent)
  FIELD(colcount)
  FIELD(sn_first)
  FIELD(sn_of_col)
  FIELD(sn_parent)
  FIELD(sn_level)
  FIELD(Rptr)
  FIELD(Ridx)
  FIELD(rel)
  FIELD(Lptr)
  FIELD(Wptr)
  FIELD(child_ptr)
  FIELD(child_idx)
  FIELD(Sdest)
  FIELD(Sgsrc)
  FIELD(Sterm_ptr)
  FIELD(Sterm_a)
  FIELD(Sterm_b)
  FIELD(Sterm_d)
  FIELD(Uoff)
  FIELD(sn_base)
  FIELD(sn_nt)
  FIELD(zero_sn)
  FIELD(lvl_ptr)
  FIELD(lvl_sn)
  FIELD(inv_phase_ptr)
  FIELD(Tptr)
  FIELD(Ksrc)
  FIELD(sst_colptr)
  FIELD(sst_rows)
  FIELD(sst_ea_src)
  FIELD(sst_ea_dst)
  FIELD(sst_gen_ptr)
#undef FIELD
  if (f == "sst_blob")
  {
    std::vector<int> v(P.sst_blob.begin(), P.sst_blob.end());
    return export_vec(v, out, count);
  }
  if (f == "sn_sparse")
  {
    std::vector<int> v(P.sn_sparse.begin(), P.sn_sparse.end());
    return export_vec(v, out, count);
  }
  if (f == "sst")
  {
    static_assert(sizeof(SstMeta) == 30 * sizeof(int), "SstMeta layout");
    return export_vec(P.sst, out, count);
  }
  if (f == "stages")
  {
    static_assert(sizeof(Stage) == 10 * sizeof(int), "Stage layout");
    return export_vec(P.stages, out, count);
  }
  if (f == "ea_tasks")
  {
    return export_vec(P.ea_tasks, out, count);
  }
  if (f == "pan_tasks")
  {
    return export_vec(P.pan_tasks, out, count);
  }
  if (f == "inv_tasks")
  {
    return export_vec(P.inv_tasks, out, count);
  }
  if (f == "tr_tasks")
  {
    return export_vec(P.tr_tasks, out, count);
  }
  if (f == "ffl_tasks")
  {
    return export_vec(P.ffl_tasks, out, count);
  }
  if (f == "bfl_tasks")
  {
    return export_vec(P.bfl_tasks, out, count);
  }
  if (f == "upd_tasks")
  {
    return export_vec(P.upd_tasks, out, count);
  }
  return set_error(B200_ERR_ARG, "unknown plan field: " + f);
}

int
b200_symbolic_free(b200_symbolic** s)
{
  if (s && *s)
  {
    delete *s;
    *s = nullptr;
  }
  return B200_OK;
}

} // extern "C"
