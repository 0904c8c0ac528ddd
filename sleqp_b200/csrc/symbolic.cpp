// symbolic.cpp -- host-side symbolic analysis, done once per sparsity pattern and cached
// (north-star subsystem 1: ordering, elimination tree, supernode amalgamation).
//
// No reference backend has an equivalent: Umfpack/CHOLMOD redo their analysis inside the
// vendor library on every set_matrix (fact_umfpack.c:139-160, fact_cholmod.c:133). The
// algorithms here are the published ones: George's automatic nested dissection on level
// structures, Liu's elimination tree with path compression, the Gilbert-Ng-Peyton column
// counts, fundamental supernodes plus relaxed amalgamation.
#include "plan.hpp"

#include "../../include/sleqp_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <numeric>
#include <thread>

namespace b200
{

static inline uint64_t
fnv1a(uint64_t h, const void* data, size_t bytes)
{
  const unsigned char* p = (const unsigned char*)data;
  for (size_t i = 0; i < bytes; ++i)
  {
    h ^= p[i];
    h *= 1099511628211ull;
  }
  return h;
}

// Fast 64-bit hash over 8-byte words, four independent lanes (this runs on every set_matrix, over
// the whole pattern: ~16 MB for config 3, so a byte-wise FNV would cost more than the GPU factorization).
static inline uint64_t
mix64(uint64_t h, uint64_t v)
{
  h ^= v;
  h *= 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  return h;
}

// second, independent mixing function: the plan cache is keyed by both hashes (and n, nnz), so a collision of one
// 64-bit hash alone cannot hand out another pattern's plan
static inline uint64_t
mix64b(uint64_t h, uint64_t v)
{
  h = ((h << 23) | (h >> 41)) + v;
  h *= 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 31;
  return h;
}

struct Hash2
{
  uint64_t a, b;
};

static Hash2
hash_words(Hash2 seed, const void* data, size_t bytes)
{
  const unsigned char* p = (const unsigned char*)data;
  uint64_t h0 = seed.a ^ 0x243F6A8885A308D3ull, h1 = seed.a ^ 0x13198A2E03707344ull, h2 = seed.a ^ 0xA4093822299F31D0ull, h3 = seed.a ^ 0x082EFA98EC4E6C89ull;
  uint64_t g0 = seed.b ^ 0x452821E638D01377ull, g1 = seed.b ^ 0xBE5466CF34E90C6Cull;
  size_t i = 0;
  for (; i + 32 <= bytes; i += 32)
  {
    uint64_t w[4];
    std::memcpy(w, p + i, 32);
    h0 = mix64(h0, w[0]);
    h1 = mix64(h1, w[1]);
    h2 = mix64(h2, w[2]);
    h3 = mix64(h3, w[3]);
    g0 = mix64b(g0, w[0] ^ (w[2] << 1 | w[2] >> 63));
    g1 = mix64b(g1, w[1] + (w[3] << 7 | w[3] >> 57));
  }
  uint64_t tail[4] = {0, 0, 0, 0};
  std::memcpy(tail, p + i, bytes - i);
  h0 = mix64(h0, tail[0]);
  h1 = mix64(h1, tail[1]);
  h2 = mix64(h2, tail[2]);
  h3 = mix64(h3, tail[3] ^ (uint64_t)bytes);
  g0 = mix64b(g0, tail[0] ^ (tail[2] << 1 | tail[2] >> 63));
  g1 = mix64b(g1, tail[1] + (tail[3] << 7 | tail[3] >> 57) + (uint64_t)bytes);
  return {mix64(mix64(mix64(h0, h1), h2), h3), mix64b(g0, g1)};
}

// O(n) sanity of the column pointers: everything the hash (and the analysis) dereferences through them is in range
// afterwards. Row indices are validated by analyze(); the hash only reads them, it never indexes with them.
bool
valid_csc_header(int n, int nnz, const int* colptr, const int* rowidx, const double* val)
{
  if (n < 0 || nnz < 0 || !colptr || (nnz > 0 && (!rowidx || !val)) || colptr[0] != 0 || colptr[n] != nnz)
  {
    return false;
  }
  for (int j = 0; j < n; ++j)
  {
    if (colptr[j + 1] < colptr[j] || colptr[j + 1] > nnz)
    {
      return false;
    }
  }
  return true;
}

uint64_t
hash_pattern(int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only, uint64_t* second)
{
  // On the critical path of every set_matrix (the plan is looked up by this key before anything can be launched):
  // large patterns are hashed in four independent slices on host threads, the slice hashes are chained in order.
  // The caller has checked the header (valid_csc_header).
  constexpr int SLICES = 4;
  const int nsl        = (long long)n + nnz >= 400000 ? SLICES : 1;
  Hash2 part[SLICES]   = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
  auto slice = [&](int sl) {
    const int j0 = (int)((long long)n * sl / nsl), j1 = (int)((long long)n * (sl + 1) / nsl);
    Hash2 h = hash_words({(uint64_t)(0x5EED + sl), (uint64_t)(0xB200 + sl)}, colptr + j0, sizeof(int) * (size_t)(j1 - j0 + 1));
    h       = hash_words(h, rowidx + colptr[j0], sizeof(int) * (size_t)(colptr[j1] - colptr[j0]));
    // the E/R classification depends on which diagonals are non-zero: part of the key
    uint64_t bits = 0;
    for (int j = j0; j < j1; ++j)
    {
      int p         = colptr[j];
      const int end = colptr[j + 1];
      if (!lower_only)
      {
        while (p < end && rowidx[p] < j)
        {
          ++p;
        }
      }
      const uint64_t nz = (p < end && rowidx[p] == j && val[p] != 0.) ? 1u : 0u;
      bits              = (bits << 1) | nz;
      if (((j - j0) & 63) == 63)
      {
        h.a  = mix64(h.a, bits);
        h.b  = mix64b(h.b, bits);
        bits = 0;
      }
    }
    part[sl] = {mix64(h.a, bits ^ 0xD1A6), mix64b(h.b, bits ^ 0x6A1D)};
  };
  if (nsl == 1)
  {
    slice(0);
  }
  else
  {
    std::thread pool[SLICES - 1];
    for (int sl = 1; sl < nsl; ++sl)
    {
      pool[sl - 1] = std::thread(slice, sl);
    }
    slice(0);
    for (int sl = 1; sl < nsl; ++sl)
    {
      pool[sl - 1].join();
    }
  }
  int hdr[4] = {n, nnz, lower_only ? 1 : 0, nsl};
  Hash2 h    = hash_words({0x5EED, 0xB200}, hdr, sizeof(hdr));
  for (int sl = 0; sl < nsl; ++sl)
  {
    h.a = mix64(h.a, part[sl].a);
    h.b = mix64b(h.b, part[sl].b);
  }
  if (second)
  {
    *second = h.b;
  }
  return h.a;
}

uint64_t
hash_kkt(int num_vars, int num_cons, int nnz_jac, const int* jac_cols, const int* jac_rows, const int* var_index, const int* cons_index, int ws_size, uint64_t* second)
{
  // This is on the critical path of every set_iterate (16 MB at config 3). The four arrays are cut into pieces of at most
  // 2 MB, every piece is hashed on its own (seeded with its number) and the piece hashes are combined in order: the
  // value does not depend on how many threads share the pieces out.
  int hdr[4] = {num_vars, num_cons, nnz_jac, ws_size};
  struct Piece
  {
    const unsigned char* p;
    size_t bytes;
  };
  std::vector<Piece> pieces;
  constexpr size_t PIECE = (size_t)2 << 20;
  auto cut = [&](const int* a, size_t count) {
    const unsigned char* p = (const unsigned char*)a;
    size_t bytes           = sizeof(int) * count;
    if (bytes == 0)
    {
      pieces.push_back({(const unsigned char*)"", 0}); // an empty array still takes its place in the sequence
      return;
    }
    do
    {
      const size_t len = std::min(bytes, PIECE);
      pieces.push_back({p, len});
      p += len;
      bytes -= len;
    } while (bytes > 0);
  };
  cut(jac_cols, (size_t)num_vars + 1);
  cut(jac_rows, (size_t)nnz_jac);
  cut(var_index, (size_t)num_vars);
  cut(cons_index, (size_t)num_cons);
  std::vector<Hash2> part(pieces.size());
  auto run = [&](size_t q) {
    const Hash2 seed = {(uint64_t)(0x4B4B54 + q), (uint64_t)(0xB2004B + q)};
    part[q]          = hash_words(seed, pieces[q].p, pieces[q].bytes);
  };
  int nthreads = 1;
  if ((long long)num_vars + nnz_jac >= 400000)
  {
    nthreads = (int)std::min<size_t>(pieces.size(), (size_t)std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency())));
    if (const char* ht = std::getenv("B200_HOST_THREADS"))
    {
      nthreads = std::max(1, std::min(nthreads, std::atoi(ht)));
    }
  }
  if (nthreads > 1)
  {
    std::vector<std::thread> pool;
    auto worker = [&](int w) {
      for (size_t q = (size_t)w; q < pieces.size(); q += (size_t)nthreads)
      {
        run(q);
      }
    };
    for (int w = 1; w < nthreads; ++w)
    {
      pool.emplace_back(worker, w);
    }
    worker(0);
    for (auto& th : pool)
    {
      th.join();
    }
  }
  else
  {
    for (size_t q = 0; q < pieces.size(); ++q)
    {
      run(q);
    }
  }
  Hash2 h = hash_words({0x4B4B54, 0xB2004B}, hdr, sizeof(hdr));
  for (size_t q = 0; q < pieces.size(); ++q)
  {
    h.a = mix64(h.a, part[q].a);
    h.b = mix64b(h.b, part[q].b);
  }
  if (second)
  {
    *second = h.b;
  }
  return h.a;
}

bool
build_kkt_lower(int num_vars, int num_cons, const int* jac_cols, const int* jac_rows, const int* var_index, const int* cons_index, int ws_size, std::vector<int>& colptr,
                std::vector<int>& rowidx, std::vector<int>& src)
{
  const int n = num_vars, N = num_vars + ws_size;
  colptr.assign((size_t)N + 1, 0);
  rowidx.clear();
  src.clear();
  rowidx.reserve((size_t)n + (size_t)jac_cols[n] + 16);
  src.reserve((size_t)n + (size_t)jac_cols[n] + 16);
  for (int j = 0; j < n; ++j)
  {
    rowidx.push_back(j); // identity part first (standard_aug_jac.c:160)
    src.push_back(-1);
    int last = -1;
    if (var_index[j] != -1)
    {
      if (var_index[j] < 0 || var_index[j] >= ws_size)
      {
        return false;
      }
      last = n + var_index[j];
      rowidx.push_back(last); // unit row of the active bound (:171-176)
      src.push_back(-1);
    }
    if (jac_cols[j + 1] < jac_cols[j])
    {
      return false;
    }
    for (int q = jac_cols[j]; q < jac_cols[j + 1]; ++q)
    {
      const int row = jac_rows[q];
      if (row < 0 || row >= num_cons)
      {
        return false;
      }
      const int ci = cons_index[row];
      if (ci == -1)
      {
        continue;
      }
      if (ci < 0 || ci >= ws_size || n + ci <= last) // rows must come out strictly increasing (mat.c:797-804)
      {
        return false;
      }
      last = n + ci;
      rowidx.push_back(last);
      src.push_back(q);
    }
    colptr[(size_t)j + 1] = (int)rowidx.size();
    if (rowidx.size() > (size_t)0x7ffffff0)
    {
      return false;
    }
  }
  for (int j = n; j < N; ++j)
  {
    colptr[(size_t)j + 1] = colptr[(size_t)n]; // empty columns of the working set (:221-226)
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// Nested dissection on rooted level structures (George 1973; George & Liu 1978).
// Subproblems are independent once split, so they are processed by a small pool of host threads
// (the analysis is on the critical path of every set_matrix with a new working set).
// ---------------------------------------------------------------------------------------
namespace
{

struct Graph
{
  int n;
  const std::vector<int>& xadj;
  const std::vector<int>& adj;
};

// Node-indexed scratch shared by all threads (they work on disjoint node sets; `tag` is also read for
// neighbours that belong to other subproblems, hence atomic with relaxed ordering) ...
struct NDShared
{
  std::vector<std::atomic<int>> tag; // subproblem id a node currently belongs to
  std::vector<int> level, level2;    // BFS level (two buffers: the pseudo-peripheral search keeps the best structure intact)
  std::vector<int> seen;             // BFS visit stamp
  std::atomic<int> next_id{1};
  std::atomic<int> next_stamp{1};
  explicit NDShared(int m) : tag(m), level(m, 0), level2(m, 0), seen(m, 0)
  {
    for (auto& t : tag)
    {
      t.store(0, std::memory_order_relaxed);
    }
  }
};

// ... and per-thread scratch
struct NDLocal
{
  std::vector<int> queue, queue2;
  std::vector<int> level_ptr, lp2;
};

// BFS restricted to nodes with tag == id, starting from root. Returns number of levels; fills
// `queue` with the visit order (first `count` entries) and `level` (node-indexed).
static int
bfs(const Graph& g, NDShared& sh, std::vector<int>& queue, std::vector<int>& level, int root, int id, int& count, std::vector<int>& level_ptr)
{
  const int stamp = sh.next_stamp.fetch_add(1, std::memory_order_relaxed);
  int head = 0, tail = 0;
  queue[tail++] = root;
  sh.seen[root] = stamp;
  level[root]   = 0;
  level_ptr.clear();
  level_ptr.push_back(0);
  int cur_level = 0;
  while (head < tail)
  {
    int v = queue[head];
    if (level[v] != cur_level)
    {
      cur_level = level[v];
      level_ptr.push_back(head);
    }
    ++head;
    for (int p = g.xadj[v]; p < g.xadj[v + 1]; ++p)
    {
      int u = g.adj[p];
      if (sh.tag[u].load(std::memory_order_relaxed) == id && sh.seen[u] != stamp)
      {
        sh.seen[u]    = stamp;
        level[u]      = cur_level + 1;
        queue[tail++] = u;
      }
    }
  }
  level_ptr.push_back(tail);
  count = tail;
  return (int)level_ptr.size() - 1;
}

struct Sub
{
  std::vector<int> nodes;
  int lo; // nodes occupy perm[lo, lo + nodes.size())
  bool connected;
};

// Processes one subproblem: writes final positions into perm, appends the children to `out`.
static void
nd_step(const Graph& g, NDShared& sh, NDLocal& loc, Sub sub, int leaf_size, std::vector<int>& perm, std::vector<Sub>& out)
{
  const int ns = (int)sub.nodes.size();
  if (ns == 0)
  {
    return;
  }
  const int id = sh.next_id.fetch_add(1, std::memory_order_relaxed);
  for (int v : sub.nodes)
  {
    sh.tag[v].store(id, std::memory_order_relaxed);
  }
  // loc.queue / sh.level / loc.level_ptr hold the level structure rooted at sub.nodes[0] over the whole subproblem
  // (left behind by the component scan when there is a single component)
  bool have_first = false;

  if (!sub.connected)
  {
    // split into connected components; each gets a consecutive range. Nodes already taken get tag -id.
    // Small components (isolated nodes are common: thousands per subproblem) are written out on the spot.
    int lo = sub.lo;
    for (int v : sub.nodes)
    {
      if (sh.tag[v].load(std::memory_order_relaxed) != id)
      {
        continue;
      }
      int cnt;
      bfs(g, sh, loc.queue, sh.level, v, id, cnt, loc.level_ptr);
      if (cnt == ns)
      {
        // a single component: continue below with its level structure (nodes in BFS order from the same first node)
        sub.nodes.assign(loc.queue.begin(), loc.queue.begin() + cnt);
        have_first = true;
        break;
      }
      for (int i = 0; i < cnt; ++i)
      {
        sh.tag[loc.queue[i]].store(-id, std::memory_order_relaxed);
      }
      if (cnt <= leaf_size)
      {
        for (int i = 0; i < cnt; ++i)
        {
          perm[lo + i] = loc.queue[i]; // BFS order
        }
      }
      else
      {
        Sub c;
        c.nodes.assign(loc.queue.begin(), loc.queue.begin() + cnt);
        c.connected = true;
        c.lo        = lo;
        out.push_back(std::move(c));
      }
      lo += cnt;
    }
    if (!have_first)
    {
      return;
    }
  }

  // connected subgraph
  int cnt  = ns;
  int nlev = (int)loc.level_ptr.size() - 1;
  if (!have_first)
  {
    nlev = bfs(g, sh, loc.queue, sh.level, sub.nodes[0], id, cnt, loc.level_ptr);
  }
  if (ns <= leaf_size)
  {
    for (int i = 0; i < cnt; ++i)
    {
      perm[sub.lo + i] = loc.queue[i];
    }
    return;
  }

  // pseudo-peripheral root. The candidate's level structure goes to the other set of buffers, so the best one so far
  // stays intact and never has to be rebuilt.
  std::vector<int>*q_cur = &loc.queue, *q_alt = &loc.queue2;
  std::vector<int>*l_cur = &sh.level, *l_alt = &sh.level2;
  std::vector<int>*lp_cur = &loc.level_ptr, *lp_alt = &loc.lp2;
  for (int iter = 0; iter < 3; ++iter)
  {
    // pick a minimum-degree node of the last level
    int best = -1, bestdeg = 0x7fffffff;
    for (int q = (*lp_cur)[nlev - 1]; q < (*lp_cur)[nlev]; ++q)
    {
      int v   = (*q_cur)[q];
      int deg = g.xadj[v + 1] - g.xadj[v];
      if (deg < bestdeg)
      {
        bestdeg = deg;
        best    = v;
      }
    }
    int cnt2;
    int nlev2 = bfs(g, sh, *q_alt, *l_alt, best, id, cnt2, *lp_alt);
    if (nlev2 > nlev)
    {
      nlev = nlev2;
      cnt  = cnt2;
      std::swap(q_cur, q_alt);
      std::swap(l_cur, l_alt);
      std::swap(lp_cur, lp_alt);
    }
    else
    {
      break;
    }
  }
  const std::vector<int>& queue     = *q_cur;
  const std::vector<int>& level     = *l_cur;
  const std::vector<int>& level_ptr = *lp_cur;

  if (nlev < 3)
  {
    // no interior level: cannot be separated by a level set; emit as one block
    for (int i = 0; i < cnt; ++i)
    {
      perm[sub.lo + i] = queue[i];
    }
    return;
  }

  // choose the separator level: smallest level among the balanced ones
  int s = -1;
  {
    const double lo_frac = 0.3;
    int best_size        = 0x7fffffff;
    int median_level     = 1;
    for (int l = 0; l < nlev; ++l)
    {
      if (level_ptr[l] <= ns / 2 && ns / 2 < level_ptr[l + 1])
      {
        median_level = l;
      }
    }
    median_level = std::min(std::max(median_level, 1), nlev - 2);
    for (int l = 1; l <= nlev - 2; ++l)
    {
      int below = level_ptr[l];
      int above = ns - level_ptr[l + 1];
      if (below >= lo_frac * ns && above >= lo_frac * ns)
      {
        int size = level_ptr[l + 1] - level_ptr[l];
        if (size < best_size || (size == best_size && std::abs(l - median_level) < std::abs(s - median_level)))
        {
          best_size = size;
          s         = l;
        }
      }
    }
    if (s < 0)
    {
      s = median_level;
    }
  }

  // thin the separator: level-s nodes without a neighbour in level s+1 join part A
  Sub A, B;
  std::vector<int> sep;
  A.nodes.reserve((size_t)level_ptr[s + 1]);
  B.nodes.reserve((size_t)(cnt - level_ptr[s + 1]));
  for (int q = 0; q < level_ptr[s]; ++q)
  {
    A.nodes.push_back(queue[q]);
  }
  for (int q = level_ptr[s]; q < level_ptr[s + 1]; ++q)
  {
    int v         = queue[q];
    bool touchesB = false;
    for (int p = g.xadj[v]; p < g.xadj[v + 1] && !touchesB; ++p)
    {
      int u = g.adj[p];
      if (sh.tag[u].load(std::memory_order_relaxed) == id && level[u] == s + 1)
      {
        touchesB = true;
      }
    }
    if (touchesB)
    {
      sep.push_back(v);
    }
    else
    {
      A.nodes.push_back(v);
    }
  }
  for (int q = level_ptr[s + 1]; q < cnt; ++q)
  {
    B.nodes.push_back(queue[q]);
  }
  // separator last
  const int hi = sub.lo + ns;
  for (size_t i = 0; i < sep.size(); ++i)
  {
    perm[hi - (int)sep.size() + (int)i] = sep[i];
  }
  A.lo        = sub.lo;
  A.connected = false; // (levels < s are connected through the BFS tree, but the thinned-in nodes may not be)
  B.lo        = sub.lo + (int)A.nodes.size();
  B.connected = false;
  out.push_back(std::move(A));
  out.push_back(std::move(B));
}

} // namespace

constexpr int ND_SHARE_MIN = 2048; // subproblems smaller than this stay with the thread that produced them

static void
nested_dissection(int m, const std::vector<int>& xadj, const std::vector<int>& adj, int leaf_size, std::vector<int>& perm)
{
  Graph g{m, xadj, adj};
  NDShared sh(m);
  perm.assign(m, -1);
  if (m == 0)
  {
    return;
  }
  // deterministic result: every subproblem's outcome depends only on its node list (ids/stamps are just
  // unique labels), whatever thread handles it
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Sub> stack;
  int in_flight = 0;
  {
    Sub root;
    root.nodes.resize(m);
    std::iota(root.nodes.begin(), root.nodes.end(), 0);
    root.lo        = 0;
    root.connected = false;
    stack.push_back(std::move(root));
  }
  unsigned hw        = std::thread::hardware_concurrency();
  int nthreads = m < 20000 ? 1 : (int)std::min<unsigned>(8, std::max<unsigned>(1, hw));
  if (const char* nt = std::getenv("B200_HOST_THREADS"))
  {
    nthreads = std::max(1, std::min(64, std::atoi(nt)));
  }
  if (const char* nt = std::getenv("B200_ND_THREADS")) // measurement knob (profiles/symbolic_threads.py)
  {
    nthreads = std::max(1, std::min(64, std::atoi(nt)));
  }
  std::atomic<int> tcount{0};
  auto worker = [&]() {
    NDLocal loc;
    loc.queue.assign(m, 0);
    loc.queue2.assign(m, 0);
    std::vector<Sub> out, local;
    long done = 0, nodes_done = 0;
    const int me = tcount.fetch_add(1);
    struct Report { long& d; long& n; int me; ~Report() { if (std::getenv("B200_SYM_TRACE")) std::fprintf(stderr, "[b200 nd] thread %d: %ld subproblems, %ld nodes\n", me, d, n); } } report{done, nodes_done, me};
    for (;;)
    {
      Sub sub;
      {
        std::unique_lock<std::mutex> lock(mu);
        cv.wait(lock, [&] { return !stack.empty() || in_flight == 0; });
        if (stack.empty())
        {
          return; // nothing queued and nothing running: done
        }
        sub = std::move(stack.back());
        stack.pop_back();
        ++in_flight;
      }
      // Only large children go back to the shared stack; the rest of the subtree is finished by this thread from a
      // private stack. (Sharing every child cost two lock hand-offs and a wake-up of all sleepers per subproblem --
      // ten thousand of them at config 2 -- which ate the whole gain of the threads.)
      local.clear();
      local.push_back(std::move(sub));
      while (!local.empty())
      {
        Sub cur = std::move(local.back());
        local.pop_back();
        out.clear();
        ++done;
        nodes_done += (long)cur.nodes.size();
        nd_step(g, sh, loc, std::move(cur), leaf_size, perm, out);
        bool shared = false;
        for (auto& c : out)
        {
          if ((int)c.nodes.size() >= ND_SHARE_MIN && nthreads > 1)
          {
            std::lock_guard<std::mutex> lock(mu);
            stack.push_back(std::move(c));
            shared = true;
          }
          else
          {
            local.push_back(std::move(c));
          }
        }
        if (shared)
        {
          cv.notify_all();
        }
      }
      {
        std::lock_guard<std::mutex> lock(mu);
        --in_flight;
      }
      cv.notify_all();
    }
  };
  if (nthreads == 1)
  {
    worker();
  }
  else
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
    {
      pool.emplace_back(worker);
    }
    for (auto& t : pool)
    {
      t.join();
    }
  }
}

// ---------------------------------------------------------------------------------------

// fn(part, lo, hi) over nparts contiguous shares of [0, n), one host thread each (part 0 on the caller's thread).
template <class F>
static void
parallel_ranges(i64 n, int nparts, F&& fn)
{
  nparts = (int)std::max<i64>(1, std::min<i64>(nparts, n));
  if (nparts == 1)
  {
    fn(0, (i64)0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 1; t < nparts; ++t)
  {
    pool.emplace_back([&fn, n, nparts, t]() { fn(t, n * t / nparts, n * (t + 1) / nparts); });
  }
  fn(0, (i64)0, n / nparts);
  for (auto& th : pool)
  {
    th.join();
  }
}

static int
host_threads(i64 work, i64 min_work)
{
  if (const char* ht = std::getenv("B200_HOST_THREADS")) // tests: the plan must not depend on the number of threads
  {
    return std::max(1, std::min(64, std::atoi(ht)));
  }
  return work < min_work ? 1 : (int)std::min<unsigned>(8, std::max<unsigned>(1, std::thread::hardware_concurrency()));
}

static int
fail(std::string& err, int code, const std::string& msg)
{
  err = msg;
  return code;
}

int
analyze(int n, int nnz, const int* colptr, const int* rowidx, const double* val, int lower_only, Plan& P, std::string& err)
{
  auto t0 = std::chrono::steady_clock::now();
  auto t_last = t0;
  const bool trace = std::getenv("B200_SYM_TRACE") != nullptr;
  auto tick = [&](const char* what) {
    if (trace)
    {
      auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[b200 symbolic] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
      t_last = now;
    }
  };
  if (n < 0 || nnz < 0 || !colptr || (nnz > 0 && (!rowidx || !val)))
  {
    return fail(err, B200_ERR_ARG, "null or negative-sized input");
  }
  if (colptr[0] != 0 || colptr[n] != nnz)
  {
    return fail(err, B200_ERR_ARG, "colptr[0] must be 0 and colptr[n] == nnz");
  }
  if (!valid_csc_header(n, nnz, colptr, rowidx, val))
  {
    return fail(err, B200_ERR_ARG, "colptr not monotone or out of range");
  }
  P       = Plan();
  P.N     = n;
  P.nnzK  = 0;
  P.nnzK_input = nnz;
  P.pattern_hash = hash_pattern(n, nnz, colptr, rowidx, val, lower_only, &P.pattern_hash2);

  // ---- classification ----------------------------------------------------------------
  std::vector<int> diag_src(n, -1);
  for (int j = 0; j < n; ++j)
  {
    if (colptr[j + 1] < colptr[j])
    {
      return fail(err, B200_ERR_ARG, "colptr not monotone");
    }
    int prev = -1;
    for (int p = colptr[j]; p < colptr[j + 1]; ++p)
    {
      int i = rowidx[p];
      if (i < 0 || i >= n || i <= prev)
      {
        return fail(err, B200_ERR_ARG, "row indices must be in range and strictly increasing per column (mat.c:797-804)");
      }
      prev = i;
      if (i < j)
      {
        if (lower_only)
        {
          return fail(err, B200_ERR_ARG, "entry above the diagonal in a matrix passed as lower-triangular");
        }
        continue;
      }
      ++P.nnzK;
      if (i == j && val[p] != 0.)
      {
        diag_src[j] = p;
      }
    }
  }
  P.e_of_k.assign(n, -1);
  P.r_of_k.assign(n, -1);
  P.k_of_e.reserve((size_t)n);
  P.dE_src.reserve((size_t)n);
  P.k_of_r.reserve((size_t)n);
  // Which indices are eliminated in closed form (E): those with a non-zero diagonal entry, as long as they form an
  // independent set of K's graph and their elimination does not densify the Schur complement. What SLEQP builds
  // (standard_aug_jac.c:135-237: identity (1,1) block) makes every variable an E node. Two kinds of candidates stay in
  // the reduced system instead (R nodes with a non-zero diagonal; S is then symmetric quasi-definite instead of
  // negative definite, which the LDL^T below factors just the same):
  //   * a candidate coupled to an E node chosen before it (an off-diagonal entry inside the (1,1) block), and
  //   * a candidate whose column is so long that its clique in S alone would hold more than 8 x nnz(K) entries (a dense
  //     column of J).
  std::vector<char> elim(n, 0);
  {
    std::vector<int> deg(n, 0);
    for (int j = 0; j < n; ++j)
    {
      for (int p = colptr[j]; p < colptr[j + 1]; ++p)
      {
        const int i = rowidx[p];
        if (i > j)
        {
          ++deg[i];
          ++deg[j];
        }
      }
    }
    const double clique_cap = 8.0 * (double)std::max<long long>(P.nnzK, 1);
    std::vector<char> blocked(n, 0);
    for (int j = 0; j < n; ++j)
    {
      const bool dense = deg[j] > 128 && (double)deg[j] * deg[j] > clique_cap;
      if (diag_src[j] < 0 || blocked[j] || dense)
      {
        P.n_demoted += diag_src[j] >= 0;
        continue;
      }
      elim[j] = 1;
      for (int p = colptr[j]; p < colptr[j + 1]; ++p)
      {
        if (rowidx[p] > j)
        {
          blocked[rowidx[p]] = 1; // a neighbour of an E node cannot be one itself
        }
      }
    }
  }
  for (int j = 0; j < n; ++j)
  {
    if (elim[j])
    {
      P.e_of_k[j] = P.nE++;
      P.k_of_e.push_back(j);
      P.dE_src.push_back(diag_src[j]);
    }
    else
    {
      P.r_of_k[j] = P.m++;
      P.k_of_r.push_back(j);
    }
  }
  const int nE = P.nE, m = P.m;

  // ---- split tril(K) into A (R x E) and G (R x R) ---------------------------------------
  struct Ent
  {
    int r, c, src;
  };
  std::vector<Ent> Aent, Gent;
  Aent.reserve((size_t)P.nnzK);
  Gent.reserve((size_t)P.nnzK / 4 + (size_t)n);
  for (int j = 0; j < n; ++j)
  {
    for (int p = colptr[j]; p < colptr[j + 1]; ++p)
    {
      int i = rowidx[p];
      if (i < j)
      {
        continue;
      }
      if (i == j)
      {
        if (P.r_of_k[j] >= 0)
        {
          Gent.push_back({P.r_of_k[j], P.r_of_k[j], p}); // explicit zero diagonal
        }
        continue;
      }
      const bool jE = P.e_of_k[j] >= 0, iE = P.e_of_k[i] >= 0;
      if (jE && iE)
      {
        return fail(err,
                    B200_ERR_UNSUPPORTED,
                    "off-diagonal entry inside the (1,1) block: the B200 backend factors K = [D A^T; A G] with diagonal D "
                    "(what standard_aug_jac.c:160 builds)");
      }
      if (jE)
      {
        Aent.push_back({P.r_of_k[i], P.e_of_k[j], p});
      }
      else if (iE)
      {
        Aent.push_back({P.r_of_k[j], P.e_of_k[i], p});
      }
      else
      {
        int a = P.r_of_k[i], b = P.r_of_k[j];
        Gent.push_back({std::max(a, b), std::min(a, b), p});
      }
    }
  }
  // A by column (E), rows ascending
  {
    auto less = [](const Ent& x, const Ent& y) { return x.c != y.c ? x.c < y.c : x.r < y.r; };
    if (!std::is_sorted(Aent.begin(), Aent.end(), less)) // the reference layout (all A entries in variable columns) arrives sorted
    {
      std::sort(Aent.begin(), Aent.end(), less);
    }
  }
  const int nnzA = (int)Aent.size();
  P.Acsc_ptr.assign(nE + 1, 0);
  P.Acsc_row.resize(nnzA);
  P.Acsc_src.resize(nnzA);
  for (const Ent& e : Aent)
  {
    ++P.Acsc_ptr[e.c + 1];
  }
  for (int e = 0; e < nE; ++e)
  {
    P.Acsc_ptr[e + 1] += P.Acsc_ptr[e];
  }
  for (int q = 0; q < nnzA; ++q)
  {
    P.Acsc_row[q] = Aent[q].r;
    P.Acsc_src[q] = Aent[q].src;
  }
  // A by row (R)
  P.Acsr_ptr.assign(m + 1, 0);
  P.Acsr_col.resize(nnzA);
  P.Acsr_src.resize(nnzA);
  for (const Ent& e : Aent)
  {
    ++P.Acsr_ptr[e.r + 1];
  }
  for (int r = 0; r < m; ++r)
  {
    P.Acsr_ptr[r + 1] += P.Acsr_ptr[r];
  }
  {
    std::vector<int> fill(P.Acsr_ptr.begin(), P.Acsr_ptr.end() - 1);
    for (const Ent& e : Aent) // ascending (c, r) => columns ascending within each row
    {
      int q         = fill[e.r]++;
      P.Acsr_col[q] = e.c;
      P.Acsr_src[q] = e.src;
    }
  }
  // G symmetric by row
  {
    P.Gsym_ptr.assign(m + 1, 0);
    for (const Ent& e : Gent)
    {
      ++P.Gsym_ptr[e.r + 1];
      if (e.r != e.c)
      {
        ++P.Gsym_ptr[e.c + 1];
      }
    }
    for (int r = 0; r < m; ++r)
    {
      P.Gsym_ptr[r + 1] += P.Gsym_ptr[r];
    }
    P.Gsym_col.resize(P.Gsym_ptr[m]);
    P.Gsym_src.resize(P.Gsym_ptr[m]);
    std::vector<int> fill(P.Gsym_ptr.begin(), P.Gsym_ptr.end() - 1);
    for (const Ent& e : Gent)
    {
      int q         = fill[e.r]++;
      P.Gsym_col[q] = e.c;
      P.Gsym_src[q] = e.src;
      if (e.r != e.c)
      {
        q             = fill[e.c]++;
        P.Gsym_col[q] = e.r;
        P.Gsym_src[q] = e.src;
      }
    }
  }

  tick("classify + split");
  // ---- adjacency of S = G - A D^-1 A^T (original R labels) -----------------------------
  {
    double work = 0;
    for (int e = 0; e < nE; ++e)
    {
      double q = P.Acsc_ptr[e + 1] - P.Acsc_ptr[e];
      work += q * q;
    }
    if (work > 4e9)
    {
      return fail(err, B200_ERR_UNSUPPORTED, "a variable couples too many working-set rows: Schur pattern would exceed 4e9 entries");
    }
  }
  std::vector<int> xadj(m + 1, 0), adj;
  {
    std::vector<int> mark(m, -1);
    for (int r = 0; r < m; ++r)
    {
      mark[r] = r;
      for (int q = P.Acsr_ptr[r]; q < P.Acsr_ptr[r + 1]; ++q)
      {
        int e = P.Acsr_col[q];
        for (int s = P.Acsc_ptr[e]; s < P.Acsc_ptr[e + 1]; ++s)
        {
          int r2 = P.Acsc_row[s];
          if (mark[r2] != r)
          {
            mark[r2] = r;
            adj.push_back(r2);
          }
        }
      }
      for (int q = P.Gsym_ptr[r]; q < P.Gsym_ptr[r + 1]; ++q)
      {
        int r2 = P.Gsym_col[q];
        if (mark[r2] != r)
        {
          mark[r2] = r;
          adj.push_back(r2);
        }
      }
      xadj[r + 1] = (int)adj.size();
      if (adj.size() > (size_t)0x7ffffff0)
      {
        return fail(err, B200_ERR_UNSUPPORTED, "Schur pattern exceeds 2^31 entries");
      }
    }
  }

  tick("S pattern");
  // ---- fill-reducing ordering -----------------------------------------------------------
  std::vector<int> perm0;
  nested_dissection(m, xadj, adj, /*leaf_size=*/24, perm0);
  std::vector<int> pinv0(m);
  for (int k = 0; k < m; ++k)
  {
    pinv0[perm0[k]] = k;
  }

  tick("nested dissection");
  // ---- elimination tree (Liu) -------------------------------------------------------------
  std::vector<int> parent0(m, -1);
  {
    std::vector<int> anc(m, -1);
    for (int k = 0; k < m; ++k)
    {
      int v = perm0[k];
      for (int p = xadj[v]; p < xadj[v + 1]; ++p)
      {
        int i = pinv0[adj[p]];
        while (i != -1 && i < k)
        {
          int inext = anc[i];
          anc[i]    = k;
          if (inext == -1)
          {
            parent0[i] = k;
          }
          i = inext;
        }
      }
    }
  }
  // ---- postorder (labels L1) ------------------------------------------------------------------
  std::vector<int> post1(m);
  {
    std::vector<int> head(m, -1), next(m, -1), stk;
    for (int j = m - 1; j >= 0; --j)
    {
      if (parent0[j] != -1)
      {
        next[j]          = head[parent0[j]];
        head[parent0[j]] = j;
      }
    }
    int k = 0;
    for (int root = 0; root < m; ++root)
    {
      if (parent0[root] != -1)
      {
        continue;
      }
      stk.push_back(root);
      while (!stk.empty())
      {
        int v = stk.back();
        int c = head[v];
        if (c == -1)
        {
          post1[k++] = v;
          stk.pop_back();
        }
        else
        {
          head[v] = next[c];
          stk.push_back(c);
        }
      }
    }
  }
  std::vector<int> parent1(m, -1), postinv1(m);
  for (int k = 0; k < m; ++k)
  {
    postinv1[post1[k]] = k;
  }
  for (int k = 0; k < m; ++k)
  {
    const int p0 = parent0[post1[k]];
    parent1[k]   = p0 == -1 ? -1 : postinv1[p0];
  }
  tick("etree + postorder");
  // ---- column counts (Gilbert, Ng, Peyton 1994) in the postorder labels; the neighbours are relabelled on the fly
  // (the algorithm does not need them sorted) ----------------------------------------------------------------
  std::vector<int> cc1(m, 0);
  {
    const std::vector<int>& parent = parent1;
    std::vector<int> first(m, -1), maxfirst(m, -1), prevleaf(m, -1), ancestor(m);
    std::vector<int>& delta = cc1;
    for (int k = 0; k < m; ++k)
    {
      int j    = k;
      delta[j] = (first[j] == -1) ? 1 : 0;
      for (; j != -1 && first[j] == -1; j = parent[j])
      {
        first[j] = k;
      }
    }
    std::iota(ancestor.begin(), ancestor.end(), 0);
    for (int j = 0; j < m; ++j)
    {
      if (parent[j] != -1)
      {
        --delta[parent[j]];
      }
      const int v = perm0[post1[j]];
      for (int p = xadj[v]; p < xadj[v + 1]; ++p)
      {
        int i = postinv1[pinv0[adj[p]]];
        if (i <= j || first[j] <= maxfirst[i])
        {
          continue;
        }
        maxfirst[i] = first[j];
        int jprev   = prevleaf[i];
        prevleaf[i] = j;
        if (jprev == -1)
        {
          ++delta[j];
        }
        else
        {
          int q = jprev;
          while (q != ancestor[q])
          {
            q = ancestor[q];
          }
          for (int s = jprev; s != q;)
          {
            int sp      = ancestor[s];
            ancestor[s] = q;
            s           = sp;
          }
          ++delta[j];
          --delta[q];
        }
      }
      if (parent[j] != -1)
      {
        ancestor[j] = parent[j];
      }
    }
    for (int j = 0; j < m; ++j)
    {
      if (parent[j] != -1)
      {
        delta[parent[j]] += delta[j];
      }
    }
  }
  tick("column counts");
  // ---- sparse subtrees: maximal complete subtrees with few entries per column (plan.hpp), in generations: what is
  // left of the tree once the subtrees of one generation are cut off may again end in sparse subtrees (the separators
  // above the chains of config 3), whose children are the subtrees of the generations before ---------------------
  std::vector<int> sst_root(m, -1); // root (L1 label) of the sparse subtree a column belongs to
  std::vector<int> sst_gen(m, -1);  // generation of the subtree rooted at a column
  {
    const char* e   = std::getenv("B200_SST"); // 0: everything through the dense supernodal path (measurements, tests)
    const bool on   = !(e && e[0] == '0');
    std::vector<int> full(m, 1); // size of the whole subtree (a contiguous range of postorder labels)
    for (int j = 0; j < m; ++j)
    {
      if (parent1[j] != -1)
      {
        full[parent1[j]] += full[j];
      }
    }
    std::vector<int> size(m);
    std::vector<i64> nnz(m);
    std::vector<char> ok(m);
    for (int gen = 0; on && gen < 16; ++gen)
    {
      // over the columns not yet in a subtree: own columns / entries below every node, limits
      std::fill(size.begin(), size.end(), 0);
      std::fill(nnz.begin(), nnz.end(), 0);
      std::fill(ok.begin(), ok.end(), 1);
      for (int j = 0; j < m; ++j) // children before parents
      {
        if (sst_root[j] == -1)
        {
          size[j] += 1;
          nnz[j] += cc1[j];
          ok[j] = ok[j] && size[j] <= SST_MAX_COLS && nnz[j] <= SST_MAX_NNZ && nnz[j] <= (i64)SST_MAX_AVG * size[j];
        }
        const int p = parent1[j];
        if (p != -1)
        {
          size[p] += size[j];
          nnz[p] += nnz[j];
          ok[p] = ok[p] && ok[j];
        }
      }
      // the highest eligible nodes, parents first
      const int min_cols = gen == 0 ? SST_MIN_COLS : 1;
      bool any           = false;
      for (int j = m - 1; j >= 0; --j)
      {
        if (sst_root[j] != -1 || !ok[j] || cc1[j] - 1 > SST_MAX_TAIL || size[j] < min_cols)
        {
          continue;
        }
        for (int c = j - full[j] + 1; c <= j; ++c)
        {
          if (sst_root[c] == -1)
          {
            sst_root[c] = j;
          }
        }
        sst_gen[j] = gen;
        any        = true;
      }
      if (!any)
      {
        break;
      }
    }
  }
  // ---- supernodes, step 1: fundamental supernodes + relaxed chain amalgamation (in the postorder labels) ----------
  // rn_last[j]: last column of the chain supernode ("R-node") column j belongs to. Chains only: column j joins the
  // supernode of j + 1 when j + 1 is its parent and the merge is fundamental (identical structure) or the supernode
  // is already at least NB wide and the merge adds at most 20 % explicit zeros.
  std::vector<int> rn_last(m, 0);
  if (m > 0)
  {
    int l      = m - 1;
    i64 k      = 1;
    i64 actual = cc1[l];
    rn_last[l] = l;
    for (int j = m - 2; j >= 0; --j)
    {
      bool jn = false;
      if (sst_root[j] != -1 || sst_root[j + 1] != -1)
      {
        jn = false; // the columns of a sparse subtree are grouped below (they need not be consecutive here)
      }
      else if (parent1[j] == j + 1)
      {
        const i64 r       = cc1[l] - 1;
        const i64 kn      = k + 1;
        const i64 stored  = kn * (kn + r) - kn * (kn - 1) / 2;
        const i64 act     = actual + cc1[j];
        const bool fundam = cc1[j] == cc1[j + 1] + 1;
        // the relaxed rule is for chains that are already wider than one panel step (big separators absorbing what
        // is nearly identical below them); everything smaller is left to the latency-driven grouping of step 2
        jn = fundam || (k >= NB && (double)(stored - act) <= 0.2 * (double)stored);
      }
      if (jn)
      {
        ++k;
        actual += cc1[j];
      }
      else
      {
        l      = j;
        k      = 1;
        actual = cc1[j];
      }
      rn_last[j] = l;
    }
  }
  // ---- supernodes, step 2: latency-driven amalgamation into groups of at most `cap` columns ------------------------
  // A supernode of up to NB columns costs ONE panel step of the factorization and ONE dependency level of the sweeps
  // whatever it holds, and every level of the supernodal tree costs a hand-off of several microseconds (profiles/):
  // so the tree of chain supernodes is collapsed bottom-up. Every node keeps an "open set" of columns below it that
  // are not yet part of a finished group; when node + open sets of its children exceed the cap, the largest open sets
  // are closed -- each becomes one dense supernode, its columns made contiguous by the final relabelling -- until the
  // rest fits. On a path graph (config 3) this turns log2(m) levels of one-column separators into log32(m) levels;
  // chain supernodes wider than the cap (the separators of 2D/3D problems) stay what they were.
  std::vector<int> grp1(m, -1); // group root (L1 label of its last column) of every column
  std::vector<int> post(m);     // final order (ND labels), groups contiguous, child groups before parent groups
  {
    int cap = NB;
    if (const char* gc = std::getenv("B200_GROUP_CAP"))
    {
      cap = std::max(1, std::atoi(gc));
    }
    // tree of R-nodes, identified by their last column; children lists in ascending order
    std::vector<int> rpar(m, -1), nchild(m, 0), width(m, 0);
    for (int j = 0; j < m; ++j)
    {
      ++width[rn_last[j]];
    }
    for (int l = 0; l < m; ++l)
    {
      if (rn_last[l] == l && parent1[l] != -1)
      {
        rpar[l] = rn_last[parent1[l]];
        ++nchild[rpar[l]];
      }
    }
    std::vector<int> cptr(m + 1, 0), cidx(std::max(m, 1));
    for (int j = 0; j < m; ++j)
    {
      cptr[j + 1] = cptr[j] + nchild[j];
    }
    {
      std::vector<int> fill(cptr.begin(), cptr.end() - 1);
      for (int l = 0; l < m; ++l)
      {
        if (rn_last[l] == l && rpar[l] != -1)
        {
          cidx[fill[rpar[l]]++] = l;
        }
      }
    }
    std::vector<int> open(m, 0);
    std::vector<char> closed(m, 0);
    std::vector<std::pair<int, int>> kids;
    for (int l = 0; l < m; ++l) // children before parents (postorder)
    {
      if (rn_last[l] != l)
      {
        continue;
      }
      kids.clear();
      int total = width[l];
      for (int q = cptr[l]; q < cptr[l + 1]; ++q)
      {
        const int c = cidx[q];
        if (open[c] > 0)
        {
          kids.push_back({open[c], c});
          total += open[c];
        }
      }
      std::sort(kids.begin(), kids.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first != y.first ? x.first > y.first : x.second < y.second; });
      for (size_t q = 0; q < kids.size() && total > cap; ++q)
      {
        closed[kids[q].second] = 1;
        total -= kids[q].first;
        open[kids[q].second] = 0;
      }
      open[l] = total;
      if (rpar[l] == -1)
      {
        closed[l] = 1;
      }
      if (sst_root[l] != -1) // the columns of a sparse subtree form their own group, nothing else merges with them
      {
        closed[l] = 1;
        open[l]   = 0;
      }
    }
    // group of an R-node: itself when closed, else its parent's; then of every column
    std::vector<int> rgrp(m, -1);
    for (int l = m - 1; l >= 0; --l) // parents before children
    {
      if (rn_last[l] == l)
      {
        rgrp[l] = closed[l] ? l : rgrp[rpar[l]];
      }
    }
    for (int j = 0; j < m; ++j)
    {
      grp1[j] = sst_root[j] != -1 ? sst_root[j] : rgrp[rn_last[j]];
    }
    // order: postorder over the tree of groups (children by ascending root), the columns of a group ascending
    std::vector<int> gcount(m + 1, 0), gnodes(std::max(m, 1));
    for (int j = 0; j < m; ++j)
    {
      ++gcount[grp1[j] + 1];
    }
    for (int g = 0; g < m; ++g)
    {
      gcount[g + 1] += gcount[g];
    }
    {
      std::vector<int> fill(gcount.begin(), gcount.end() - 1);
      for (int j = 0; j < m; ++j)
      {
        gnodes[fill[grp1[j]]++] = j;
      }
    }
    // child groups of a group: the closed children of its columns, i.e. every group root r != tree root hangs below
    // the group of its parent
    std::vector<int> ghead(m, -1), gnext(m, -1);
    for (int r = m - 1; r >= 0; --r)
    {
      if (grp1[r] == r && parent1[r] != -1) // r is the root (last column) of its group
      {
        const int pg = grp1[parent1[r]];
        gnext[r]     = ghead[pg];
        ghead[pg]    = r;
      }
    }
    std::vector<int> stk;
    int k = 0;
    for (int root = 0; root < m; ++root)
    {
      if (parent1[root] != -1)
      {
        continue;
      }
      stk.push_back(root);
      while (!stk.empty())
      {
        const int g = stk.back();
        const int c = ghead[g];
        if (c == -1)
        {
          for (int q = gcount[g]; q < gcount[g + 1]; ++q)
          {
            post[k++] = gnodes[q]; // L1 label for now
          }
          stk.pop_back();
        }
        else
        {
          ghead[g] = gnext[c];
          stk.push_back(c);
        }
      }
    }
  }
  // final labels: position k holds L1 label post[k]; compose with the postorder to ND labels
  std::vector<int> cc_final(m), grp_final(m);
  std::vector<int> sst_final(m, -1); // generation of the sparse subtree a column belongs to (-1: none)
  {
    std::vector<int> l1_to_final(m);
    for (int k = 0; k < m; ++k)
    {
      l1_to_final[post[k]] = k;
    }
    for (int k = 0; k < m; ++k)
    {
      cc_final[k]  = cc1[post[k]];
      grp_final[k] = l1_to_final[grp1[post[k]]];
      sst_final[k] = sst_root[post[k]] != -1 ? sst_gen[sst_root[post[k]]] : -1;
    }
    for (int k = 0; k < m; ++k)
    {
      post[k] = post1[post[k]];
    }
  }
  std::vector<int>().swap(post1);
  std::vector<int>().swap(parent1);
  std::vector<int>().swap(postinv1);
  std::vector<int>().swap(cc1);
  std::vector<int>().swap(grp1);
  tick("amalgamation groups + order");
  P.perm.resize(m);
  P.pinv.resize(m);
  P.parent.assign(m, -1);
  {
    std::vector<int> postinv(m);
    for (int k = 0; k < m; ++k)
    {
      postinv[post[k]] = k;
    }
    for (int k = 0; k < m; ++k)
    {
      P.perm[k]         = perm0[post[k]];
      P.pinv[P.perm[k]] = k;
      int p0            = parent0[post[k]];
      P.parent[k]       = p0 == -1 ? -1 : postinv[p0];
    }
  }
  // adjacency in new labels
  std::vector<int> xadj2(m + 1, 0), adj2(adj.size());
  for (int k = 0; k < m; ++k)
  {
    int v        = P.perm[k];
    xadj2[k + 1] = xadj2[k] + (xadj[v + 1] - xadj[v]);
  }
  parallel_ranges(m, host_threads((i64)adj.size(), 500000), [&](int, i64 klo, i64 khi) {
    for (int k = (int)klo; k < (int)khi; ++k)
    {
      int v = P.perm[k];
      int o = xadj2[k];
      for (int p = xadj[v]; p < xadj[v + 1]; ++p)
      {
        adj2[o++] = P.pinv[adj[p]];
      }
      std::sort(adj2.begin() + xadj2[k], adj2.begin() + xadj2[k + 1]);
    }
  });
  std::vector<int>().swap(adj);

  tick("relabel");
  const std::vector<int>& parent = P.parent;
  P.colcount = cc_final;
  const std::vector<int>& cc = P.colcount;

  // ---- supernodes = the amalgamation groups (contiguous in the final labels, root column last) ---------------------
  std::vector<char> join(std::max(m, 1), 0); // join[j]: j and j+1 share a supernode
  for (int j = 0; j + 1 < m; ++j)
  {
    join[j] = grp_final[j] == grp_final[j + 1];
  }
  P.sn_of_col.assign(m, 0);
  P.sn_first.clear();
  for (int j = 0; j < m; ++j)
  {
    if (j == 0 || !join[j - 1])
    {
      P.sn_first.push_back(j);
    }
    P.sn_of_col[j] = (int)P.sn_first.size() - 1;
  }
  P.nsuper = (int)P.sn_first.size();
  P.sn_first.push_back(m);
  const int ns = P.nsuper;

  // supernodal tree
  P.sn_parent.assign(ns, -1);
  for (int T = 0; T < ns; ++T)
  {
    int l = P.sn_first[T + 1] - 1;
    if (parent[l] != -1)
    {
      P.sn_parent[T] = P.sn_of_col[parent[l]];
    }
  }
  P.child_ptr.assign(ns + 1, 0);
  for (int T = 0; T < ns; ++T)
  {
    if (P.sn_parent[T] >= 0)
    {
      ++P.child_ptr[P.sn_parent[T] + 1];
    }
  }
  for (int T = 0; T < ns; ++T)
  {
    P.child_ptr[T + 1] += P.child_ptr[T];
  }
  P.child_idx.resize(P.child_ptr[ns]);
  {
    std::vector<int> fill(P.child_ptr.begin(), P.child_ptr.end() - 1);
    for (int T = 0; T < ns; ++T)
    {
      if (P.sn_parent[T] >= 0)
      {
        P.child_idx[fill[P.sn_parent[T]]++] = T;
      }
    }
  }

  tick("supernodes");
  // ---- row structures ---------------------------------------------------------------------------------
  P.Rptr.assign(ns + 1, 0);
  P.Ridx.clear();
  {
    std::vector<int> mark(m, -1);
    std::vector<int> rows;
    for (int T = 0; T < ns; ++T)
    {
      const int f = P.sn_first[T], l = P.sn_first[T + 1] - 1;
      rows.clear();
      for (int j = f; j <= l; ++j)
      {
        for (int p = xadj2[j]; p < xadj2[j + 1]; ++p)
        {
          int i = adj2[p];
          if (i > l && mark[i] != T)
          {
            mark[i] = T;
            rows.push_back(i);
          }
        }
      }
      for (int q = P.child_ptr[T]; q < P.child_ptr[T + 1]; ++q)
      {
        int c = P.child_idx[q];
        for (i64 p = P.Rptr[c]; p < P.Rptr[c + 1]; ++p)
        {
          int i = P.Ridx[p];
          if (i > l && mark[i] != T)
          {
            mark[i] = T;
            rows.push_back(i);
          }
        }
      }
      std::sort(rows.begin(), rows.end());
      if ((int)rows.size() != cc[l] - 1)
      {
        return fail(err, B200_ERR_ARG, "internal: supernode row structure disagrees with the column count");
      }
      P.Ridx.insert(P.Ridx.end(), rows.begin(), rows.end());
      P.Rptr[T + 1] = (i64)P.Ridx.size();
    }
  }
  // relative indices into the parent's front
  P.rel.assign(P.Ridx.size(), -1);
  {
    std::vector<int> pos(m, -1);
    for (int T = 0; T < ns; ++T)
    {
      if (P.child_ptr[T] == P.child_ptr[T + 1])
      {
        continue;
      }
      const int f = P.sn_first[T], k = P.sn_first[T + 1] - f;
      for (int c = 0; c < k; ++c)
      {
        pos[f + c] = c;
      }
      for (i64 p = P.Rptr[T]; p < P.Rptr[T + 1]; ++p)
      {
        pos[P.Ridx[p]] = k + (int)(p - P.Rptr[T]);
      }
      for (int q = P.child_ptr[T]; q < P.child_ptr[T + 1]; ++q)
      {
        int c = P.child_idx[q];
        for (i64 p = P.Rptr[c]; p < P.Rptr[c + 1]; ++p)
        {
          P.rel[p] = pos[P.Ridx[p]];
          if (P.rel[p] < 0)
          {
            return fail(err, B200_ERR_ARG, "internal: child update row missing from the parent front");
          }
        }
      }
      // pos entries are overwritten by later parents; stale values are never read because a
      // child's rows are always a subset of its parent's front (checked above via rel >= 0 on a
      // freshly reset map would be stricter; reset to be safe)
      for (int c = 0; c < k; ++c)
      {
        pos[f + c] = -1;
      }
      for (i64 p = P.Rptr[T]; p < P.Rptr[T + 1]; ++p)
      {
        pos[P.Ridx[p]] = -1;
      }
    }
  }

  tick("row structures + rel");
  // ---- sparse subtrees: exact column structures, chains ("segments") and their levels, child assembly maps ---------
  // Device layout per subtree (sst.cu): values in the panel buffer (compact, column by column, diagonal first) and
  // ONE blob of 16-bit indices [level pointers over segments | segment starts | segment lengths | column pointers |
  // front-local rows], every part padded to 16 bytes. A segment is a maximal single-child path of the subtree's own
  // elimination tree (consecutive columns): one thread walks it sequentially, segments of one level run in parallel.
  P.sn_sparse.assign(ns, 0);
  std::vector<i64> sst_nnz(ns, 0);       // entries of a sparse subtree (its share of the panel buffer)
  std::vector<int> sst_index(ns, -1);    // position in P.sst
  std::vector<int> sst_col_of(m, -1);    // for the columns of sparse subtrees: offset of their column pointer in sst_colptr
  std::vector<int> sst_ea_child;         // per assembly entry: the child supernode (until its workspace offset is known)
  for (int T = 0; T < ns; ++T)
  {
    P.sn_sparse[T] = sst_final[P.sn_first[T]] >= 0;
  }
  auto pad8 = [](std::vector<unsigned short>& v) { v.resize((v.size() + 7) & ~(size_t)7, 0); };
  for (int T = 0; T < ns; ++T)
  {
    const int f = P.sn_first[T], l = P.sn_first[T + 1] - 1, k = l - f + 1;
    if (!P.sn_sparse[T])
    {
      continue;
    }
    const int* rows_T = P.Ridx.data() + P.Rptr[T];
    const int r       = (int)(P.Rptr[T + 1] - P.Rptr[T]);
    SstMeta M         = SstMeta();
    M.sn      = T;
    M.first   = f;
    M.k       = k;
    M.r       = r;
    M.Rptr    = (int)P.Rptr[T];
    M.gen     = sst_final[f];
    M.signal     = P.sn_parent[T];
    M.parent_sst = (P.sn_parent[T] >= 0 && P.sn_sparse[P.sn_parent[T]]) ? P.sn_parent[T] : -1;
    M.nchild     = P.child_ptr[T + 1] - P.child_ptr[T];
    // front-local index of a row (new labels): a column of the subtree or one of its update rows
    auto local = [&](int i) -> int {
      if (i <= l)
      {
        return i >= f ? i - f : -1;
      }
      const int* it = std::lower_bound(rows_T, rows_T + r, i);
      return (it != rows_T + r && *it == i) ? k + (int)(it - rows_T) : -1;
    };
    // struct(j) = lower adjacency of j + the structures of its children without j; a child is a column of the subtree
    // or the root of a child subtree (whose structure is its update rows)
    std::vector<std::vector<int>> st((size_t)k);
    std::vector<int> nchild((size_t)k, 0); // children inside the subtree
    for (int q = P.child_ptr[T]; q < P.child_ptr[T + 1]; ++q)
    {
      const int c = P.child_idx[q];
      if (!P.sn_sparse[c])
      {
        return fail(err, B200_ERR_ARG, "internal: a sparse subtree has a dense child supernode");
      }
      const int pc = parent[P.sn_first[c + 1] - 1]; // the column of T the child hangs below
      for (i64 t = P.Rptr[c]; t < P.Rptr[c + 1]; ++t)
      {
        if (P.Ridx[t] != pc)
        {
          st[(size_t)(pc - f)].push_back(P.Ridx[t]);
        }
      }
    }
    for (int j = f; j <= l; ++j)
    {
      std::vector<int>& cur = st[(size_t)(j - f)];
      for (int p = xadj2[j]; p < xadj2[j + 1]; ++p)
      {
        if (adj2[p] > j)
        {
          cur.push_back(adj2[p]);
        }
      }
      std::sort(cur.begin(), cur.end());
      cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
      if ((int)cur.size() != cc[j] - 1)
      {
        return fail(err, B200_ERR_ARG, "internal: column structure of a sparse subtree disagrees with the column count");
      }
      const int par = parent[j];
      if (par != -1 && par <= l)
      {
        std::vector<int>& up = st[(size_t)(par - f)];
        for (int i : cur)
        {
          if (i != par)
          {
            up.push_back(i);
          }
        }
        ++nchild[(size_t)(par - f)];
      }
    }
    // column pointers and front-local rows (diagonal first)
    std::vector<int> colptr((size_t)k + 1, 0), rowloc;
    for (int j = f; j <= l; ++j)
    {
      colptr[(size_t)(j - f)] = (int)rowloc.size();
      rowloc.push_back(j - f);
      for (int i : st[(size_t)(j - f)])
      {
        const int loc = local(i);
        if (loc < 0)
        {
          return fail(err, B200_ERR_ARG, "internal: row of a sparse subtree missing from its update rows");
        }
        rowloc.push_back(loc);
      }
    }
    const int nnz     = (int)rowloc.size();
    colptr[(size_t)k] = nnz;
    // segments: column c continues the segment of c - 1 iff c - 1 is its only child inside the subtree
    std::vector<int> seg_start, seg_len, seg_of((size_t)k, 0), seg_lev;
    for (int c = 0; c < k; ++c)
    {
      const bool cont = c > 0 && parent[f + c - 1] == f + c && nchild[(size_t)c] == 1;
      if (!cont)
      {
        seg_start.push_back(c);
        seg_len.push_back(0);
        seg_lev.push_back(0);
      }
      seg_of[(size_t)c] = (int)seg_start.size() - 1;
      ++seg_len.back();
    }
    const int nseg = (int)seg_start.size();
    int nslev      = 0;
    for (int q = 0; q < nseg; ++q) // children before parents: levels of the child segments are final
    {
      const int top = f + seg_start[(size_t)q] + seg_len[(size_t)q] - 1;
      const int par = parent[top];
      nslev         = std::max(nslev, seg_lev[(size_t)q] + 1);
      if (par != -1 && par <= l)
      {
        int& pl = seg_lev[(size_t)seg_of[(size_t)(par - f)]];
        pl      = std::max(pl, seg_lev[(size_t)q] + 1);
      }
    }
    std::vector<int> cnt((size_t)nslev + 1, 0);
    for (int q = 0; q < nseg; ++q)
    {
      ++cnt[(size_t)seg_lev[(size_t)q] + 1];
    }
    for (int q = 0; q < nslev; ++q)
    {
      cnt[(size_t)q + 1] += cnt[(size_t)q];
    }
    std::vector<int> order((size_t)nseg);
    {
      std::vector<int> fillp(cnt.begin(), cnt.end() - 1);
      for (int q = 0; q < nseg; ++q)
      {
        order[(size_t)fillp[(size_t)seg_lev[(size_t)q]]++] = q;
      }
    }
    // the blob
    pad8(P.sst_blob);
    M.blob = (int)P.sst_blob.size();
    for (int q = 0; q <= nslev; ++q)
    {
      P.sst_blob.push_back((unsigned short)cnt[(size_t)q]);
    }
    pad8(P.sst_blob);
    M.o_segstart = (int)P.sst_blob.size() - M.blob;
    for (int q : order)
    {
      P.sst_blob.push_back((unsigned short)seg_start[(size_t)q]);
    }
    pad8(P.sst_blob);
    M.o_seglen = (int)P.sst_blob.size() - M.blob;
    for (int q : order)
    {
      P.sst_blob.push_back((unsigned short)seg_len[(size_t)q]);
    }
    pad8(P.sst_blob);
    M.o_colptr = (int)P.sst_blob.size() - M.blob;
    for (int c = 0; c <= k; ++c)
    {
      P.sst_blob.push_back((unsigned short)colptr[(size_t)c]);
    }
    pad8(P.sst_blob);
    M.o_rows = (int)P.sst_blob.size() - M.blob;
    for (int v : rowloc)
    {
      P.sst_blob.push_back((unsigned short)v);
    }
    pad8(P.sst_blob);
    {
      // the same entries by row (diagonal left out), columns ascending: what a column gathers in the forward sweep
      std::vector<int> rptr((size_t)(k + r) + 1, 0);
      for (int c = 0; c < k; ++c)
      {
        for (int q = colptr[(size_t)c] + 1; q < colptr[(size_t)c + 1]; ++q)
        {
          ++rptr[(size_t)rowloc[(size_t)q] + 1];
        }
      }
      for (int i = 0; i < k + r; ++i)
      {
        rptr[(size_t)i + 1] += rptr[(size_t)i];
      }
      std::vector<int> rcol((size_t)rptr[(size_t)(k + r)]), rpos(rcol.size()), fillp(rptr.begin(), rptr.end() - 1);
      for (int c = 0; c < k; ++c) // columns ascending => ascending inside every row
      {
        for (int q = colptr[(size_t)c] + 1; q < colptr[(size_t)c + 1]; ++q)
        {
          const int o      = fillp[(size_t)rowloc[(size_t)q]]++;
          rcol[(size_t)o] = c;
          rpos[(size_t)o] = q;
        }
      }
      M.o_rowptr = (int)P.sst_blob.size() - M.blob;
      for (int v : rptr)
      {
        P.sst_blob.push_back((unsigned short)v);
      }
      pad8(P.sst_blob);
      M.o_rcol = (int)P.sst_blob.size() - M.blob;
      for (int v : rcol)
      {
        P.sst_blob.push_back((unsigned short)v);
      }
      pad8(P.sst_blob);
      M.o_rpos = (int)P.sst_blob.size() - M.blob;
      for (int v : rpos)
      {
        P.sst_blob.push_back((unsigned short)v);
      }
      pad8(P.sst_blob);
    }
    M.blob_len16 = ((int)P.sst_blob.size() - M.blob) / 8;
    M.nslev      = nslev;
    M.nseg       = nseg;
    M.nnz        = nnz;
    // int copies for the host side (assembly destinations below, emulation in the tests)
    for (int c = 0; c < k; ++c)
    {
      sst_col_of[f + c] = (int)P.sst_colptr.size() + c;
    }
    M.col_ptr = (int)P.sst_colptr.size();
    M.row_ptr = (int)P.sst_rows.size();
    P.sst_colptr.insert(P.sst_colptr.end(), colptr.begin(), colptr.end());
    P.sst_rows.insert(P.sst_rows.end(), rowloc.begin(), rowloc.end());
    // assembly of the children: where every entry (a >= b) of a child's update block goes -- a slot of the compact
    // values (>= 0) or an entry of this subtree's own update block (-1 - index)
    M.ea_begin = (int)P.sst_ea_dst.size();
    for (int q = P.child_ptr[T]; q < P.child_ptr[T + 1]; ++q)
    {
      const int c  = P.child_idx[q];
      const int rc = (int)(P.Rptr[c + 1] - P.Rptr[c]);
      for (int bcol = 0; bcol < rc; ++bcol)
      {
        const int ib = local(P.Ridx[P.Rptr[c] + bcol]);
        for (int arow = bcol; arow < rc; ++arow)
        {
          const int ia = local(P.Ridx[P.Rptr[c] + arow]);
          if (ia < 0 || ib < 0)
          {
            return fail(err, B200_ERR_ARG, "internal: update row of a child missing from the sparse subtree");
          }
          sst_ea_child.push_back(c);
          P.sst_ea_src.push_back(arow + (i64)bcol * rc); // + the child's workspace offset once that is allocated
          if (ib >= k)
          {
            P.sst_ea_dst.push_back(-1 - ((ia - k) + (ib - k) * r));
          }
          else
          {
            const int* b0 = rowloc.data() + colptr[(size_t)ib];
            const int* b1 = rowloc.data() + colptr[(size_t)ib + 1];
            const int* it = std::lower_bound(b0, b1, ia);
            if (it == b1 || *it != ia)
            {
              return fail(err, B200_ERR_ARG, "internal: fill of a child missing from the structure of the sparse subtree");
            }
            P.sst_ea_dst.push_back((int)(it - rowloc.data()));
          }
        }
      }
    }
    M.ea_end = (int)P.sst_ea_dst.size();
    sst_nnz[T]   = nnz;
    sst_index[T] = (int)P.sst.size();
    P.sst.push_back(M);
    {
      // dynamic shared memory of the subtree's CTA (sst.cu): values, front vector / update block, index blob
      const size_t vec = (size_t)std::max(k + r, r * r);
      const size_t b   = sizeof(double) * ((((size_t)nnz + 1) & ~(size_t)1) + ((vec + 1) & ~(size_t)1)) + 16 * (size_t)M.blob_len16;
      P.sst_smem_bytes = std::max(P.sst_smem_bytes, b + 64);
      if (P.sst_smem_bytes > SST_SMEM_LIMIT)
      {
        return fail(err, B200_ERR_ARG, "internal: a sparse subtree exceeds the shared memory of one CTA");
      }
    }
  }
  pad8(P.sst_blob);
  P.sst_blob.resize(P.sst_blob.size() + 8, 0);
  // launch order: generation by generation (a subtree after its children); P.sst is in supernode order, which already
  // puts children first, but one launch covers one generation
  {
    std::vector<int> ord(P.sst.size());
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return P.sst[(size_t)x].gen < P.sst[(size_t)y].gen; });
    std::vector<SstMeta> sorted;
    sorted.reserve(P.sst.size());
    for (int q : ord)
    {
      sst_index[P.sst[(size_t)q].sn] = (int)sorted.size();
      sorted.push_back(P.sst[(size_t)q]);
    }
    P.sst.swap(sorted);
    P.sst_gen_ptr.assign(1, 0);
    for (size_t q = 0; q < P.sst.size(); ++q)
    {
      while ((int)P.sst_gen_ptr.size() - 1 < P.sst[q].gen)
      {
        P.sst_gen_ptr.push_back((int)q);
      }
    }
    P.sst_gen_ptr.push_back((int)P.sst.size());
    if (P.sst.empty())
    {
      P.sst_gen_ptr.assign(1, 0);
    }
  }
  tick("sparse subtrees");
  // ---- storage offsets, levels, statistics ---------------------------------------------------------
  P.Lptr.assign(ns + 1, 0);
  P.Wptr.assign(ns + 1, 0);
  P.sn_level.assign(ns, 0);
  P.max_front = 0;
  for (int T = 0; T < ns; ++T)
  {
    const i64 k = P.sn_first[T + 1] - P.sn_first[T];
    const i64 r = P.Rptr[T + 1] - P.Rptr[T];
    const i64 h = k + r;
    i64 sz      = P.sn_sparse[T] ? sst_nnz[T] : panel_ld(h) * k; // leading dimension padded to an even number of rows (plan.hpp)
    sz          = (sz + 3) & ~(i64)3;
    P.Lptr[T + 1] = P.Lptr[T] + sz;
    P.Wptr[T + 1] = P.Wptr[T] + h;
    if (P.sn_sparse[T])
    {
      P.sst[(size_t)sst_index[T]].Lptr = P.Lptr[T];
      P.nnzL_stored += sst_nnz[T];
      for (int j = P.sn_first[T]; j < P.sn_first[T + 1]; ++j)
      {
        P.flops_stored += (double)cc[j] * (double)cc[j];
      }
      continue;
    }
    P.max_front   = std::max<i64>(P.max_front, h);
    P.nnzL_stored += k * h - k * (k - 1) / 2;
    for (i64 c = 0; c < k; ++c)
    {
      P.flops_stored += (double)(h - c) * (double)(h - c);
    }
  }
  for (int j = 0; j < m; ++j)
  {
    P.nnzL += cc[j];
    P.flops += (double)cc[j] * (double)cc[j];
  }
  for (int T = 0; T < ns; ++T)
  {
    if (P.sn_parent[T] >= 0)
    {
      P.sn_level[P.sn_parent[T]] = std::max(P.sn_level[P.sn_parent[T]], P.sn_level[T] + 1);
    }
  }
  P.nlevels = 0;
  for (int T = 0; T < ns; ++T)
  {
    P.nlevels = std::max(P.nlevels, P.sn_level[T] + 1);
  }
  P.lvl_ptr.assign(P.nlevels + 1, 0);
  for (int T = 0; T < ns; ++T)
  {
    ++P.lvl_ptr[P.sn_level[T] + 1];
  }
  for (int l = 0; l < P.nlevels; ++l)
  {
    P.lvl_ptr[l + 1] += P.lvl_ptr[l];
  }
  P.lvl_sn.resize(ns);
  {
    std::vector<int> fill(P.lvl_ptr.begin(), P.lvl_ptr.end() - 1);
    for (int T = 0; T < ns; ++T)
    {
      P.lvl_sn[fill[P.sn_level[T]]++] = T;
    }
  }

  tick("offsets + levels");
  // ---- assembly map of S into the panels ------------------------------------------------------------------
  {
    // lower pattern of S by column (new labels), rows ascending incl. diagonal
    std::vector<i64> Sptr(m + 1, 0);
    for (int j = 0; j < m; ++j)
    {
      i64 cnt = 1;
      for (int p = xadj2[j]; p < xadj2[j + 1]; ++p)
      {
        cnt += adj2[p] > j;
      }
      Sptr[j + 1] = Sptr[j] + cnt;
    }
    P.nnzS = Sptr[m];
    std::vector<int> Srow((size_t)P.nnzS);
    P.Sdest.resize((size_t)P.nnzS);
    P.Sdiag.resize((size_t)m);
    std::atomic<bool> outside{false};
    parallel_ranges(m, host_threads(P.nnzS, 200000), [&](int, i64 jlo, i64 jhi) {
    for (int j = (int)jlo; j < (int)jhi; ++j)
    {
      i64 o     = Sptr[j];
      Srow[o++] = j;
      for (int p = xadj2[j]; p < xadj2[j + 1]; ++p)
      {
        if (adj2[p] > j)
        {
          Srow[o++] = adj2[p];
        }
      }
      const int T = P.sn_of_col[j];
      const int f = P.sn_first[T], l = P.sn_first[T + 1] - 1, k = l - f + 1;
      const i64 h      = k + (P.Rptr[T + 1] - P.Rptr[T]);
      const int* rows  = P.Ridx.data() + P.Rptr[T];
      const int nrows  = (int)(P.Rptr[T + 1] - P.Rptr[T]);
      if (P.sn_sparse[T])
      {
        // sparse subtree: the entry goes to its slot in the compact column (rows there are front-local indices)
        const int cp        = sst_col_of[j];
        const SstMeta& M    = P.sst[(size_t)sst_index[T]];
        const int c0        = P.sst_colptr[(size_t)cp], c1 = P.sst_colptr[(size_t)cp + 1];
        const int* crow     = P.sst_rows.data() + M.row_ptr + c0;
        for (i64 q = Sptr[j]; q < Sptr[j + 1]; ++q)
        {
          const int i = Srow[q];
          int loc;
          if (i <= l)
          {
            loc = i - f;
          }
          else
          {
            const int* it = std::lower_bound(rows, rows + nrows, i);
            loc           = (it != rows + nrows && *it == i) ? k + (int)(it - rows) : -1;
          }
          const int* it = std::lower_bound(crow, crow + (c1 - c0), loc);
          if (loc < 0 || it == crow + (c1 - c0) || *it != loc)
          {
            outside = true;
            it      = crow;
          }
          P.Sdest[q] = P.Lptr[T] + c0 + (it - crow);
        }
        P.Sdiag[j] = P.Sdest[Sptr[j]];
        continue;
      }
      for (i64 q = Sptr[j]; q < Sptr[j + 1]; ++q)
      {
        int i = Srow[q];
        i64 rowpos;
        if (i <= l)
        {
          rowpos = i - f;
        }
        else
        {
          const int* it = std::lower_bound(rows, rows + nrows, i);
          if (it == rows + nrows || *it != i)
          {
            outside = true;
            it      = rows;
          }
          rowpos = k + (it - rows);
        }
        P.Sdest[q] = P.Lptr[T] + (i64)(j - f) * panel_ld(h) + rowpos;
      }
      P.Sdiag[j] = P.Sdest[Sptr[j]];
    }
    });
    if (outside)
    {
      return fail(err, B200_ERR_ARG, "internal: S entry outside the supernode structure");
    }
    tick("  assembly: destinations");
    auto entry_id = [&](int row, int col) -> i64 {
      const int* b  = Srow.data() + Sptr[col];
      const int* e  = Srow.data() + Sptr[col + 1];
      const int* it = std::lower_bound(b, e, row);
      return (it != e && *it == row) ? (i64)(it - Srow.data()) : -1;
    };
    P.Sgsrc.assign((size_t)P.nnzS, -1);
    for (const Ent& g : Gent)
    {
      int a = P.pinv[g.r], b = P.pinv[g.c];
      i64 id = entry_id(std::max(a, b), std::min(a, b));
      if (id < 0)
      {
        return fail(err, B200_ERR_ARG, "internal: G entry missing from the S pattern");
      }
      P.Sgsrc[id] = g.src;
    }
    // product terms. The entry of S every term belongs to is found by binary search in the column of S: that is
    // the expensive part (c (c + 1) / 2 terms per eliminated variable with c entries) and independent per variable,
    // so it runs on a few host threads; counting and filling are cheap linear passes.
    std::vector<i64> tptr((size_t)nE + 1, 0);
    for (int e = 0; e < nE; ++e)
    {
      const i64 c = P.Acsc_ptr[e + 1] - P.Acsc_ptr[e];
      tptr[e + 1] = tptr[e] + c * (c + 1) / 2;
    }
    std::vector<i64> term_ids((size_t)tptr[nE]); // entry id of every product term, in (e, s, t) order
    {
      std::atomic<bool> missing{false};
      auto search = [&](int e0, int e1) {
        for (int e = e0; e < e1; ++e)
        {
          i64 o = tptr[e];
          for (int s = P.Acsc_ptr[e]; s < P.Acsc_ptr[e + 1]; ++s)
          {
            const int a = P.pinv[P.Acsc_row[s]];
            for (int t = P.Acsc_ptr[e]; t <= s; ++t)
            {
              const int b  = P.pinv[P.Acsc_row[t]];
              const i64 id = entry_id(std::max(a, b), std::min(a, b));
              if (id < 0)
              {
                missing = true;
              }
              term_ids[(size_t)o++] = id;
            }
          }
        }
      };
      const int nth = host_threads(tptr[nE], 200000);
      if (nth == 1)
      {
        search(0, nE);
      }
      else
      {
        std::vector<std::thread> pool;
        for (int th = 0; th < nth; ++th)
        {
          // equal shares of the terms, not of the variables
          const i64 lo = tptr[nE] * th / nth, hi = tptr[nE] * (th + 1) / nth;
          const int e0 = (int)(std::lower_bound(tptr.begin(), tptr.end(), lo) - tptr.begin());
          const int e1 = th + 1 == nth ? nE : (int)(std::lower_bound(tptr.begin(), tptr.end(), hi) - tptr.begin());
          pool.emplace_back(search, std::min(e0, nE), std::min(e1, nE));
        }
        for (auto& th : pool)
        {
          th.join();
        }
      }
      if (missing)
      {
        return fail(err, B200_ERR_ARG, "internal: product term missing from the S pattern");
      }
    }
    tick("  assembly: term search");
    // Terms are grouped by the entry of S they belong to, in (e, s, t) order inside a group. Each thread owns a
    // contiguous range of entries and walks all terms, keeping those of its range: the order inside a group -- and with
    // it the summation order on the device -- does not depend on the number of threads.
    P.Sterm_ptr.assign((size_t)P.nnzS + 1, 0);
    const int nth_fill = host_threads(tptr[nE], 2000000);
    parallel_ranges(P.nnzS, nth_fill, [&](int, i64 lo, i64 hi) {
      for (i64 id : term_ids)
      {
        if (id >= lo && id < hi)
        {
          ++P.Sterm_ptr[(size_t)id + 1];
        }
      }
    });
    for (i64 q = 0; q < P.nnzS; ++q)
    {
      P.Sterm_ptr[q + 1] += P.Sterm_ptr[q];
    }
    {
      const size_t nt = (size_t)P.Sterm_ptr[P.nnzS];
      P.Sterm_a.resize(nt);
      P.Sterm_b.resize(nt);
      P.Sterm_d.resize(nt);
      std::vector<i64> fill(P.Sterm_ptr.begin(), P.Sterm_ptr.end() - 1);
      parallel_ranges(P.nnzS, nth_fill, [&](int, i64 lo, i64 hi) {
        size_t tq = 0;
        for (int e = 0; e < nE; ++e)
        {
          for (int s = P.Acsc_ptr[e]; s < P.Acsc_ptr[e + 1]; ++s)
          {
            for (int t = P.Acsc_ptr[e]; t <= s; ++t)
            {
              const i64 id = term_ids[tq++];
              if (id < lo || id >= hi)
              {
                continue;
              }
              const i64 o  = fill[(size_t)id]++;
              P.Sterm_a[o] = P.Acsc_src[s];
              P.Sterm_b[o] = P.Acsc_src[t];
              P.Sterm_d[o] = P.dE_src[e];
            }
          }
        }
      });
    }
  }

  tick("assembly map");
  // ---- numeric schedule: stages of panel steps -------------------------------------------------------------
  P.sn_base.assign(ns, 0);
  P.sn_nt.assign(ns, 1);
  int nstages = 0;
  for (int T = 0; T < ns; ++T)
  {
    const int k = P.sn_first[T + 1] - P.sn_first[T];
    P.sn_nt[T]  = P.sn_sparse[T] ? 0 : (k + NB - 1) / NB; // sparse subtrees are factored before the first stage (k_sst_factor)
  }
  for (int T = 0; T < ns; ++T) // children precede parents
  {
    nstages = std::max(nstages, P.sn_base[T] + P.sn_nt[T]);
    int p   = P.sn_parent[T];
    if (p >= 0)
    {
      P.sn_base[p] = std::max(P.sn_base[p], P.sn_base[T] + P.sn_nt[T]);
    }
  }
  // update-matrix workspace: U_T lives from stage base[T] to stage base[parent] (inclusive)
  P.Uoff.assign(ns, -1);
  {
    const int alloc_stages = std::max(nstages, 1); // update blocks of sparse subtrees exist even without a single dense stage
    std::vector<std::vector<int>> alloc_at(alloc_stages + 1), free_at(alloc_stages + 2);
    for (int T = 0; T < ns; ++T)
    {
      i64 r = P.Rptr[T + 1] - P.Rptr[T];
      if (r == 0)
      {
        continue;
      }
      alloc_at[P.sn_base[T]].push_back(T);
      int p = P.sn_parent[T];
      free_at[(p >= 0 ? P.sn_base[p] : alloc_stages - 1) + 1].push_back(T);
    }
    std::map<i64, i64> freelist; // offset -> size
    i64 top = 0;
    auto release = [&](i64 off, i64 sz) {
      auto it = freelist.insert({off, sz}).first;
      auto nx = std::next(it);
      if (nx != freelist.end() && it->first + it->second == nx->first)
      {
        it->second += nx->second;
        freelist.erase(nx);
      }
      if (it != freelist.begin())
      {
        auto pv = std::prev(it);
        if (pv->first + pv->second == it->first)
        {
          pv->second += it->second;
          freelist.erase(it);
          it = pv;
        }
      }
      if (it->first + it->second == top)
      {
        top = it->first;
        freelist.erase(it);
      }
    };
    std::vector<i64> usz(ns, 0);
    for (int s = 0; s < alloc_stages; ++s)
    {
      for (int T : free_at[s])
      {
        release(P.Uoff[T], usz[T]);
      }
      for (int T : alloc_at[s])
      {
        i64 r  = P.Rptr[T + 1] - P.Rptr[T];
        i64 sz = (r * r + 3) & ~(i64)3;
        usz[T] = sz;
        i64 off = -1;
        // first fit among the largest few: bounded scan keeps this O(n log n) in practice
        int scanned = 0;
        for (auto it = freelist.begin(); it != freelist.end() && scanned < 64; ++it, ++scanned)
        {
          if (it->second >= sz)
          {
            off = it->first;
            i64 rem = it->second - sz;
            i64 o2  = it->first + sz;
            freelist.erase(it);
            if (rem > 0)
            {
              freelist.insert({o2, rem});
            }
            break;
          }
        }
        if (off < 0)
        {
          off = top;
          top += sz;
        }
        P.Uoff[T]  = off;
        P.Utotal   = std::max(P.Utotal, top);
      }
    }
  }
  for (SstMeta& M : P.sst)
  {
    M.Uoff = M.r > 0 ? P.Uoff[M.sn] : 0;
  }
  for (size_t e = 0; e < P.sst_ea_src.size(); ++e)
  {
    P.sst_ea_src[e] += P.Uoff[(size_t)sst_ea_child[e]];
  }
  // tasks
  {
    std::vector<std::vector<int>> starts(nstages), active(nstages);
    for (int T = 0; T < ns; ++T)
    {
      if (P.sn_sparse[T])
      {
        continue; // no children to assemble, no panel steps (and possibly no stage at all)
      }
      starts[P.sn_base[T]].push_back(T);
      for (int t = 0; t < P.sn_nt[T]; ++t)
      {
        active[P.sn_base[T] + t].push_back(T);
      }
    }
    P.stages.resize(nstages);
    P.n_scratch_slots = 0;
    for (int s = 0; s < nstages; ++s)
    {
      Stage& st     = P.stages[s];
      st.zero_begin = (int)P.zero_sn.size();
      st.ea_begin   = (int)P.ea_tasks.size();
      for (int T : starts[s])
      {
        if (P.child_ptr[T] == P.child_ptr[T + 1])
        {
          continue;
        }
        if (P.Rptr[T + 1] > P.Rptr[T])
        {
          P.zero_sn.push_back(T);
        }
        for (int q = P.child_ptr[T]; q < P.child_ptr[T + 1]; ++q)
        {
          int c  = P.child_idx[q];
          int rc = (int)(P.Rptr[c + 1] - P.Rptr[c]);
          for (int jb = 0; jb * EA_COLS < rc; ++jb)
          {
            P.ea_tasks.push_back({c, jb});
          }
        }
      }
      st.zero_end  = (int)P.zero_sn.size();
      st.ea_end    = (int)P.ea_tasks.size();
      st.pan_begin  = (int)P.pan_tasks.size();
      st.upd_begin  = (int)P.upd_tasks.size();
      for (int T : active[s])
      {
        const int t  = s - P.sn_base[T];
        const int k  = P.sn_first[T + 1] - P.sn_first[T];
        const int r  = (int)(P.Rptr[T + 1] - P.Rptr[T]);
        const int h  = k + r;
        const int c0 = t * NB;
        const int w  = std::min(NB, k - c0);
        const int below = h - c0 - w;
        for (int rb = 0; rb == 0 || rb * RB < below; ++rb)
        {
          P.pan_tasks.push_back({T, t, rb, 0});
        }
      }
      // Trailing update inside the panel: front columns [c0+w, k), rows >= column. Two-level blocking: columns of the
      // same NBO-wide outer block ("near") are updated after every panel step with K = w; the columns beyond it ("far")
      // once, after the last step of the outer block, with K = the whole outer block -- DMMA tiles with a decent
      // arithmetic intensity. Look-ahead: the tiles of the NEXT panel's NB columns come first (pass 0); the next
      // panel step waits for those only, the rest (pass 1, incl. the Schur complement) runs next to it.
      for (int pass = 0; pass < 2; ++pass)
      {
        if (pass == 1)
        {
          st.upd_mid = (int)P.upd_tasks.size();
        }
        for (int T : active[s])
        {
          const int t  = s - P.sn_base[T];
          const int k  = P.sn_first[T + 1] - P.sn_first[T];
          const int r  = (int)(P.Rptr[T + 1] - P.Rptr[T]);
          const int h  = k + r;
          const int c0 = t * NB;
          const int w  = std::min(NB, k - c0);
          const int ob_begin = (c0 / NBO) * NBO;
          const int ob_end   = std::min(k, ob_begin + NBO);
          const int np0 = c0 + w, np1 = std::min(k, np0 + NB); // columns of the next panel step
          auto tiles = [&](int jb, int je, int kb, int ke) {  // columns [jb, je), all rows below
            for (int j0 = jb; j0 < je; j0 += TILE)
            {
              for (int i0 = j0; i0 < h; i0 += TILE)
              {
                P.upd_tasks.push_back({T, je, UPD_INPANEL, i0, j0, kb, ke});
              }
            }
          };
          if (pass == 0)
          {
            if (np0 < ob_end)
            {
              tiles(np0, np1, c0, c0 + w); // near
            }
            else if (np0 < k)
            {
              tiles(np0, np1, ob_begin, ob_end); // the next panel opens a new outer block: its share of the far update
            }
          }
          else
          {
            tiles(std::min(ob_end, np0 + NB), ob_end, c0, c0 + w);
            if (c0 + w == ob_end)
            {
              tiles(std::min(k, ob_end + NB), k, ob_begin, ob_end);
            }
            if (t == P.sn_nt[T] - 1 && r > 0)
            {
              // Schur tiles start at EVEN front rows (a TMA box must start on a 16-byte boundary of the panel column):
              // for an odd k the tile grid is shifted up by one row, tile (i0, j0) covers the rows i0 - 1 .. i0 + 62 of
              // the update matrix (row -1 does not exist and is masked), so the grid spans r + 1 rows
              const int span = r + (k & 1);
              for (int j0 = 0; j0 < span; j0 += TILE)
              {
                for (int i0 = j0; i0 < span; i0 += TILE)
                {
                  P.upd_tasks.push_back({T, r, UPD_SCHUR, i0, j0, 0, k});
                }
              }
            }
          }
        }
      }
      st.pan_end        = (int)P.pan_tasks.size();
      st.upd_end        = (int)P.upd_tasks.size();
    }
  }

  tick("stages + U alloc + tasks");
  // ---- selective inversion schedule ------------------------------------------------------------------------------
  // After the factorization every panel [L11; L21] is turned into Minv = [L11^-1; -L21 L11^-1], so that both
  // sweeps of a solve are plain matrix-vector products per supernode (Raghavan's selective inversion). L11^-1
  // is built by recursive halving over NB-blocks: inv([A 0; B C]) = [Ainv 0; -Cinv B Ainv, Cinv]; the NB x NB
  // diagonal blocks are inverted by k_panel itself. Nodes of equal height (over all supernodes) share a launch.
  {
    P.Tptr.assign(ns + 1, 0);
    struct Node
    {
      int sn, a, b, c, ht;
    };
    std::vector<Node> nodes;
    int hmax = 0;
    // iterative-recursive split; returns the height of block range [a, c)
    struct Rec
    {
      std::vector<Node>& nodes;
      int sn;
      int run(int a, int c)
      {
        if (c - a <= 1)
        {
          return 0;
        }
        int half = 1;
        while (half * 2 < c - a)
        {
          half *= 2;
        }
        const int b  = a + half;
        const int hl = run(a, b), hr = run(b, c);
        const int ht = 1 + std::max(hl, hr);
        nodes.push_back({sn, a, b, c, ht});
        return ht;
      }
    };
    for (int T = 0; T < ns; ++T)
    {
      const i64 k  = P.sn_first[T + 1] - P.sn_first[T];
      const int nb = P.sn_sparse[T] ? 0 : (int)((k + NB - 1) / NB); // sparse subtrees keep their factor as it is (substitution)
      P.Tptr[T + 1] = P.Tptr[T] + (nb > 1 ? ((k * k + 3) & ~(i64)3) : 0);
      if (nb > 1)
      {
        Rec rec{nodes, T};
        hmax = std::max(hmax, rec.run(0, nb));
      }
    }
    P.inv_phase_ptr.assign(1, 0);
    for (int ht = 1; ht <= hmax; ++ht)
    {
      for (int kind = INV_T1; kind <= INV_T2; ++kind)
      {
        for (const Node& nd : nodes)
        {
          if (nd.ht != ht)
          {
            continue;
          }
          const int k    = P.sn_first[nd.sn + 1] - P.sn_first[nd.sn];
          const int ca   = nd.a * NB, cb = nd.b * NB;       // A columns [ca, cb)
          const int rend = std::min(nd.c * NB, k);            // C rows [cb, rend)
          for (int j0 = ca; j0 < cb; j0 += TILE)
          {
            for (int i0 = cb; i0 < rend; i0 += TILE)
            {
              if (kind == INV_T1)
              {
                P.inv_tasks.push_back({nd.sn, INV_T1, i0, j0, j0, cb}); // Ainv[q, j] = 0 for q < j
              }
              else
              {
                P.inv_tasks.push_back({nd.sn, INV_T2, i0, j0, cb, std::min(i0 + TILE, rend)}); // Cinv[i, q] = 0 for q > i
              }
            }
          }
        }
        P.inv_phase_ptr.push_back((int)P.inv_tasks.size());
      }
    }
    for (int T = 0; T < ns; ++T)
    {
      const int k = P.sn_sparse[T] ? 0 : P.sn_first[T + 1] - P.sn_first[T];
      const int r = (int)(P.Rptr[T + 1] - P.Rptr[T]);
      for (int j0 = 0; j0 < k; j0 += TILE)
      {
        for (int i0 = 0; i0 < r; i0 += TILE)
        {
          P.inv_tasks.push_back({T, INV_Z, i0, j0, j0, k});
        }
      }
    }
    P.inv_phase_ptr.push_back((int)P.inv_tasks.size());
    // row-major copy of every inverse panel for the forward sweep
    for (int T = 0; T < ns; ++T)
    {
      const int k = P.sn_sparse[T] ? 0 : P.sn_first[T + 1] - P.sn_first[T];
      const int h = k + (int)(P.Rptr[T + 1] - P.Rptr[T]);
      for (int j0 = 0; j0 < k; j0 += 32)
      {
        // one tile band above the diagonal is copied too (zeros): a warp's row group may straddle a tile edge
        for (int i0 = std::max(0, j0 - 32); i0 < h; i0 += 32)
        {
          P.tr_tasks.push_back({T, i0, j0});
        }
      }
    }
  }

  // useful flops of the tile classes (roofline numerators of the DMMA kernels): 2 na nb K per tile, half of it on the
  // diagonal tiles of a symmetric update (only the lower triangle is kept)
  for (const Task5& t : P.upd_tasks)
  {
    const i64 k = P.sn_first[t.sn + 1] - P.sn_first[t.sn], r = P.Rptr[t.sn + 1] - P.Rptr[t.sn];
    const bool schur = t.kind == UPD_SCHUR;
    const i64 sh     = schur ? (k & 1) : 0; // shifted tile grid of the Schur tiles (see the task generation)
    const i64 u0 = t.i0 - sh, v0 = t.j0 - sh;
    const double na = (double)(std::min<i64>(u0 + TILE, schur ? r : k + r) - std::max<i64>(u0, 0));
    const double nb = (double)(std::min<i64>(v0 + TILE, schur ? r : (i64)t.jend) - std::max<i64>(v0, 0));
    P.flops_update += (t.i0 == t.j0 ? 1.0 : 2.0) * na * nb * (double)(t.ke - t.kb);
  }
  for (const InvTask& t : P.inv_tasks)
  {
    const i64 k = P.sn_first[t.sn + 1] - P.sn_first[t.sn], r = P.Rptr[t.sn + 1] - P.Rptr[t.sn];
    double na, nb;
    if (t.kind == INV_T1)
    {
      na = (double)std::min<i64>(TILE, k - t.i0), nb = (double)std::min<i64>(TILE, t.ke - t.j0);
    }
    else if (t.kind == INV_T2)
    {
      na = (double)std::min<i64>(TILE, t.ke - t.i0), nb = (double)std::min<i64>(TILE, t.kb - t.j0);
    }
    else
    {
      na = (double)std::min<i64>(TILE, r - t.i0), nb = (double)std::min<i64>(TILE, k - t.j0);
    }
    P.flops_inv += 2.0 * na * nb * (double)(t.ke - t.kb);
  }
  tick("inversion + transpose tasks");
  // ---- dataflow sweeps: warp tasks in ticket order, dependency counters per supernode -------------------------
  {
    constexpr int LANES = 32;
    constexpr int FLOW_WARPS = 148 * FLOW_CTAS_PER_SM * (FLOW_THREADS / 32);
    if ((i64)P.Ridx.size() > 0x7ffffff0)
    {
      return fail(err, B200_ERR_UNSUPPORTED, "row-index array exceeds 2^31 entries");
    }
    // Depth of a task: 16 panel entries per lane (one batch of loads, all in flight before the warp waits) where a
    // level has fewer tasks than the machine has warps -- the narrow levels are latency-bound --, 32 (two batches)
    // where it has more: there the fixed cost of a task (record, publication, ticket) is what limits the bandwidth;
    // FLOW_DEEP where even that many tasks are left (large 3D fronts).
    std::vector<int> depth_of_level(P.nlevels, 16);
    i64 deep_min = FLOW_WARPS / 2; // tasks a level must still have at depth FLOW_DEEP (B200_FLOW_DEEP_TASKS: tests, tuning)
    if (const char* dm = std::getenv("B200_FLOW_DEEP_TASKS"))
    {
      deep_min = std::max<i64>(1, std::atoll(dm));
    }
    for (int l = 0; l < P.nlevels; ++l)
    {
      i64 entries = 0;
      for (int q = P.lvl_ptr[l]; q < P.lvl_ptr[l + 1]; ++q)
      {
        const int T = P.lvl_sn[q];
        if (!P.sn_sparse[T])
        {
          entries += (P.Wptr[T + 1] - P.Wptr[T]) * (i64)(P.sn_first[T + 1] - P.sn_first[T]);
        }
      }
      // (measured on B200: a lower threshold for the second batch is slower on every configuration.) Levels that
      // still give every warp a task at a depth of FLOW_DEEP are bandwidth-bound: there the panel is streamed with a
      // rolling window of loads and the fixed cost of a task is amortised over four times as much data.
      depth_of_level[l] = entries / (LANES * FLOW_DEEP) >= deep_min ? FLOW_DEEP : entries / (LANES * 16) >= (i64)FLOW_WARPS ? 32 : 16;
    }
    // forward: lanes = rows, depth = columns (row r of the triangular top block needs columns <= r)
    // backward: lanes = columns (blocks aligned to 32 like the tiles of the row-major copy), depth = rows (column j
    //           needs rows >= j)
    // A deep chunk must really be deep: what is left of a depth range below FLOW_DEEP_MIN goes out in chunks of 32 as in
    // the other wide levels (small supernodes in a level that qualifies as a whole; the tail of a large one).
    // And only supernodes with at least FLOW_DEEP columns get them: the small supernodes of a level that qualifies as a
    // whole (bottom levels of large 2D problems) keep the task shapes they were measured with.
    constexpr int FLOW_DEEP_MIN = 64;
    auto depth_of_sn = [&](int T) {
      const int d = depth_of_level[P.sn_level[T]];
      return (d > 32 && P.sn_first[T + 1] - P.sn_first[T] < FLOW_DEEP) ? 32 : d;
    };
    auto chunk = [&](int remaining, int DEPTH) { return (DEPTH > 32 && remaining < FLOW_DEEP_MIN) ? 32 : DEPTH; };
    auto fwd_blocks = [&](int k, int h, int DEPTH, auto&& emit) {
      for (int i0 = 0; i0 < h; i0 += LANES)
      {
        const int i1 = std::min(h, i0 + LANES);
        const int je = std::min(k, i1);
        for (int j0 = 0, d; j0 < je; j0 += d)
        {
          d = chunk(je - j0, DEPTH);
          emit(i0, i1, j0, std::min(je, j0 + d));
        }
      }
    };
    auto bwd_blocks = [&](int k, int h, int DEPTH, auto&& emit) {
      for (int j0 = 0; j0 < k; j0 += LANES)
      {
        for (int i0 = j0, d; i0 < h; i0 += d) // j0 is a multiple of 32, so the depth blocks are aligned to 16
        {
          d = chunk(h - i0, DEPTH);
          emit(i0, std::min(h, i0 + d), j0, std::min(k, j0 + LANES));
        }
      }
    };
    std::vector<int> nf(ns, 0), nbk(ns, 0);
    for (int T = 0; T < ns; ++T)
    {
      const int k = P.sn_first[T + 1] - P.sn_first[T];
      const int h = (int)(P.Wptr[T + 1] - P.Wptr[T]);
      if (P.sn_sparse[T])
      {
        nf[T] = 1; // a sparse subtree is swept by one CTA before the dataflow kernel starts: one signal to its parent
        continue;
      }
      fwd_blocks(k, h, depth_of_sn(T), [&](int, int, int, int) { ++nf[T]; });
      bwd_blocks(k, h, depth_of_sn(T), [&](int, int, int, int) { ++nbk[T]; });
    }
    for (int l = 0; l < P.nlevels; ++l)
    {
      std::vector<int> order(P.lvl_sn.begin() + P.lvl_ptr[l], P.lvl_sn.begin() + P.lvl_ptr[l + 1]);
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return (P.Wptr[x + 1] - P.Wptr[x]) > (P.Wptr[y + 1] - P.Wptr[y]); });
      for (int T : order)
      {
        if (P.sn_sparse[T])
        {
          continue;
        }
        const int k = P.sn_first[T + 1] - P.sn_first[T];
        const int h = (int)(P.Wptr[T + 1] - P.Wptr[T]);
        int need    = 0;
        for (int q = P.child_ptr[T]; q < P.child_ptr[T + 1]; ++q)
        {
          need += nf[P.child_idx[q]];
        }
        const int wait_idx = need > 0 ? T : -1;
        fwd_blocks(k, h, depth_of_sn(T), [&](int i0, int i1, int j0, int j1) {
          P.ffl_tasks.push_back({P.Lptr[T], (int)P.Rptr[T], P.sn_first[T], k, (int)panel_ld(h), i0, i1, j0, j1, wait_idx, need, P.sn_parent[T], l, depth_of_level[l] == 16, 0});
        });
      }
    }
    for (int l = P.nlevels - 1; l >= 0; --l)
    {
      std::vector<int> order(P.lvl_sn.begin() + P.lvl_ptr[l], P.lvl_sn.begin() + P.lvl_ptr[l + 1]);
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return (P.Wptr[x + 1] - P.Wptr[x]) > (P.Wptr[y + 1] - P.Wptr[y]); });
      for (int T : order)
      {
        if (P.sn_sparse[T])
        {
          continue;
        }
        const int k   = P.sn_first[T + 1] - P.sn_first[T];
        const int h   = (int)(P.Wptr[T + 1] - P.Wptr[T]);
        const int par = P.sn_parent[T];
        const int signal_idx = P.child_ptr[T + 1] > P.child_ptr[T] ? T : -1;
        bwd_blocks(k, h, depth_of_sn(T), [&](int i0, int i1, int j0, int j1) {
          const bool tail = i1 > k && par >= 0; // touches x of the ancestors
          P.bfl_tasks.push_back({P.Lptr[T], (int)P.Rptr[T], P.sn_first[T], k, (int)panel_ld(h), i0, i1, j0, j1, tail ? par : -1, tail ? nbk[par] : 0, signal_idx, l, depth_of_level[l] == 16, 0});
        });
      }
    }
  }

  // ---- operators of the E-block elimination with their index chains resolved once (k_pre / k_post: one dependent
  //      load less per entry) -------------------------------------------------------------------------------------
  P.Acsr_k.resize(P.Acsr_col.size());
  P.Acsr_dsrc.resize(P.Acsr_col.size());
  for (size_t q = 0; q < P.Acsr_col.size(); ++q)
  {
    P.Acsr_k[q]    = P.k_of_e[P.Acsr_col[q]];
    P.Acsr_dsrc[q] = P.dE_src[P.Acsr_col[q]];
  }
  P.Acsc_p.resize(P.Acsc_row.size());
  for (size_t q = 0; q < P.Acsc_row.size(); ++q)
  {
    P.Acsc_p[q] = P.pinv[P.Acsc_row[q]];
  }
  tick("sweep tasks");
  // ---- hash of the full permutation -----------------------------------------------------------------------------
  {
    std::vector<int> fp(n);
    full_structure(P, fp.data(), nullptr, nullptr, nullptr, nullptr);
    P.perm_hash = fnv1a(14695981039346656037ull, fp.data(), sizeof(int) * (size_t)n);
  }
  P.ms_symbolic = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return B200_OK;
}

void
full_structure(const Plan& P, int* perm, int* parent, int* colcount, int* n_super_total, int* super_first)
{
  const int nE = P.nE, m = P.m;
  if (perm)
  {
    for (int e = 0; e < nE; ++e)
    {
      perm[e] = P.k_of_e[e];
    }
    for (int j = 0; j < m; ++j)
    {
      perm[nE + j] = P.k_of_r[P.perm[j]];
    }
  }
  if (parent || colcount)
  {
    for (int e = 0; e < nE; ++e)
    {
      int best = -1;
      for (int q = P.Acsc_ptr[e]; q < P.Acsc_ptr[e + 1]; ++q)
      {
        int pos = nE + P.pinv[P.Acsc_row[q]];
        if (best == -1 || pos < best)
        {
          best = pos;
        }
      }
      if (parent)
      {
        parent[e] = best;
      }
      if (colcount)
      {
        colcount[e] = 1 + (P.Acsc_ptr[e + 1] - P.Acsc_ptr[e]);
      }
    }
    for (int j = 0; j < m; ++j)
    {
      if (parent)
      {
        parent[nE + j] = P.parent[j] == -1 ? -1 : nE + P.parent[j];
      }
      if (colcount)
      {
        colcount[nE + j] = P.colcount[j];
      }
    }
  }
  if (n_super_total)
  {
    *n_super_total = nE + P.nsuper;
  }
  if (super_first)
  {
    for (int e = 0; e < nE; ++e)
    {
      super_first[e] = e;
    }
    for (int T = 0; T <= P.nsuper; ++T)
    {
      super_first[nE + T] = nE + P.sn_first[T];
    }
  }
}

} // namespace b200
