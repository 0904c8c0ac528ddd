// device.cuh -- CUDA helpers shared by the kernels and the C-ABI layer.
#pragma once
#include "common.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace b200
{

struct CudaError : std::runtime_error
{
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define B200_CUDA(expr)                                                                                                  \
  do                                                                                                                     \
  {                                                                                                                      \
    cudaError_t e__ = (expr);                                                                                            \
    if (e__ != cudaSuccess)                                                                                              \
    {                                                                                                                    \
      throw ::b200::CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +              \
                              std::to_string(__LINE__) + ")");                                                           \
    }                                                                                                                    \
  } while (0)

extern std::atomic<int64_t> g_launches;

// Counts kernel launches: directly when launched eagerly, into `captured` while a graph is being
// recorded (each replay then adds the recorded count).
struct LaunchCounter
{
  int64_t* captured = nullptr;
  // optional tracer (profiling entry points): called after each launch with a kernel-class tag
  void (*trace)(void* ctx, const char* tag) = nullptr;
  void* trace_ctx                           = nullptr;
  inline void tick(const char* tag = nullptr)
  {
    if (trace && tag)
    {
      trace(trace_ctx, tag);
    }
    if (captured)
    {
      ++*captured;
    }
    else
    {
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
  }
};

template <typename T>
struct DevBuf
{
  T* p       = nullptr;
  size_t cap = 0;
  DevBuf() {}
  DevBuf(const DevBuf&)            = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release()
  {
    if (p)
    {
      cudaFree(p);
      p   = nullptr;
      cap = 0;
    }
  }
  // grows (never shrinks); contents are NOT preserved
  void reserve(size_t n)
  {
    if (n > cap)
    {
      release();
      size_t want = n + std::min<size_t>(n / 8, (size_t)1 << 22) + 16; // slack for slowly growing patterns, bounded for the big panel buffers
      B200_CUDA(cudaMalloc((void**)&p, want * sizeof(T)));
      cap = want;
    }
  }
  void upload(const std::vector<T>& v, cudaStream_t s)
  {
    reserve(v.size());
    if (!v.empty())
    {
      B200_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    }
  }
};

template <typename T>
struct PinnedBuf
{
  T* p       = nullptr;
  size_t cap = 0;
  PinnedBuf() {}
  PinnedBuf(const PinnedBuf&)            = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf()
  {
    if (p)
    {
      cudaFreeHost(p);
    }
  }
  void reserve(size_t n)
  {
    if (n > cap)
    {
      if (p)
      {
        cudaFreeHost(p);
        p = nullptr;
      }
      size_t want = n + n / 8 + 16;
      B200_CUDA(cudaMallocHost((void**)&p, want * sizeof(T)));
      cap = want;
    }
  }
};

// True if `p` is page-locked host memory (cudaMallocHost / cudaHostRegister): the copy engines can read or write it
// directly, so the library skips its own pinned staging copy.
inline bool
is_pinned_host(const void* p)
{
  cudaPointerAttributes attr;
  if (!p || cudaPointerGetAttributes(&attr, p) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

// ... over its whole extent: a registration that another owner made for an array that has since moved may cover only
// the head of `p`
inline bool
is_pinned_host(const void* p, size_t bytes)
{
  return is_pinned_host(p) && (bytes == 0 || is_pinned_host(static_cast<const char*>(p) + bytes - 1));
}

// Per-supernode geometry on the device.
struct SnMeta
{
  long long Lptr;
  long long Uoff;
  long long Rptr;
  long long Tptr;
  int first;
  int k;
  int r;
  int parent;
  int child_begin;
  int child_end;
  int ld;   // leading dimension of the column-major panels L / Mt (panel_ld(k + r))
  int tmap; // index of the panel's TMA tensor map, -1: none (small fronts use the cp.async path)
  int pad2, pad3, pad4, pad5;
};

// A TMA tensor map (CUtensorMap: 128 opaque bytes, 64-byte aligned) of one supernode's panel, encoded on the host
// (fact.cu) and read by the copy engine through a generic address in global memory.
struct alignas(64) PanelTensorMap
{
  unsigned long long opaque[16];
};

// Device copy of a Plan (per handle).
struct DevPlan
{
  std::shared_ptr<const Plan> plan;
  DevBuf<SnMeta> sn;
  DevBuf<PanelTensorMap> tmaps; // one per supernode with a front of at least TMA_MIN_FRONT rows (SnMeta::tmap)
  DevBuf<int> Ridx, rel, child_idx;
  // assembly
  DevBuf<long long> Sdest, Sterm_ptr, Sdiag;
  DevBuf<int> Sgsrc, Sterm_a, Sterm_b, Sterm_d;
  // tasks
  DevBuf<int> zero_sn;
  DevBuf<EaTask> ea_tasks;
  DevBuf<PanelTask> pan_tasks;
  DevBuf<Task5> upd_tasks;
  DevBuf<int> lvl_sn;
  DevBuf<InvTask> inv_tasks;
  DevBuf<TrTask> tr_tasks;
  DevBuf<SweepTask> ffl_tasks, bfl_tasks; // dataflow sweeps
  // E-part / residual operators
  DevBuf<int> k_of_e, k_of_r, pinv, perm, dE_src;
  DevBuf<int> Acsc_ptr, Acsc_row, Acsc_src, Acsr_ptr, Acsr_col, Acsr_src, Gsym_ptr, Gsym_col, Gsym_src;
  DevBuf<int> Acsr_k, Acsr_dsrc, Acsc_p;
  DevBuf<SstMeta> sst; // sparse subtrees (sst.cu)
  DevBuf<unsigned short> sst_blob; // their index structure, 16-bit (a subtree has < 2^16 entries and rows)
  DevBuf<long long> sst_ea_src;    // assembly of child subtrees into their parents
  DevBuf<int> sst_ea_dst;
  DevBuf<int> Ksrc; // set_kkt plans: source of every value of tril(K) in the Jacobian's value array (-1: the constant 1)
};

} // namespace b200
