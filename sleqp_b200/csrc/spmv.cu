// spmv.cu -- CSC sparse matrix-vector products of the EQP loop on the device.
//
//   b200_mat_mult_vec        y = A x     <-> sleqp_mat_mult_vec        (sparse/mat.c:282-310)
//   b200_mat_mult_vec_trans  y = A^T v   <-> sleqp_mat_mult_vec_trans  (sparse/mat.c:312-363)
//
// The reference computes y = A x by scattering column by column into a zeroed dense vector; on
// the device the same sums are formed as gathers over a CSR mirror of the matrix (built once per
// pattern), so no atomics are needed and, for a fixed row, the products are still added in
// ascending column order. y = A^T v is a gather over the CSC columns themselves (the reference
// merges two sorted index lists per column, mat.c:333-353; entries of v that are not stored
// contribute nothing there and exact zeros here).
// Both are bandwidth-bound: one pass over (ptr, idx, val) plus the dense vectors; a sub-warp of
// G lanes works on one row/column, G chosen from the average segment length.
#include "numeric.cuh"

#include <cstring>

using namespace b200;

namespace b200
{

template <int G>
__global__ void __launch_bounds__(256)
k_spmv_gather(int nseg,
              const int* __restrict__ ptr,
              const int* __restrict__ idx,
              const double* __restrict__ val,
              const double* __restrict__ x,
              double* __restrict__ y,
              const int* __restrict__ skip) // optional device flag: non-zero = leave y alone (device-side control flow of the CG)
{
  if (skip && *skip)
  {
    return;
  }
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long seg  = gtid / G;
  const int gl         = threadIdx.x % G;
  double acc           = 0.0;
  if (seg < nseg)
  {
    const int b = ptr[seg], e = ptr[seg + 1];
    for (int q = b + gl; q < e; q += G)
    {
      acc += val[q] * __ldg(x + idx[q]);
    }
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1)
  {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (seg < nseg && gl == 0)
  {
    y[seg] = acc;
  }
}

__global__ void
k_gather_perm(int n, const int* __restrict__ src, const double* __restrict__ in, double* __restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
  {
    out[i] = in[src[i]];
  }
}

__global__ void
k_scatter_sparse(int nnz, const int* __restrict__ idx, int first, const double* __restrict__ val, double* __restrict__ out)
{
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q < nnz)
  {
    out[idx ? idx[q] : first + q] = val[q];
  }
}

static void
launch_spmv(int nseg, long long nnz, const int* ptr, const int* idx, const double* val, const double* x, double* y, cudaStream_t s, const int* skip = nullptr)
{
  if (nseg <= 0)
  {
    return;
  }
  const double avg = (double)nnz / (double)nseg;
  const int T      = 256;
#define B200_SPMV(G)                                                                                                    \
  k_spmv_gather<G><<<(unsigned)(((long long)nseg * G + T - 1) / T), T, 0, s>>>(nseg, ptr, idx, val, x, y, skip)
  if (avg <= 4.0)
  {
    B200_SPMV(1);
  }
  else if (avg <= 8.0)
  {
    B200_SPMV(2);
  }
  else if (avg <= 16.0)
  {
    B200_SPMV(4);
  }
  else if (avg <= 32.0)
  {
    B200_SPMV(8);
  }
  else if (avg <= 64.0)
  {
    B200_SPMV(16);
  }
  else
  {
    B200_SPMV(32);
  }
#undef B200_SPMV
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaGetLastError());
}

} // namespace b200

struct b200_mat
{
  int device          = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int num_rows = 0, num_cols = 0, nnz = 0;
  bool have = false;
  std::vector<int> h_cols, h_rows; // cached pattern
  DevBuf<int> cols, rows, csr_ptr, csr_col, csr_src;
  DevBuf<double> data, csr_val, x, y, sp_val, cp_val;
  DevBuf<int> sp_idx, cp_idx, cp_cnt;
  PinnedBuf<int> h_cnt;
  PinnedBuf<double> h_val, h_out;
  PinnedBuf<int> h_idx;
};

namespace
{

template <typename F>
int
guarded(F&& f)
{
  try
  {
    return f();
  }
  catch (const CudaError& e)
  {
    return set_error(B200_ERR_CUDA, e.what());
  }
  catch (const std::exception& e)
  {
    return set_error(B200_ERR_CUDA, e.what());
  }
}

// dense device vector from a sparse host vector (zero-fill + scatter)
void
stage_sparse(b200_mat* M, int dim, int nnz, const int* idx, const double* val, double* d_dense)
{
  B200_CUDA(cudaStreamSynchronize(M->stream));
  B200_CUDA(cudaMemsetAsync(d_dense, 0, sizeof(double) * (size_t)dim, M->stream));
  if (nnz <= 0)
  {
    return;
  }
  M->sp_val.reserve((size_t)nnz);
  if (is_pinned_host(val, sizeof(double) * (size_t)nnz)) // borrowed buffer, but every caller synchronises the stream before returning
  {
    B200_CUDA(cudaMemcpyAsync(M->sp_val.p, val, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, M->stream));
  }
  else
  {
    M->h_val.reserve((size_t)nnz);
    std::memcpy(M->h_val.p, val, sizeof(double) * (size_t)nnz);
    B200_CUDA(cudaMemcpyAsync(M->sp_val.p, M->h_val.p, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, M->stream));
  }
  const bool contiguous = (idx[nnz - 1] - idx[0]) == nnz - 1;
  const int* d_idx      = nullptr;
  if (!contiguous)
  {
    M->h_idx.reserve((size_t)nnz);
    M->sp_idx.reserve((size_t)nnz);
    std::memcpy(M->h_idx.p, idx, sizeof(int) * (size_t)nnz);
    B200_CUDA(cudaMemcpyAsync(M->sp_idx.p, M->h_idx.p, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, M->stream));
    d_idx = M->sp_idx.p;
  }
  k_scatter_sparse<<<(unsigned)((nnz + 255) / 256), 256, 0, M->stream>>>(nnz, d_idx, idx[0], M->sp_val.p, d_dense);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  B200_CUDA(cudaGetLastError());
}

int
check_sparse(int nnz, const int* idx, const double* val, int dim)
{
  if (nnz < 0 || nnz > dim || (nnz > 0 && (!idx || !val)))
  {
    return set_error(B200_ERR_ARG, "malformed sparse vector");
  }
  if (nnz > 0 && (idx[0] < 0 || idx[nnz - 1] >= dim))
  {
    return set_error(B200_ERR_ARG, "sparse vector index out of range");
  }
  return B200_OK;
}

} // namespace

extern "C" {

int
b200_mat_create(b200_mat** handle, int device)
{
  if (!handle)
  {
    return set_error(B200_ERR_ARG, "null handle pointer");
  }
  *handle = nullptr;
  return guarded([&]() {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    {
      cudaGetLastError();
      return set_error(B200_ERR_CUDA, "no CUDA device available: the B200 backend has no CPU fallback");
    }
    int dev = device;
    if (dev < 0)
    {
      dev = 0;
      for (const char* name : {"B200_DEVICE", "LOCAL_RANK"})
      {
        const char* v = std::getenv(name);
        if (v && *v)
        {
          dev = std::atoi(v);
          break;
        }
      }
    }
    if (dev >= count)
    {
      return set_error(B200_ERR_CUDA, "requested device not visible");
    }
    B200_CUDA(cudaSetDevice(dev));
    std::unique_ptr<b200_mat> M(new b200_mat());
    M->device = dev;
    B200_CUDA(cudaStreamCreateWithFlags(&M->own_stream, cudaStreamNonBlocking));
    M->stream = M->own_stream;
    *handle = M.release();
    return (int)B200_OK;
  });
}

int
b200_mat_set(b200_mat* M, int num_rows, int num_cols, int nnz, const int* cols, const int* rows, const double* data)
{
  if (!M)
  {
    return set_error(B200_ERR_ARG, "null handle");
  }
  if (num_rows < 0 || num_cols < 0 || nnz < 0 || !cols || (nnz > 0 && (!rows || !data)) || cols[0] != 0 || cols[num_cols] != nnz)
  {
    return set_error(B200_ERR_ARG, "malformed CSC matrix");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    B200_CUDA(cudaStreamSynchronize(M->stream));
    const bool same = M->have && M->num_rows == num_rows && M->num_cols == num_cols && M->nnz == nnz &&
                      std::memcmp(M->h_cols.data(), cols, sizeof(int) * (size_t)(num_cols + 1)) == 0 &&
                      (nnz == 0 || std::memcmp(M->h_rows.data(), rows, sizeof(int) * (size_t)nnz) == 0);
    if (!same)
    {
      for (int j = 0; j < num_cols; ++j)
      {
        if (cols[j + 1] < cols[j])
        {
          return set_error(B200_ERR_ARG, "column pointers not monotone");
        }
      }
      for (int q = 0; q < nnz; ++q)
      {
        if (rows[q] < 0 || rows[q] >= num_rows)
        {
          return set_error(B200_ERR_ARG, "row index out of range");
        }
      }
      M->h_cols.assign(cols, cols + num_cols + 1);
      M->h_rows.assign(rows, rows + nnz);
      // CSR mirror by counting sort: columns ascending inside each row
      std::vector<int> ptr((size_t)num_rows + 1, 0), col((size_t)nnz), src((size_t)nnz);
      for (int q = 0; q < nnz; ++q)
      {
        ++ptr[rows[q] + 1];
      }
      for (int i = 0; i < num_rows; ++i)
      {
        ptr[i + 1] += ptr[i];
      }
      std::vector<int> fill(ptr.begin(), ptr.end() - 1);
      for (int j = 0; j < num_cols; ++j)
      {
        for (int q = cols[j]; q < cols[j + 1]; ++q)
        {
          int o  = fill[rows[q]]++;
          col[o] = j;
          src[o] = q;
        }
      }
      M->cols.upload(M->h_cols, M->stream);
      M->rows.upload(M->h_rows, M->stream);
      M->csr_ptr.upload(ptr, M->stream);
      M->csr_col.upload(col, M->stream);
      M->csr_src.upload(src, M->stream);
      B200_CUDA(cudaStreamSynchronize(M->stream)); // ptr/col/src are locals
      M->num_rows = num_rows;
      M->num_cols = num_cols;
      M->nnz      = nnz;
      M->data.reserve((size_t)nnz + 8);
      M->csr_val.reserve((size_t)nnz + 8);
      M->x.reserve((size_t)std::max(num_rows, num_cols) + 8);
      M->y.reserve((size_t)std::max(num_rows, num_cols) + 8);
      M->h_out.reserve((size_t)std::max(num_rows, num_cols) + 8);
      M->have = true;
    }
    if (nnz > 0)
    {
      B200_CUDA(cudaMemcpyAsync(M->data.p, data, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, M->stream));
      k_gather_perm<<<(unsigned)((nnz + 255) / 256), 256, 0, M->stream>>>(nnz, M->csr_src.p, M->data.p, M->csr_val.p);
      g_launches.fetch_add(1, std::memory_order_relaxed);
      B200_CUDA(cudaGetLastError());
    }
    B200_CUDA(cudaStreamSynchronize(M->stream)); // `data` is borrowed
    return (int)B200_OK;
  });
}

int
b200_mat_mult_vec_device(b200_mat* M, const double* d_x, double* d_y)
{
  if (!M || !M->have || !d_x || !d_y)
  {
    return set_error(B200_ERR_STATE, "matrix not set or null vector");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    launch_spmv(M->num_rows, M->nnz, M->csr_ptr.p, M->csr_col.p, M->csr_val.p, d_x, d_y, M->stream);
    return (int)B200_OK;
  });
}

int
b200_mat_mult_vec_device_if(b200_mat* M, const double* d_x, double* d_y, const int* d_skip)
{
  if (!M || !M->have || !d_x || !d_y)
  {
    return set_error(B200_ERR_STATE, "matrix not set or null vector");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    launch_spmv(M->num_rows, M->nnz, M->csr_ptr.p, M->csr_col.p, M->csr_val.p, d_x, d_y, M->stream, d_skip);
    return (int)B200_OK;
  });
}

int
b200_mat_mult_vec_trans_device(b200_mat* M, const double* d_v, double* d_y)
{
  if (!M || !M->have || !d_v || !d_y)
  {
    return set_error(B200_ERR_STATE, "matrix not set or null vector");
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    launch_spmv(M->num_cols, M->nnz, M->cols.p, M->rows.p, M->data.p, d_v, d_y, M->stream);
    return (int)B200_OK;
  });
}

int
b200_mat_mult_vec(b200_mat* M, int nnz_x, const int* idx, const double* val, double* result_dense)
{
  if (!M || !M->have)
  {
    return set_error(B200_ERR_STATE, "matrix not set");
  }
  int rc = check_sparse(nnz_x, idx, val, M->num_cols);
  if (rc != B200_OK)
  {
    return rc;
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    stage_sparse(M, M->num_cols, nnz_x, idx, val, M->x.p);
    launch_spmv(M->num_rows, M->nnz, M->csr_ptr.p, M->csr_col.p, M->csr_val.p, M->x.p, M->y.p, M->stream);
    const bool direct = is_pinned_host(result_dense, sizeof(double) * (size_t)M->num_rows);
    if (M->num_rows > 0)
    {
      B200_CUDA(cudaMemcpyAsync(direct ? result_dense : M->h_out.p, M->y.p, sizeof(double) * (size_t)M->num_rows, cudaMemcpyDeviceToHost, M->stream));
    }
    B200_CUDA(cudaStreamSynchronize(M->stream));
    if (M->num_rows > 0 && !direct)
    {
      std::memcpy(result_dense, M->h_out.p, sizeof(double) * (size_t)M->num_rows);
    }
    return (int)B200_OK;
  });
}

int
b200_mat_mult_vec_trans(b200_mat* M, int nnz_v, const int* idx, const double* val, double* result_dense)
{
  if (!M || !M->have)
  {
    return set_error(B200_ERR_STATE, "matrix not set");
  }
  int rc = check_sparse(nnz_v, idx, val, M->num_rows);
  if (rc != B200_OK)
  {
    return rc;
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    stage_sparse(M, M->num_rows, nnz_v, idx, val, M->x.p);
    launch_spmv(M->num_cols, M->nnz, M->cols.p, M->rows.p, M->data.p, M->x.p, M->y.p, M->stream);
    const bool direct = is_pinned_host(result_dense, sizeof(double) * (size_t)M->num_cols);
    if (M->num_cols > 0)
    {
      B200_CUDA(cudaMemcpyAsync(direct ? result_dense : M->h_out.p, M->y.p, sizeof(double) * (size_t)M->num_cols, cudaMemcpyDeviceToHost, M->stream));
    }
    B200_CUDA(cudaStreamSynchronize(M->stream));
    if (M->num_cols > 0 && !direct)
    {
      std::memcpy(result_dense, M->h_out.p, sizeof(double) * (size_t)M->num_cols);
    }
    return (int)B200_OK;
  });
}

int
b200_mat_mult_vec_trans_sparse(b200_mat* M, int nnz_v, const int* idx, const double* val, double eps, int* idx_out, double* val_out, int* nnz_out)
{
  if (!M || !M->have)
  {
    return set_error(B200_ERR_STATE, "matrix not set");
  }
  if (!nnz_out || (M->num_cols > 0 && (!idx_out || !val_out)))
  {
    return set_error(B200_ERR_ARG, "null output");
  }
  int rc = check_sparse(nnz_v, idx, val, M->num_rows);
  if (rc != B200_OK)
  {
    return rc;
  }
  return guarded([&]() {
    B200_CUDA(cudaSetDevice(M->device));
    const int n = M->num_cols;
    if (n == 0)
    {
      *nnz_out = 0;
      return (int)B200_OK;
    }
    stage_sparse(M, M->num_rows, nnz_v, idx, val, M->x.p);
    launch_spmv(n, M->nnz, M->cols.p, M->rows.p, M->data.p, M->x.p, M->y.p, M->stream);
    // the result is sparsified on the device (mat.c:355-358: entries with |s| <= eps are dropped) and only the kept
    // entries cross the bus: a product with a few violated-constraint multipliers touches a few columns
    const int nchunks = compact_chunks(n);
    M->cp_idx.reserve((size_t)n);
    M->cp_val.reserve((size_t)n);
    M->cp_cnt.reserve((size_t)nchunks + 1);
    M->h_cnt.reserve(2);
    LaunchCounter eager;
    enqueue_compact(M->y.p, n, eps, M->cp_cnt.p, M->cp_idx.p, M->cp_val.p, M->stream, eager);
    B200_CUDA(cudaMemcpyAsync(M->h_cnt.p, M->cp_cnt.p + nchunks, sizeof(int), cudaMemcpyDeviceToHost, M->stream));
    B200_CUDA(cudaStreamSynchronize(M->stream));
    const int kept = M->h_cnt.p[0];
    if (kept > 0)
    {
      B200_CUDA(cudaMemcpyAsync(val_out, M->cp_val.p, sizeof(double) * (size_t)kept, cudaMemcpyDeviceToHost, M->stream));
      B200_CUDA(cudaMemcpyAsync(idx_out, M->cp_idx.p, sizeof(int) * (size_t)kept, cudaMemcpyDeviceToHost, M->stream));
      B200_CUDA(cudaStreamSynchronize(M->stream));
    }
    *nnz_out = kept;
    return (int)B200_OK;
  });
}

int
b200_mat_set_stream(b200_mat* M, void* stream)
{
  if (!M)
  {
    return set_error(B200_ERR_ARG, "null handle");
  }
  cudaSetDevice(M->device);
  cudaStreamSynchronize(M->stream);
  M->stream = stream ? (cudaStream_t)stream : M->own_stream;
  return B200_OK;
}

void*
b200_mat_stream(b200_mat* M)
{
  return M ? (void*)M->stream : nullptr;
}

int
b200_mat_free(b200_mat** handle)
{
  if (!handle || !*handle)
  {
    return B200_OK;
  }
  b200_mat* M = *handle;
  cudaSetDevice(M->device);
  if (M->stream)
  {
    cudaStreamSynchronize(M->stream);
  }
  if (M->own_stream)
  {
    cudaStreamDestroy(M->own_stream);
  }
  delete M;
  *handle = nullptr;
  return B200_OK;
}

} // extern "C"
