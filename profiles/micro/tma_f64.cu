// Microtest: one 16 x 16 box of doubles through cp.async.bulk.tensor.2d, tensor map in global memory.
// usage: tma_f64 <swizzle: 0 none | 3 128B> <h> <k> <row0> <col0> <map in: 0 global | 1 grid constant>
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

struct alignas(64) Map
{
  unsigned long long o[16];
};

__device__ __forceinline__ unsigned
s32(const void* p)
{
  return (unsigned)__cvta_generic_to_shared(p);
}

__device__ void
body(const void* tmap, int row0, int col0, double* out, bool fence)
{
  __shared__ __align__(1024) double box[256];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(s32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    if (fence)
    {
      asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;\n" ::"l"(tmap) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(&bar)), "r"(2048u) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(s32(box)), "l"(tmap), "r"(row0),
                 "r"(col0), "r"(s32(&bar))
                 : "memory");
  }
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(s32(&bar)), "r"(0u)
               : "memory");
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
  {
    out[i] = box[i];
  }
}

__global__ void
k_global(const Map* tmap, int row0, int col0, double* out)
{
  body(tmap, row0, col0, out, true);
}

__global__ void
k_param(const __grid_constant__ Map tmap, int row0, int col0, double* out)
{
  body(&tmap, row0, col0, out, false);
}

int
main(int argc, char** argv)
{
  const int swz = argc > 1 ? atoi(argv[1]) : 3, h = argc > 2 ? atoi(argv[2]) : 64, k = argc > 3 ? atoi(argv[3]) : 64;
  const int row0 = argc > 4 ? atoi(argv[4]) : 0, col0 = argc > 5 ? atoi(argv[5]) : 0, mode = argc > 6 ? atoi(argv[6]) : 0;
  const int ld = (h + 1) & ~1;
  std::vector<double> A((size_t)ld * k);
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < ld; ++i)
      A[(size_t)j * ld + i] = i < h ? 1000.0 * j + i : -1.0;
  double *dA, *dout;
  cudaMalloc(&dA, A.size() * 8);
  cudaMalloc(&dout, 256 * 8);
  cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                          CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Map m;
  cuuint64_t dims[2] = {(cuuint64_t)h, (cuuint64_t)k}, str[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, 16}, es[2] = {1, 1};
  CUresult rc = ((Enc)fn)((CUtensorMap*)&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swz,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d swizzle=%d h=%d k=%d ld=%d row0=%d col0=%d mode=%d\n", (int)rc, swz, h, k, ld, row0, col0, mode);
  Map* dm;
  cudaMalloc(&dm, sizeof(Map));
  cudaMemcpy(dm, &m, sizeof(Map), cudaMemcpyHostToDevice);
  if (mode == 0)
    k_global<<<1, 128>>>(dm, row0, col0, dout);
  else
    k_param<<<1, 128>>>(m, row0, col0, dout);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess)
  {
    std::vector<double> o(256);
    cudaMemcpy(o.data(), dout, 256 * 8, cudaMemcpyDeviceToHost);
    // unswizzled expectation: element (kk, r) = 1000 (col0 + kk) + row0 + r at kk * 16 + (swz ? ((r/2)^(kk%8))*2 + r%2 : r)
    int bad = 0;
    for (int kk = 0; kk < 16; ++kk)
      for (int r = 0; r < 16; ++r)
      {
        const int pos = kk * 16 + (swz == 3 ? ((((r >> 1) ^ (kk & 7)) << 1) | (r & 1)) : r);
        const int gi = row0 + r, gj = col0 + kk;
        const double want = (gi < h && gj < k && gi >= 0 && gj >= 0) ? 1000.0 * gj + gi : 0.0;
        bad += o[pos] != want;
      }
    printf("mismatches: %d (first values %g %g %g %g)\n", bad, o[0], o[1], o[2], o[16]);
  }
  return 0;
}
