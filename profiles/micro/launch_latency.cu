// Microbenchmark: per-node cost of dependent kernels inside a CUDA graph on this GPU (what bounds the
// level-scheduled sweeps from below). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_latency launch_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__global__ void k_empty(int* p) { if (p && threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1; }
__global__ void k_chain(const double* __restrict__ a, double* __restrict__ out, int hops)
{
  // a dependent chain of `hops` global loads per thread (pointer chasing through doubles used as indices)
  extern __shared__ double sm[];
  double v = a[blockIdx.x * blockDim.x + threadIdx.x];
  for (int i = 0; i < hops; ++i) v = a[((long long)v) & 0xFFFFF];
  sm[threadIdx.x] = v; __syncthreads();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sm[threadIdx.x ^ 1];
}
static float run_graph(cudaStream_t s, int nodes, int grid, int smem_lo, int smem_hi, int hops, double* a, double* out, int reps)
{
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < nodes; ++i)
  {
    int smem = smem_lo + (smem_hi - smem_lo) * i / (nodes > 1 ? nodes - 1 : 1);
    if (hops < 0) k_empty<<<grid, 128, smem, s>>>(nullptr);
    else k_chain<<<grid, 128, smem < 1024 ? 1024 : smem, s>>>(a, out, hops);
  }
  cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / reps / nodes;
}
int main()
{
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  double *a, *out; cudaMalloc(&a, 8 << 20); cudaMalloc(&out, 8 << 20);
  std::vector<double> h(1 << 20); for (size_t i = 0; i < h.size(); ++i) h[i] = (double)((i * 2654435761u) & 0xFFFFF);
  cudaMemcpy(a, h.data(), 8 << 20, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  printf("us per graph node, 26 dependent nodes:\n");
  for (int grid : {1, 148, 1184, 13000})
  {
    printf(" grid %5d: empty %.2f | empty smem 1K..12K %.2f | chain 4 hops %.2f | chain 8 hops %.2f | chain 8 hops smem 1K..12K %.2f | chain 16 hops %.2f\n", grid,
           run_graph(s, 26, grid, 0, 0, -1, a, out, 50), run_graph(s, 26, grid, 1024, 12288, -1, a, out, 50), run_graph(s, 26, grid, 1024, 1024, 4, a, out, 50),
           run_graph(s, 26, grid, 1024, 1024, 8, a, out, 50), run_graph(s, 26, grid, 1024, 12288, 8, a, out, 50), run_graph(s, 26, grid, 1024, 1024, 16, a, out, 50));
  }
  return 0;
}
