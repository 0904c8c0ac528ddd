// Dependent-chain latencies that bound the segment walk of the sparse-subtree sweeps (sst.cu), one thread:
//   DFMA -> DFMA, LDS -> LDS (pointer chase in shared memory), STS -> LDS of the same address, LDS.U16 -> LDS.64 (index -> value)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_latency chain_latency.cu && ./chain_latency
#include <cstdio>
#include <cuda_runtime.h>

__global__ void
k(long long* out, double a0, int n)
{
  __shared__ int next[1024];
  __shared__ double val[1024];
  __shared__ unsigned short idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x)
  {
    next[i] = (i + 17) & 1023;
    val[i]  = 1.0 + 1e-9 * i;
    idx[i]  = (unsigned short)((i + 33) & 1023);
  }
  __syncthreads();
  if (threadIdx.x != 0)
  {
    return;
  }
  long long t0 = clock64();
  double a     = a0;
  for (int i = 0; i < n; ++i)
  {
    a = fma(a, 1.0000001, 1e-9);
  }
  long long t1 = clock64();
  int p        = 0;
  for (int i = 0; i < n; ++i)
  {
    p = next[p];
  }
  long long t2 = clock64();
  volatile double* v = val;
  double s           = a;
  for (int i = 0; i < n; ++i)
  {
    v[i & 1023] = s;
    s           = v[i & 1023] + 1.0;
  }
  long long t3 = clock64();
  int q        = p;
  double acc   = s;
  for (int i = 0; i < n; ++i)
  {
    q = idx[q];
    acc += val[q];
    q = (q + (int)acc) & 1023; // the value feeds the next index: u16 load -> f64 load -> chain
  }
  long long t4 = clock64();
  out[0]       = t1 - t0;
  out[1]       = t2 - t1;
  out[2]       = t3 - t2;
  out[3]       = t4 - t3;
  out[4]       = (long long)a + p + (long long)s + q + (long long)acc;
}

int
main()
{
  long long* d;
  cudaMalloc(&d, 64);
  const int n = 4096;
  for (int rep = 0; rep < 2; ++rep)
  {
    k<<<1, 64>>>(d, 1.0, n);
  }
  long long h[5];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("cycles per step: DFMA chain %.1f, LDS pointer chase %.1f, STS->LDS same address (+DADD) %.1f, LDS.U16 -> LDS.64 -> DADD -> F2I chain %.1f\n", (double)h[0] / n, (double)h[1] / n,
         (double)h[2] / n, (double)h[3] / n);
  return 0;
}
