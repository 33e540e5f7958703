// DFMA issue-rate microbenchmark: FP64 throughput of one SM sub-partition against warps per scheduler and ILP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double *out, int iters, double a, double b)
{
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}
template <int ILP> void run(int warps_per_sm)
{
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 4000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<ILP><<<148, warps_per_sm * 32>>>(out, 10, 1.0000001, 1e-9);
  cudaEventRecord(e0);
  k<ILP><<<148, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fmas = 148.0 * warps_per_sm * 32 * iters * 8.0 * ILP;
  printf("ILP %2d warps/SM %2d (per scheduler %4.1f): %7.2f TFLOP/s  (%.1f DFMA lanes/clk/SM at 1.9 GHz)\n", ILP, warps_per_sm, warps_per_sm / 4.0,
         2 * fmas / ms * 1e-9, fmas / (ms * 1e-3) / 148 / 1.9e9);
  cudaFree(out);
}
int main()
{
  for (int w : {4, 8, 12, 16, 32})
    {
      run<1>(w); run<2>(w); run<4>(w); run<8>(w); run<16>(w);
    }
  return 0;
}
