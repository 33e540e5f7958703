// Cost of red.global.add.f64 per warp instruction against lane count and address pattern (12 warps per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_cost red_cost.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// pattern: lane -> element offset inside a 4096-double window
__device__ __forceinline__ int lane_off(int pat, int lane)
{
  switch (pat)
    {
      case 0: return lane;            // contiguous: 8 sectors, 2 lines
      case 1: return lane * 4;        // one sector per lane: 32 sectors, 8 lines
      case 2: return lane * 16;       // one line per lane
      case 3: { const int c = lane / 5, i = lane % 5; return c * 160 + (i == 0 ? 0 : i == 4 ? 64 : 31 + i); }  // cell-like: vertex, run of 3, vertex
      case 4: return lane / 4 * 16 + lane % 4;  // 4 lanes per sector, one sector per line: 8 sectors, 8 lines
      default: return lane;
    }
}

template <bool LOAD> __global__ void k(double *v, size_t n, int pat, int active, int iters, double *sink)
{
  const int lane = threadIdx.x & 31;
  const size_t gw = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  unsigned s = (unsigned)gw * 2654435761u + 77u;
  const int off = lane_off(pat, lane);
  double acc = 0;
  for (int it = 0; it < iters; ++it)
    {
      s = s * 1664525u + 1013904223u;
      const size_t base = ((size_t)(s >> 8) % (n / 4096 - 1)) * 4096;
      if (lane < active)
        {
          if (LOAD) acc += __ldg(v + base + off);
          else atomicAdd(v + base + off, 1.0);
        }
    }
  if (acc == 1.2345) *sink = acc;
}

int main()
{
  const size_t n = 16974593;
  double *v, *sink;
  CK(cudaMalloc(&v, n * 8));
  CK(cudaMalloc(&sink, 8));
  CK(cudaMemset(v, 0, n * 8));
  const int iters = 2000, warps = 12;
  for (int load = 0; load < 2; ++load)
    for (int pat = 0; pat < 5; ++pat)
      for (int active : {32, 30, 16, 8, 1})
        {
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          if (load) k<true><<<148, warps * 32>>>(v, n, pat, active, 10, sink); else k<false><<<148, warps * 32>>>(v, n, pat, active, 10, sink);
          cudaEventRecord(e0);
          if (load) k<true><<<148, warps * 32>>>(v, n, pat, active, iters, sink); else k<false><<<148, warps * 32>>>(v, n, pat, active, iters, sink);
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          printf("%s pattern %d active lanes %2d: %6.1f cycles per warp instruction per SM (at 1.9 GHz), %.3f ms\n", load ? "ldg" : "red", pat, active,
                 ms * 1e-3 * 1.9e9 / (iters * warps), ms);
        }
  return 0;
}
