// Throughput of small 1-D TMA operations at arbitrary (unaligned) element coordinates on an FP64 vector:
//   cp.async.bulk.tensor.1d          global -> shared  (gather of a contiguous DoF run)
//   cp.reduce.async.bulk.tensor.1d   shared -> global  add.f64 (scatter-add of a run)
// box = 2, 4, 10, 28 doubles (run lengths 1, 3, 9, 27 of a Q4 cell padded to 16-byte multiples).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_small tma_small.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned sptr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

#define BATCH 64
// every issuing warp: lane 0 issues `ops` operations at pseudo-random coordinates
template <bool REDUCE, int MODE> __global__ void k(const __grid_constant__ CUtensorMap tm_param, const CUtensorMap *tm_glob, double *vec, int box, int ops, unsigned n, int issuers, unsigned long long *cycles)
{
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *buf = reinterpret_cast<double *>(smem) + warp * 64;  // 512 B per warp (all operations of a warp use the same buffer)
  for (int i = lane; i < 64; i += 32) buf[i] = 1.0;
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sptr(&bar[warp])));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const CUtensorMap *tmp = MODE == 1 ? tm_glob : &tm_param;
  const long long t0 = clock64();
  if (lane == 0 && warp < issuers)
    {
      unsigned s = (blockIdx.x * 32 + warp) * 2654435761u + 12345u;
      unsigned phase = 0;
      for (int it = 0; it < ops; it += BATCH)
        {
          if (!REDUCE) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(&bar[warp])), "r"(BATCH * box * 8) : "memory");
#pragma unroll 8
          for (int u = 0; u < BATCH; ++u)
            {
              s = s * 1664525u + 1013904223u;
              const int c = (int)((s >> 4) % (n - 64));
              if (MODE == 2)
                {
                  double *g = vec + (c & ~1);
                  if (REDUCE)
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(g), "r"(sptr(buf)), "r"(box * 8) : "memory");
                  else
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sptr(buf)), "l"(g), "r"(box * 8),
                                 "r"(sptr(&bar[warp]))
                                 : "memory");
                }
              else if (REDUCE)
                asm volatile("cp.reduce.async.bulk.tensor.1d.global.shared::cta.add.tile.bulk_group [%0, {%1}], [%2];" ::"l"(tmp), "r"(c), "r"(sptr(buf))
                             : "memory");
              else
                asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(sptr(buf)),
                             "l"(tmp), "r"(c), "r"(sptr(&bar[warp]))
                             : "memory");
            }
          if (REDUCE)
            {
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
          else
            {
              asm volatile("{\n.reg .pred p;\nW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(sptr(&bar[warp])),
                           "r"(phase)
                           : "memory");
              phase ^= 1;
            }
        }
      if (REDUCE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main(int argc, char **argv)
{
  const int only_reduce = argc > 1 ? atoi(argv[1]) : -1, only_box = argc > 2 ? atoi(argv[2]) : -1, dt = argc > 3 ? atoi(argv[3]) : 0, nops = argc > 4 ? atoi(argv[4]) : 2048, mode = argc > 5 ? atoi(argv[5]) : 0;
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qr));
  const unsigned n = 16974593;
  double *v;
  CK(cudaMalloc(&v, (size_t)n * 8));
  unsigned long long *cyc;
  CK(cudaMalloc(&cyc, 148 * 8));
  const int ops = nops;
  for (int reduce = 0; reduce < 2; ++reduce)
    for (int box : {2, 4, 10, 28})
      for (int issuers : {1, 2, 4, 8})
        {
          if ((only_reduce >= 0 && reduce != only_reduce) || (only_box >= 0 && box != only_box)) continue;
          CUtensorMap tm;
          cuuint64_t gdim[1] = {n};
          cuuint64_t gstr[1] = {0};
          cuuint32_t bdim[1] = {(cuuint32_t)box};
          cuuint32_t estr[1] = {1};
          CUresult r = encode(&tm, dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : dt == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT64, 1, v, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode failed %d (box %d)\n", (int)r, box); continue; }
          CK(cudaMemset(v, 0, (size_t)n * 8));
          cudaEvent_t e0, e1;
          cudaEventCreate(&e0);
          cudaEventCreate(&e1);
          cudaEventRecord(e0);
          CUtensorMap *tmg;
          CK(cudaMalloc(&tmg, sizeof(tm)));
          CK(cudaMemcpy(tmg, &tm, sizeof(tm), cudaMemcpyHostToDevice));
#define L(R, M) k<R, M><<<148, 256, 8 * 512>>>(tm, tmg, v, box, ops, n, issuers, cyc)
          if (mode == 0) { if (reduce) L(true, 0); else L(false, 0); }
          if (mode == 1) { if (reduce) L(true, 1); else L(false, 1); }
          if (mode == 2) { if (reduce) L(true, 2); else L(false, 2); }
          cudaEventRecord(e1);
          CK(cudaDeviceSynchronize());
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          unsigned long long h[148];
          CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
          double avg = 0;
          for (int i = 0; i < 148; ++i) avg += h[i];
          avg /= 148;
          double sum = 0;
          if (reduce)
            {
              std::vector<double> hv(n);
              CK(cudaMemcpy(hv.data(), v, (size_t)n * 8, cudaMemcpyDeviceToHost));
              for (unsigned i = 0; i < n; ++i) sum += hv[i];
            }
          printf("%s box %2d doubles, %d issuing warps/SM: %7.1f cycles/op/SM (%.3f ms, %.1f Mops/s chip)%s", reduce ? "reduce-add" : "load      ", box, issuers,
                 avg / (ops * issuers), ms, 148.0 * ops * issuers / ms * 1e-3, reduce ? "" : "\n");
          if (reduce) printf("  checksum %.0f expected %.0f\n", sum, 148.0 * ops * issuers * box);
        }
  return 0;
}
