// kernels_general_inst.cu -- instantiation + dispatch of the general-geometry cell kernel (one TU per dim and dtype).
#include "kernels_general.cuh"

#ifndef MFG_INST_DIM
#error "compile with -DMFG_INST_DIM=2|3 -DMFG_INST_F64=0|1"
#endif

namespace mfg {

#if MFG_INST_F64
typedef double inst_number;
#else
typedef float inst_number;
#endif

template <int dim, int n, typename Number>
static void launch_gen(bool atomic, const uint32_t *idx, const Number *gsym, const Number *src, Number *dst, uint32_t cell_begin, uint32_t cell_end,
                       const double *N, const double *D, cudaStream_t stream)
{
  if (cell_end <= cell_begin) return;
  ShapeMats<Number, n> sh;
  for (int i = 0; i < n * n; ++i) { sh.N[i] = (Number)N[i]; sh.D[i] = (Number)D[i]; }
  constexpr int    NPC = ipow(n, dim), CPB = gen_cells_per_block(dim, n);
  constexpr int    threads = ((NPC * CPB + 31) / 32) * 32;
  constexpr size_t smem = sizeof(Number) * (1 + dim) * NPC * CPB;
  static_assert(threads <= 1024, "one thread per tensor entry");
  const uint32_t blocks = (cell_end - cell_begin + CPB - 1) / CPB;
  auto ka = laplace_cell_general<dim, n, Number, true>;
  auto kc = laplace_cell_general<dim, n, Number, false>;
  static bool attr_set[64] = {false};  // (function attributes are per device)
  int dev = 0;
  MFG_CUDA(cudaGetDevice(&dev));
  MFG_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  if (!attr_set[dev] && smem > 48 * 1024)
    {
      MFG_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MFG_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[dev] = true;
    }
  if (atomic) ka<<<blocks, threads, smem, stream>>>(idx, gsym, src, dst, cell_begin, cell_end, sh);
  else kc<<<blocks, threads, smem, stream>>>(idx, gsym, src, dst, cell_begin, cell_end, sh);
  MFG_CUDA_LAST();
}

template <int dim, typename Number>
void launch_laplace_general_dim(int degree, bool atomic, const uint32_t *idx, const Number *gsym, const Number *src, Number *dst, uint32_t cell_begin,
                                uint32_t cell_end, const double *N, const double *D, cudaStream_t stream)
{
  switch (degree)
    {
#define MFG_CASE(P) \
  case P: launch_gen<dim, P + 1, Number>(atomic, idx, gsym, src, dst, cell_begin, cell_end, N, D, stream); break;
      MFG_CASE(1) MFG_CASE(2) MFG_CASE(3) MFG_CASE(4) MFG_CASE(5) MFG_CASE(6) MFG_CASE(7) MFG_CASE(8)
#undef MFG_CASE
      default: throw Error(MFG_ERR_UNSUPPORTED, "degree must be in 1..8");
    }
}

template void launch_laplace_general_dim<MFG_INST_DIM, inst_number>(int, bool, const uint32_t *, const inst_number *, const inst_number *, inst_number *,
                                                                    uint32_t, uint32_t, const double *, const double *, cudaStream_t);

}  // namespace mfg
