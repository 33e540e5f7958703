// kernels_v0_inst.cu -- explicit instantiation + dispatch of the generic cell
// kernel.  Compiled once per (MFG_INST_DIM, MFG_INST_F64) so the 64 kernel
// variants build in parallel translation units.
#include "kernels_v0.cuh"

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
static void launch_n(bool atomic, const uint32_t *idx, const Number *cw, const Number *src, Number *dst, uint32_t cell_begin,
                     uint32_t cell_end, const double *N, const double *D, cudaStream_t stream, const uint32_t *hn_mask, const double *hn_weights)
{
  if (cell_end <= cell_begin) return;
  ShapeMats<Number, n> sh;
  for (int i = 0; i < n * n; ++i) { sh.N[i] = (Number)N[i]; sh.D[i] = (Number)D[i]; }
  constexpr int    CPB     = v0_cells_per_block(dim, n);
  constexpr int    threads = v0_block_threads(dim, n);
  constexpr size_t smem    = v0_smem_bytes<Number>(dim, n);
  const uint32_t   blocks  = (cell_end - cell_begin + CPB - 1) / CPB;
  HangingMat<Number, n> hm;
  for (int i = 0; i < n * n; ++i) hm.W[i] = hn_weights ? (Number)hn_weights[i] : Number(0);
  auto             ka      = laplace_cell_v0<dim, n, Number, true, false>;
  auto             kc      = laplace_cell_v0<dim, n, Number, false, false>;
  auto             kh      = laplace_cell_v0<dim, n, Number, true, true>;
  static bool      attr_set[64] = {false};  // (function attributes are per device)
  int dev = 0;
  MFG_CUDA(cudaGetDevice(&dev));
  MFG_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  if (!attr_set[dev] && smem > 48 * 1024)
    {
      MFG_CUDA(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MFG_CUDA(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      MFG_CUDA(cudaFuncSetAttribute(kh, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set[dev] = true;
    }
  if (hn_mask)
    {
      if (!atomic) throw Error(MFG_ERR_UNSUPPORTED, "hanging nodes need the atomic scatter");
      kh<<<blocks, threads, smem, stream>>>(idx, cw, src, dst, cell_begin, cell_end, sh, hn_mask, hm);
    }
  else if (atomic) ka<<<blocks, threads, smem, stream>>>(idx, cw, src, dst, cell_begin, cell_end, sh, nullptr, hm);
  else kc<<<blocks, threads, smem, stream>>>(idx, cw, src, dst, cell_begin, cell_end, sh, nullptr, hm);
  MFG_CUDA_LAST();
}

template <int dim, typename Number>
void launch_laplace_v0_dim(int degree, bool atomic, const uint32_t *idx, const Number *cw, const Number *src, Number *dst,
                           uint32_t cell_begin, uint32_t cell_end, const double *N, const double *D, cudaStream_t stream,
                           const uint32_t *hn_mask, const double *hn_weights)
{
  switch (degree)
    {
#define MFG_CASE(P) \
  case P: launch_n<dim, P + 1, Number>(atomic, idx, cw, src, dst, cell_begin, cell_end, N, D, stream, hn_mask, hn_weights); break;
      MFG_CASE(1) MFG_CASE(2) MFG_CASE(3) MFG_CASE(4) MFG_CASE(5) MFG_CASE(6) MFG_CASE(7) MFG_CASE(8)
#undef MFG_CASE
      default: throw Error(MFG_ERR_UNSUPPORTED, "degree must be in 1..8");
    }
}

template void launch_laplace_v0_dim<MFG_INST_DIM, inst_number>(int, bool, const uint32_t *, const inst_number *, const inst_number *,
                                                               inst_number *, uint32_t, uint32_t, const double *, const double *,
                                                               cudaStream_t, const uint32_t *, const double *);

}  // namespace mfg
