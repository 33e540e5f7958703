// kernels_slab_inst.cu -- instantiation + dispatch of the slab kernel (one TU per dtype).
#include "kernels_slab.cuh"

#ifndef MFG_INST_F64
#error "compile with -DMFG_INST_F64=0|1"
#endif

namespace mfg {

#if MFG_INST_F64
typedef double inst_number;
#else
typedef float inst_number;
#endif

template <int n, typename Number, int MINB_>
static void launch_n(const uint32_t *idx, const Number *cw, const Number *src, Number *dst, uint32_t n_cells, const double *N,
                     const double *D, int sm_count, cudaStream_t stream)
{
  using Cfg = SlabCfg<n, Number, MINB_>;
  if (n_cells == 0) return;
  ShapeMats<Number, n> sh;
  for (int i = 0; i < n * n; ++i) { sh.N[i] = (Number)N[i]; sh.D[i] = (Number)D[i]; }
  const uint32_t n_groups = (n_cells + Cfg::CW - 1) / Cfg::CW;
  auto           kern     = laplace_cell_slab<n, Number, MINB_>;
  static int     blocks_per_sm = 0;
  if (blocks_per_sm == 0)
    {
      MFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
      MFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, Cfg::WPB * 32, Cfg::SMEM));
      if (blocks_per_sm < 1) throw Error(MFG_ERR_CUDA, "slab kernel does not fit on an SM");
    }
  const uint32_t want = (n_groups + Cfg::WPB - 1) / Cfg::WPB;
  const uint32_t grid = std::min<uint32_t>(want, (uint32_t)(sm_count * blocks_per_sm));
  kern<<<grid, Cfg::WPB * 32, Cfg::SMEM, stream>>>(idx, cw, src, dst, n_cells, n_groups, sh);
  MFG_CUDA_LAST();
}

template <>
void launch_laplace_slab<inst_number>(int degree, int min_blocks, const uint32_t *idx, const inst_number *cw, const inst_number *src, inst_number *dst,
                                      uint32_t n_cells, const double *N, const double *D, int sm_count, cudaStream_t stream)
{
  switch (degree)
    {
      case 1: launch_n<2, inst_number, 0>(idx, cw, src, dst, n_cells, N, D, sm_count, stream); break;
      case 2: launch_n<3, inst_number, 0>(idx, cw, src, dst, n_cells, N, D, sm_count, stream); break;
      case 3: launch_n<4, inst_number, 0>(idx, cw, src, dst, n_cells, N, D, sm_count, stream); break;
      case 4:
        if (min_blocks == 2) launch_n<5, inst_number, 2>(idx, cw, src, dst, n_cells, N, D, sm_count, stream);
        else if (min_blocks == 15) launch_n<5, inst_number, 15>(idx, cw, src, dst, n_cells, N, D, sm_count, stream);
        else if (min_blocks == 12) launch_n<5, inst_number, 12>(idx, cw, src, dst, n_cells, N, D, sm_count, stream);
        else launch_n<5, inst_number, 0>(idx, cw, src, dst, n_cells, N, D, sm_count, stream);
        break;
      default: throw Error(MFG_ERR_UNSUPPORTED, "slab kernel: degree must be in 1..4");
    }
}

#if MFG_INST_F64
bool slab_supported(int dim, int degree, mfg_dtype) { return dim == 3 && degree >= 1 && degree <= 4; }
int  slab_cells_per_group(int degree) { return 32 / (degree + 1); }
size_t slab_cw_padded_cells(uint32_t n_cells) { return (size_t)n_cells + 32; }
#endif

}  // namespace mfg
