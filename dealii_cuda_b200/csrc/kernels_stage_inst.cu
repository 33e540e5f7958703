// kernels_stage_inst.cu -- instantiation + dispatch of the staged cell kernel (one TU per dtype).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "kernels_stage.cuh"

#ifndef MFG_INST_F64
#error "compile with -DMFG_INST_F64=0|1"
#endif

namespace mfg {

#if MFG_INST_F64
typedef double inst_number;
#else
typedef float inst_number;
#endif

template <int n, typename Number>
static void launch_n(const uint32_t *gdesc, const uint32_t *halo, const uint16_t *ptab, int pstride, const uint32_t *class_pat, const Number *cwP,
                     const Number *src, Number *dst, uint32_t n_groups, const double *N, const double *D, int sm_count, cudaStream_t stream,
                     const uint32_t *glist, uint32_t n_list, int mode, bool pdl, bool dep_wait, bool add, int device, bool sync)
{
  using Cfg = StageCfg<n, Number>;
  const uint32_t n_items = mode == 1 ? n_list : n_groups;
  if (n_items == 0) return;
  MFG_REQUIRE(pstride == Cfg::PSTRIDE, "staged cell kernel: the plan was built for another table layout");
  EoMats<Number, n> em;
  make_eo_tables<Number, n>(N, D, em);
  StageClasses cls;
  for (int a = 0; a < 8; ++a) cls.pat[a] = class_pat[a < Cfg::NCLASS ? a : 0];
  auto       kern = sync ? laplace_cell_stage<n, Number, true> : laplace_cell_stage<n, Number, false>;
  // (function attributes are per device: one cache entry per device of the process)
  static int blocks_per_sm[64] = {0};  // (both instantiations have the same resources)
  MFG_REQUIRE(device >= 0 && device < 64, "device index out of range");
  if (blocks_per_sm[device] == 0)
    {
      MFG_CUDA(cudaFuncSetAttribute(laplace_cell_stage<n, Number, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
      MFG_CUDA(cudaFuncSetAttribute(laplace_cell_stage<n, Number, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
      int b = 0;
      MFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, Cfg::WPB * 32, Cfg::SMEM));
      if (b < 1) throw Error(MFG_ERR_CUDA, "staged cell kernel does not fit on an SM");
      blocks_per_sm[device] = b;
    }
  const int reserve = 4;  // CTA slots left to the exchange kernels beside the interior groups of a multi-GPU apply
  const uint32_t full = (uint32_t)(sm_count * blocks_per_sm[device]);
  uint32_t       grid = pdl && !dep_wait && full > (uint32_t)reserve + Cfg::NCLASS ? full - reserve : full;
  if (mode == 1) grid = std::min<uint32_t>(grid, (n_items + Cfg::WPB - 1) / Cfg::WPB);
  else
    {  // class order: CTA b takes the groups g = NCLASS m + b % NCLASS
      const uint32_t per_class = (n_items + Cfg::NCLASS - 1) / Cfg::NCLASS;
      grid = std::min<uint32_t>(grid / Cfg::NCLASS, (per_class + Cfg::WPB - 1) / Cfg::WPB) * Cfg::NCLASS;
      if (grid == 0) grid = Cfg::NCLASS;
    }
  const uint4 *gd = reinterpret_cast<const uint4 *>(gdesc);
  if (pdl)
    {
      cudaLaunchConfig_t cfg;
      std::memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::WPB * 32); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      MFG_CUDA(cudaLaunchKernelEx(&cfg, kern, gd, halo, ptab, cwP, src, dst, n_groups, em, glist, n_list, mode, cls, (int)dep_wait, (int)add));
    }
  else
    {
      kern<<<grid, Cfg::WPB * 32, Cfg::SMEM, stream>>>(gd, halo, ptab, cwP, src, dst, n_groups, em, glist, n_list, mode, cls, 0, (int)add);
      MFG_CUDA_LAST();
    }
}

template <>
void launch_laplace_stage<inst_number>(int degree, const uint32_t *gdesc, const uint32_t *halo, const uint16_t *ptab, int pstride, const uint32_t *class_pat,
                                       const inst_number *cwP, const inst_number *src, inst_number *dst, uint32_t n_groups, const double *N, const double *D,
                                       int sm_count, cudaStream_t stream, const uint32_t *glist, uint32_t n_list, int mode, bool pdl, bool dep_wait, bool add,
                                       int device, bool sync)
{
#define MFG_STAGE_ARGS gdesc, halo, ptab, pstride, class_pat, cwP, src, dst, n_groups, N, D, sm_count, stream, glist, n_list, mode, pdl, dep_wait, add, device, sync
  switch (degree)
    {
      case 2: launch_n<3, inst_number>(MFG_STAGE_ARGS); break;
      case 3: launch_n<4, inst_number>(MFG_STAGE_ARGS); break;
      case 4: launch_n<5, inst_number>(MFG_STAGE_ARGS); break;
      case 5: launch_n<6, inst_number>(MFG_STAGE_ARGS); break;
      default: throw Error(MFG_ERR_UNSUPPORTED, "staged cell kernel: degree must be in 2..5");
    }
#undef MFG_STAGE_ARGS
}

#if MFG_INST_F64
bool stage_supported(int dim, int degree, mfg_dtype) { return dim == 3 && degree >= 2 && degree <= 5; }

template <int n, typename Number> static StageGeom sgeom()
{
  using Cfg = StageCfg<n, Number>;
  return StageGeom{n, Cfg::CW, Cfg::CW % 2 == 0 ? Cfg::CW / 2 : Cfg::CW, Cfg::XCAP, Cfg::HMAX, Cfg::OCAP, Cfg::LCAP, Cfg::PSTRIDE, Cfg::NCLASS};
}
StageGeom stage_geom(int degree, mfg_dtype dt)
{
  const bool f64 = dt == MFG_F64;
  switch (degree)
    {
      case 2: return f64 ? sgeom<3, double>() : sgeom<3, float>();
      case 3: return f64 ? sgeom<4, double>() : sgeom<4, float>();
      case 4: return f64 ? sgeom<5, double>() : sgeom<5, float>();
      case 5: return f64 ? sgeom<6, double>() : sgeom<6, float>();
      default: throw Error(MFG_ERR_UNSUPPORTED, "staged cell kernel: degree must be in 2..5");
    }
}
#endif

}  // namespace mfg
