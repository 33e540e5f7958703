// kernels_slab3_inst.cu -- instantiation + dispatch of the slab3 kernel (one TU per dtype).
#include <cstdlib>
#include <cstring>
#include "kernels_slab3.cuh"

#ifndef MFG_INST_F64
#error "compile with -DMFG_INST_F64=0|1"
#endif

namespace mfg {

#if MFG_INST_F64
typedef double inst_number;
#else
typedef float inst_number;
#endif

template <int n, typename Number, bool ASYNC, bool EARLY, bool DOT = false>
static void launch_f(const uint32_t *idxP, const Number *cwP, const Number *src, Number *dst, uint32_t n_groups, const double *N, const double *D,
                     int sm_count, cudaStream_t stream, const uint32_t *mergeP, const uint32_t *glist, bool pdl, int dep_wait, int device,
                     const uint32_t *clist, uint32_t n_clist, double *dot_out, uint32_t *n_dot)
{
  using Cfg = Slab3Cfg<n, Number, ASYNC>;
  if (n_groups == 0) return;
  EoMats<Number, n> em;
  make_eo_tables<Number, n>(N, D, em);
  auto       kern = laplace_cell_slab3<n, Number, ASYNC, EARLY, DOT>;
  static int blocks_per_sm[64] = {0};  // (function attributes are per device)
  MFG_REQUIRE(device >= 0 && device < 64, "device index out of range");
  if (blocks_per_sm[device] == 0)
    {
      MFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
      int b = 0;
      MFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kern, Cfg::WPB * 32, Cfg::SMEM));
      if (b < 1) throw Error(MFG_ERR_CUDA, "slab3 kernel does not fit on an SM");
      blocks_per_sm[device] = b;
    }
  const uint32_t want = (n_groups + Cfg::WPB - 1) / Cfg::WPB;
  // pdl without dep_wait = interior groups of a multi-GPU apply: a few CTA slots stay free for the exchange kernels
  const uint32_t reserve = 4, full = (uint32_t)(sm_count * blocks_per_sm[device]);
  const uint32_t grid = std::min<uint32_t>(want, pdl && !dep_wait && full > reserve + 1 ? full - reserve : full);
  if (n_dot) *n_dot = std::min<uint32_t>(grid * Cfg::WPB, n_groups);  // warps with work = partial sums written
  if (pdl)
    {
      cudaLaunchConfig_t cfg;
      std::memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::WPB * 32); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      MFG_CUDA(cudaLaunchKernelEx(&cfg, kern, idxP, cwP, src, dst, n_groups, em, mergeP, glist, dep_wait, clist, n_clist, dot_out));
    }
  else
    {
      kern<<<grid, Cfg::WPB * 32, Cfg::SMEM, stream>>>(idxP, cwP, src, dst, n_groups, em, mergeP, glist, 0, clist, n_clist, dot_out);
      MFG_CUDA_LAST();
    }
}

template <int n, typename Number>
static void launch_n(const uint32_t *idxP, const Number *cwP, const Number *src, Number *dst, uint32_t n_groups, const double *N, const double *D,
                     int sm_count, cudaStream_t stream, const uint32_t *mergeP, const uint32_t *glist, bool pdl, int dep_wait, int device, int flavour,
                     const uint32_t *clist, uint32_t n_clist, double *dot_out, uint32_t *n_dot)
{
  if (dot_out)
    {
      if (flavour & 1) throw Error(MFG_ERR_UNSUPPORTED, "slab3 kernel: the fused dot product exists for the register gather only");
#define MFG_B2 idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, mergeP, glist, pdl, dep_wait, device, clist, n_clist, dot_out, n_dot
      if (flavour & 2) launch_f<n, Number, false, true, true>(MFG_B2);
      else launch_f<n, Number, false, false, true>(MFG_B2);
#undef MFG_B2
      return;
    }
#define MFG_B idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, mergeP, glist, pdl, dep_wait, device, clist, n_clist, dot_out, n_dot
  switch (flavour & 3)
    {
      case 0: launch_f<n, Number, false, false>(MFG_B); break;
      case 1: launch_f<n, Number, true, false>(MFG_B); break;
      case 2: launch_f<n, Number, false, true>(MFG_B); break;
      default: launch_f<n, Number, true, true>(MFG_B); break;
    }
#undef MFG_B
}

template <>
void launch_laplace_slab3<inst_number>(int degree, const uint32_t *idxP, const inst_number *cwP, const inst_number *src, inst_number *dst,
                                       uint32_t n_groups, const double *N, const double *D, int sm_count, cudaStream_t stream, const uint32_t *mergeP,
                                       const uint32_t *glist, bool pdl, int dep_wait, int device, int flavour, const uint32_t *clist, uint32_t n_clist, double *dot_out,
                                       uint32_t *n_dot)
{
#define MFG_A idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, mergeP, glist, pdl, dep_wait, device, flavour, clist, n_clist, dot_out, n_dot
  switch (degree)
    {
      case 1: launch_n<2, inst_number>(MFG_A); break;
      case 2: launch_n<3, inst_number>(MFG_A); break;
      case 3: launch_n<4, inst_number>(MFG_A); break;
      case 4: launch_n<5, inst_number>(MFG_A); break;
      case 5: launch_n<6, inst_number>(MFG_A); break;
      default: throw Error(MFG_ERR_UNSUPPORTED, "slab3 kernel: degree must be in 1..5");
    }
#undef MFG_A
}

#if MFG_INST_F64
bool slab2_supported(int dim, int degree, mfg_dtype) { return dim == 3 && degree >= 1 && degree <= 5; }

template <int n, int WB> static Slab2Geom geom_of()
{
  using Tab = Slab2Tab<n, WB>;
  constexpr int CW = 32 / n;
  return Slab2Geom{n, CW, CW % 2 == 0 ? CW / 2 : CW, Tab::F, Tab::BC()};
}
Slab2Geom slab2_geom(int degree, mfg_dtype dt)
{
  const bool f64 = dt == MFG_F64;
  switch (degree)
    {
      case 1: return f64 ? geom_of<2, 8>() : geom_of<2, 4>();
      case 2: return f64 ? geom_of<3, 8>() : geom_of<3, 4>();
      case 3: return f64 ? geom_of<4, 8>() : geom_of<4, 4>();
      case 4: return f64 ? geom_of<5, 8>() : geom_of<5, 4>();
      case 5: return f64 ? geom_of<6, 8>() : geom_of<6, 4>();
      default: throw Error(MFG_ERR_UNSUPPORTED, "slab kernels: degree must be in 1..5");
    }
}
#endif

}  // namespace mfg
