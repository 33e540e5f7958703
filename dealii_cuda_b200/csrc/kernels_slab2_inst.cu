// kernels_slab2_inst.cu -- instantiation + dispatch of the slab2 kernel (one TU per dtype).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "kernels_slab2_ws.cuh"

#ifndef MFG_INST_F64
#error "compile with -DMFG_INST_F64=0|1"
#endif

namespace mfg {

#if MFG_INST_F64
typedef double inst_number;
#else
typedef float inst_number;
#endif

template <int n, typename Number, int CFG>
static void launch_n(const uint32_t *idxP, const Number *cwP, const Number *src, Number *dst, uint32_t n_groups, const double *N,
                     const double *D, int sm_count, cudaStream_t stream, cudaTextureObject_t tex, const uint32_t *mergeP, const uint32_t *glist, bool pdl, bool dep_wait, const uint32_t *idxLex, const uint32_t *idxJ, uint32_t n_cells)
{
  using Cfg = Slab2Cfg<n, Number, CFG>;
  if (n_groups == 0) return;
  EoMats<Number, n> em;
  make_eo_tables<Number, n>(N, D, em);
  auto       kern          = laplace_cell_slab2<n, Number, CFG>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0)
    {
      MFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
      MFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, Cfg::WPB * 32, Cfg::SMEM));
      if (blocks_per_sm < 1) throw Error(MFG_ERR_CUDA, "slab2 kernel does not fit on an SM");
    }
#ifdef MFG_SLAB2_ABLATE
  {
    int d[2] = {getenv("MFG_SLAB2_DELAY") ? atoi(getenv("MFG_SLAB2_DELAY")) : 0, getenv("MFG_SLAB2_DELAY_MODE") ? atoi(getenv("MFG_SLAB2_DELAY_MODE")) : 0};
    MFG_CUDA(cudaMemcpyToSymbolAsync(g_slab2_delay, d, sizeof(d), 0, cudaMemcpyHostToDevice, stream));
  }
#endif
  const uint32_t want = (n_groups + Cfg::WPB - 1) / Cfg::WPB;
  // pdl = interior groups of a multi-GPU apply: MFG_SLAB2_RESERVE CTA slots are left free for the kernels of the
  // exchange that runs beside it (pack, NCCL send/recv, accumulate): a persistent grid that fills every slot would
  // push them behind its last CTA
  static const int reserve = std::getenv("MFG_SLAB2_RESERVE") ? std::atoi(std::getenv("MFG_SLAB2_RESERVE")) : 4;
  const uint32_t full = (uint32_t)(sm_count * blocks_per_sm);
  const uint32_t grid = std::min<uint32_t>(want, pdl && !dep_wait && full > (uint32_t)reserve + 1 ? full - reserve : full);
  if (pdl)
    {
      // programmatic dependent launch: the CTAs of this launch may start while the kernel before it in the stream (the
      // interface cell groups of the same apply, which trigger at once) is still running; they depend on nothing it writes
      cudaLaunchConfig_t cfg;
      std::memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::WPB * 32); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      MFG_CUDA(cudaLaunchKernelEx(&cfg, kern, idxP, cwP, src, dst, n_groups, em, tex, mergeP, glist, (int)dep_wait, idxLex, idxJ, n_cells));
    }
  else
    {
      kern<<<grid, Cfg::WPB * 32, Cfg::SMEM, stream>>>(idxP, cwP, src, dst, n_groups, em, tex, mergeP, glist, 0, idxLex, idxJ, n_cells);
      MFG_CUDA_LAST();
    }
}

template <int n, typename Number>
static void launch_cfg(int cfg, const uint32_t *idxP, const Number *cwP, const Number *src, Number *dst, uint32_t n_groups, const double *N,
                       const double *D, int sm_count, cudaStream_t stream, cudaTextureObject_t tex, const uint32_t *mergeP, const uint32_t *glist, bool pdl, bool dep_wait, const uint32_t *idxLex, const uint32_t *idxJ, uint32_t n_cells)
{
  switch (cfg)
    {
      case 1: launch_n<n, Number, 1>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 3: launch_n<n, Number, 3>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 4: launch_n<n, Number, 4>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 7: launch_n<n, Number, 7>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 9: launch_n<n, Number, 9>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 11: launch_n<n, Number, 11>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 13: launch_n<n, Number, 13>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 15: launch_n<n, Number, 15>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 17: launch_n<n, Number, 17>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 19: launch_n<n, Number, 19>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 21: launch_n<n, Number, 21>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 23: launch_n<n, Number, 23>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 0: launch_n<n, Number, 0>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 2: launch_n<n, Number, 2>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      // configurations 3 and 7 with the face merge compiled in
      case 259: launch_n<n, Number, 259>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 263: launch_n<n, Number, 263>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      // plane-layout gather / scatter (dense transpose buffers are conflict free for n = 5 only)
      case 515:
        if constexpr (n == 5) launch_n<n, Number, 515>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells);
        else throw Error(MFG_ERR_UNSUPPORTED, "slab2 plane-layout configuration exists for degree 4 only");
        break;
      case 513:
        if constexpr (n == 5) launch_n<n, Number, 513>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells);
        else throw Error(MFG_ERR_UNSUPPORTED, "slab2 plane-layout configuration exists for degree 4 only");
        break;
#ifdef MFG_SLAB2_ABLATE
#define MFG_ABL(a) case 7 + 32 * a: if constexpr (n == 5) launch_n<n, Number, 7 + 32 * a>(idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      MFG_ABL(1) MFG_ABL(2) MFG_ABL(3) MFG_ABL(4) MFG_ABL(5) MFG_ABL(6) MFG_ABL(7)
#undef MFG_ABL
#endif
      default: throw Error(MFG_ERR_UNSUPPORTED, "slab2 kernel: unknown configuration");
    }
}

template <>
void launch_laplace_slab2<inst_number>(int degree, int cfg, const uint32_t *idxP, const inst_number *cwP, const inst_number *src, inst_number *dst,
                                       uint32_t n_groups, const double *N, const double *D, int sm_count, cudaStream_t stream, cudaTextureObject_t tex, const uint32_t *mergeP, const uint32_t *glist, bool pdl, bool dep_wait, const uint32_t *idxLex, const uint32_t *idxJ, uint32_t n_cells)
{
  switch (degree)
    {
      case 1: launch_cfg<2, inst_number>(cfg, idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 2: launch_cfg<3, inst_number>(cfg, idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 3: launch_cfg<4, inst_number>(cfg, idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 4: launch_cfg<5, inst_number>(cfg, idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      case 5: launch_cfg<6, inst_number>(cfg, idxP, cwP, src, dst, n_groups, N, D, sm_count, stream, tex, mergeP, glist, pdl, dep_wait, idxLex, idxJ, n_cells); break;
      default: throw Error(MFG_ERR_UNSUPPORTED, "slab2 kernel: degree must be in 1..5");
    }
}

template <int NCW, int NLW, int MINB, bool LS, int RC = 0, int RL = 0>
static void launch_ws(const uint32_t *idxLex, const uint32_t *idxJ, const inst_number *cwP, const inst_number *src, inst_number *dst, uint32_t n_groups,
                      uint32_t n_cells, const double *N, const double *D, int sm_count, cudaStream_t stream)
{
  constexpr int n = 5;
  using Cfg = Slab2WsCfg<n, inst_number, NCW, NLW, MINB, LS, RC, RL>;
  if (n_groups == 0) return;
  EoMats<inst_number, n> em;
  make_eo_tables<inst_number, n>(N, D, em);
  auto       kern = laplace_cell_slab2_ws<n, inst_number, NCW, NLW, MINB, LS, RC, RL>;
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0)
    {
      MFG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
      MFG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, Cfg::THREADS, Cfg::SMEM));
      if (blocks_per_sm < 1) throw Error(MFG_ERR_CUDA, "warp-specialised slab2 kernel does not fit on an SM");
    }
  const uint32_t want = (n_groups + Cfg::NCW - 1) / Cfg::NCW;
  const uint32_t grid = std::min<uint32_t>(want, (uint32_t)(sm_count * blocks_per_sm));
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, em);
  MFG_CUDA_LAST();
}

// shape: 0 = 2 CTAs x (4 contraction + 2 loader warps), 1 = 3 CTAs x (3 + 1), 2 = 2 CTAs x (4 + 1), 3 = 2 CTAs x (4 + 4)
template <>
void launch_laplace_slab2_ws<inst_number>(int degree, int shape, const uint32_t *idxLex, const uint32_t *idxJ, const inst_number *cwP,
                                          const inst_number *src, inst_number *dst, uint32_t n_groups, uint32_t n_cells, const double *N,
                                          const double *D, int sm_count, cudaStream_t stream)
{
  if (degree != 4) throw Error(MFG_ERR_UNSUPPORTED, "the warp-specialised slab2 kernel exists for degree 4 only");
  switch (shape)
    {
      case 0: launch_ws<4, 2, 2, false>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      case 1: launch_ws<3, 1, 3, false>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      case 2: launch_ws<4, 1, 2, false>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      case 3: launch_ws<4, 4, 1, false>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      // 4, 5: the loader warps also scatter (contraction warps touch global memory only through the coefficient copy)
      case 4: launch_ws<4, 2, 2, true>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      case 5: launch_ws<4, 4, 1, true>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      // 6: two CTAs of 4 + 4 warps at 128 registers; setmaxnreg gives the contraction warps 168, the loader / scatter warps 88
      case 6: launch_ws<4, 4, 2, true, 168, 88>(idxLex, idxJ, cwP, src, dst, n_groups, n_cells, N, D, sm_count, stream); break;
      default: throw Error(MFG_ERR_UNSUPPORTED, "warp-specialised slab2 kernel: unknown shape");
    }
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
      default: throw Error(MFG_ERR_UNSUPPORTED, "slab2 kernel: degree must be in 1..5");
    }
}
#endif

}  // namespace mfg
