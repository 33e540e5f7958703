// emu_stubs.cc -- TEST INFRASTRUCTURE: what libmfgpu_emu.so has in place of the units that are not built for the CPU emulation
// (tests/emu/build_emu_lib.py): the staged cell kernel (variant 40; inline PTX).  It reports "unsupported".
#include "kernels_slab3.cuh"
#include "kernels_stage.cuh"
#include "operators.cuh"

namespace mfg {

static void unsupported(const char *what) { throw Error(MFG_ERR_UNSUPPORTED, std::string(what) + " is not part of the CPU emulation build"); }

bool      stage_supported(int, int, mfg_dtype) { return false; }
StageGeom stage_geom(int, mfg_dtype) { unsupported("the staged kernel"); return StageGeom(); }

template <typename Number>
void launch_laplace_stage(int, const uint32_t *, const uint32_t *, const uint16_t *, int, const uint32_t *, const Number *, const Number *, Number *, uint32_t,
                          const double *, const double *, int, cudaStream_t, const uint32_t *, uint32_t, int, bool, bool, bool, int, bool)
{
  unsupported("the staged kernel");
}
template void launch_laplace_stage<float>(int, const uint32_t *, const uint32_t *, const uint16_t *, int, const uint32_t *, const float *, const float *, float *, uint32_t,
                                          const double *, const double *, int, cudaStream_t, const uint32_t *, uint32_t, int, bool, bool, bool, int, bool);
template void launch_laplace_stage<double>(int, const uint32_t *, const uint32_t *, const uint16_t *, int, const uint32_t *, const double *, const double *, double *, uint32_t,
                                           const double *, const double *, int, cudaStream_t, const uint32_t *, uint32_t, int, bool, bool, bool, int, bool);

}  // namespace mfg
