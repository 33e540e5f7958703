// emu_mg_transfer.cc -- TEST INFRASTRUCTURE: the device code of the multigrid transfer kernel (csrc/mg_transfer.cu: pass, valence /
// block weights, mg_kernel -- cut out by tests/test_mg_transfer_emulated.py) on the CPU emulation of tests/emu/cuda_emu.h, launched
// like mgt_run does: one CTA of 128 threads per block (refined coarse cell), dynamic shared memory for two (2p+1)^dim tensors.
#include "cuda_emu.h"
#define __restrict__
alignas(16) double sm[2 * 17 * 17 * 17];   // the kernel's `extern __shared__ double sm[]`
#include "mg_kernel_device_part.h"          // generated

// prolong != 0: dst_fine = P src_coarse (dst zeroed first); else dst_coarse += P^T src_fine
extern "C" int emu_mg_transfer(int prolong, int dim, int degree, uint32_t n_blocks, const uint32_t *coarse_idx, const uint32_t *fine_idx,
                               const double *weights, const uint32_t *cell_xyz, const uint32_t *nc, const double *P1d, const double *src, double *dst,
                               uint32_t n_dst)
{
  MgArgs A;
  A.dim = dim; A.p = degree; A.n_cells = n_blocks;
  for (int d = 0; d < 3; ++d) A.nc[d] = nc ? nc[d] : 1u;
  A.coarse_idx = coarse_idx; A.fine_idx = fine_idx; A.cxyz = cell_xyz; A.wtab = weights;
  PMat pm;
  std::memset(pm.P, 0, sizeof(pm.P));
  std::memcpy(pm.P, P1d, sizeof(double) * (2 * degree + 1) * (degree + 1));
  if (prolong) std::memset(dst, 0, sizeof(double) * n_dst);
  if (n_blocks == 0) return 0;
  if (prolong) emu_launch(n_blocks, 128, mg_kernel<double, true>, A, pm, dst, src);
  else emu_launch(n_blocks, 128, mg_kernel<double, false>, A, pm, dst, src);
  return 0;
}
