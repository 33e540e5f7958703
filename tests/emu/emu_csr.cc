// emu_csr.cc -- TEST INFRASTRUCTURE: the CSR kernel of csrc/sparse_matrix.cu (one warp per row, shuffle reduction; cut out by
// tests/test_sparse_matrix.py) on the CPU emulation of tests/emu/cuda_emu.h, launched like mfg_spm_vmult does.
#include "cuda_emu.h"
#define __restrict__
#include "csr_kernel_device_part.h"   // generated

extern "C" int emu_csr_vmult(uint32_t n, const uint32_t *row_ptr, const uint32_t *col, const double *val, const double *x, double *y)
{
  const unsigned threads = 256, rows_per_block = threads / 32, blocks = (n + rows_per_block - 1) / rows_per_block;
  emu_launch(blocks, threads, csr_vmult_warp_per_row<double>, n, row_ptr, col, val, x, y);
  return 0;
}
