// mesh.cuh -- the uniform box mesh + FE_Q DoF substrate (device resident).
// Stands in for what the reference gets from deal.II: Triangulation
// (hyper_cube + refine_global, poisson_common.h:58-72, bmop_common.h:108-120),
// DoFHandler::distribute_dofs (bmop.cu:116), the lexicographic renumbering of
// cell DoFs (matrix_free_gpu.cu:283-300) and the boundary ConstraintMatrix
// (bmop.cu:118-124).
#pragma once
#include "common.cuh"
#include "fe_data.h"

struct mfg_mesh
{
  mfg_ctx *ctx = nullptr;
  int      dim = 0, p = 0, n = 0;
  int      lg[3] = {0, 0, 0};     // log2 cells per direction
  uint32_t nc[3] = {1, 1, 1};     // cells per direction
  double   origin[3] = {0, 0, 0}, h = 0;
  uint32_t dirichlet_faces = 0;
  uint32_t n_cells = 0, n_dofs = 0, npc = 0, n_constrained = 0;
  mfg::DevBuf<uint32_t> l2g;         // [n_cells][npc] lexicographic
  mfg::DevBuf<uint32_t> cell_first;  // [n_cells] first DoF numbered by the cell
  mfg::DevBuf<uint16_t> rank_table;  // [8][npc]
  mfg::DevBuf<uint32_t> constrained; // ascending
  mfg::DevBuf<uint8_t>  cflag;       // [n_dofs] 1 = constrained
  mfg::FEData1D         fe;
};

namespace mfg {

// Morton (Z-order) <-> cell coordinates with per-direction bit counts; x is the
// least significant bit (deal.II child order: child = x + 2y + 4z).
struct MortonMap
{
  int dim, lg[3];
  __host__ __device__ inline void decode(uint32_t c, uint32_t x[3]) const
  {
    x[0] = x[1] = x[2] = 0;
    int pos = 0;
    for (int b = 0; b < 11; ++b)
      for (int d = 0; d < dim; ++d)
        if (b < lg[d]) { x[d] |= ((c >> pos) & 1u) << b; ++pos; }
  }
  __host__ __device__ inline uint32_t encode(const uint32_t x[3]) const
  {
    uint32_t c = 0; int pos = 0;
    for (int b = 0; b < 11; ++b)
      for (int d = 0; d < dim; ++d)
        if (b < lg[d]) { c |= ((x[d] >> b) & 1u) << pos; ++pos; }
    return c;
  }
};

// hierarchic (deal.II FE_Q) -> lexicographic local numbering, built by
// enumerating vertices, lines, quads, hex with deal.II's reference-cell conventions
std::vector<uint32_t> hierarchic_to_lexicographic(int dim, int p);

}  // namespace mfg
