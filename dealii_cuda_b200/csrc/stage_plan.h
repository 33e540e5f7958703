// stage_plan.h -- host-side plan of the staged Laplace cell kernel (kernels_stage.cuh).
//
// The staged kernel replaces the per-entry gather (read_dof_values, fee_gpu.cuh:323-338) and the per-entry atomic
// scatter (distribute_local_to_global, fee_gpu.cuh:346-365 + atomic.cuh:11-32) of a warp group of cells by
//   * one coalesced asynchronous copy of the group's OWN DoFs (deal.II numbers the DoFs a cell touches first
//     consecutively, so the cells of a group own one contiguous range of the vector) plus a short list of HALO DoFs
//     (owned by earlier cells) into a shared-memory staging buffer, from which every lane reads its slab;
//   * face merges inside the group in registers, one staged copy of the group's results, then a coalesced write-out:
//     plain stores for the DoFs no cell outside the group touches, red.add for the rest.
// Everything the kernel needs is derived here from the plain index array (loc2glob with the constrained bit), so any
// mesh works; groups whose structure does not fit (staging buffer too small) are left to the slab2 kernel.
#pragma once
#include <cstdint>
#include <vector>

namespace mfg {

constexpr int STAGE_PH      = 16;      // uint16 entries of a pattern header
constexpr uint16_t STAGE_DEAD = 0x8000u;  // pos entry: not written back (handed over to a neighbour cell or constrained)
constexpr uint32_t STAGE_NOPAT = 0xffffu; // gdesc: group is not staged (slab2 kernel processes it)

struct StagePlanIn
{
  int             n = 0;        // points per direction
  int             cw = 0;       // cells per warp group
  int             hc = 0;       // cells per half warp (lane = 16 ch + n cl + i when cw is even, else n c + i)
  int             wb = 8;       // sizeof(Number)
  int             xcap = 0;     // slots of the staging buffer; slot xcap-1 is the zero (input) / trash (output) slot
  int             hmax = 0;     // maximum number of halo entries of a group (32 x registers of the kernel)
  int             ocap = 0;     // capacity of the own range: cw * n^3 rounded up to a multiple of 32
  int             lcap = 0;     // capacity of the load list: ocap + hmax
  uint32_t        n_plain = 0;  // cells [0, n_plain) are processed by the kernel
  uint32_t        n_cells = 0;  // all cells (cells behind n_plain count for the DoF multiplicities)
  uint32_t        n_dofs = 0;
  const uint32_t *idx = nullptr;  // [n_cells][n^3] lexicographic, bit 31 = constrained
  int             merge_dirs = 7; // bit d: merge faces in direction d inside a group
  int             nclass = 1;     // groups g and g + nclass have the same position in the cell order (same tables on a uniform mesh)
};

struct StagePlan
{
  uint32_t              n_groups = 0, n_patterns = 0, n_staged = 0;
  int                   pstride = 0;  // uint16 entries per pattern: header [16] | pos, rows in pairs: uint32 [ceil(n^2/2)][32] |
                                      // own range: uint32 slot | flag << 16 [ocap] | halo slots: uint16 [hmax]
  std::vector<uint32_t> gdesc;        // [n_groups][4]: own_base, halo_off, n_halo | pattern << 16, merge mask
  std::vector<uint32_t> halo;         // halo DoF lists of all groups
  std::vector<uint16_t> ptab;         // [n_patterns][pstride]
  std::vector<uint32_t> fallback;     // groups the staged kernel skips
  uint32_t              class_pat[8];  // most frequent pattern among the groups g % nclass == a (STAGE_NOPAT: none)
  // statistics (per staged group averages x 1000 are computed by the caller)
  uint64_t n_own = 0, n_halo = 0, n_plain_dofs = 0, n_red_dofs = 0, rd_wavefronts = 0, wr_wavefronts = 0, cp_wavefronts = 0;
};

// pattern header fields
enum { STAGE_H_OWN = 0, STAGE_H_NHALO = 1 };

void build_stage_plan(const StagePlanIn &in, StagePlan &out);

}  // namespace mfg
