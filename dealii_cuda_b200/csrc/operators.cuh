// operators.cuh -- host objects behind the C ABI handles:
//   mfg_mf      <-> MatrixFreeGpu<dim,Number>          (matrix_free_gpu.h:81-229)
//   mfg_ch      <-> ConstraintHandlerGpu<Number>       (constraint_handler_gpu.h:13-59)
//   mfg_laplace <-> LaplaceOperatorGpu<dim,p,Number>   (laplace_operator_gpu.h:35-96)
#pragma once
#include <memory>
#include "common.cuh"
#include "fe_data.h"
#include "mesh.cuh"
#include "vector.cuh"

struct mfg_mf
{
  mfg_ctx    *ctx = nullptr;
  int         dim = 0, p = 0, n = 0;
  uint32_t    npc = 0, n_cells = 0, n_dofs = 0;
  mfg_dtype   dt = MFG_F64;
  mfg_scatter scatter = MFG_SCATTER_ATOMIC;
  std::vector<uint32_t> color_offsets;  // [n_colors+1] over the (color-sorted) cell order
  mfg::DevBuf<uint32_t> idx;            // [n_cells][npc] lexicographic; bit 31 = constrained DoF
  mfg::DevBuf<uint32_t> cell_perm;      // sorted position -> original cell (empty: identity)
  // hanging nodes: cells [0, n_plain) have mask 0, cells [n_plain, n_cells) carry hn_mask[cell]
  uint32_t              n_plain = 0;
  mfg::DevBuf<uint32_t> hn_mask;        // [n_cells] in kernel cell order (empty: no hanging nodes)
  mfg::FEData1D         fe;
  // geometry: either a uniform mesh (origin/h/Morton map) or host arrays from the explicit description
  const mfg_mesh       *mesh = nullptr;         // not owned
  std::vector<double>   geom_host;              // [n_cells][npc]: inv_jac^2 * JxW_q, original cell order
  bool                  general = false;        // MFG_GEOM_GENERAL: full J^-1 per quadrature point
  std::vector<double>   gsym_host;              // general: [n_cells][dim(dim+1)/2][npc]: JxW_q K K^T (00, 11[, 22], 01[, 02, 12])
  std::vector<double>   qpoints_host;           // optional [n_cells][npc][dim]
  // host copies of the geometry as given in mfg_mf_desc (original cell order) and the device arrays of the generic
  // FEEvaluationGpu path (mfg_mf_get_gpu_data), built on first request in kernel cell order
  std::vector<double>   jxw_host, invjac_host;
  mfg::DevBuf<uint8_t>  gd_jxw, gd_invjac, gd_qpts;
  uint32_t              n_colors() const { return (uint32_t)color_offsets.size() - 1; }
};

struct mfg_ch
{
  mfg_ctx  *ctx = nullptr;
  mfg_dtype dt = MFG_F64;
  mfg::DevBuf<uint32_t> constrained, edge;
  mfg::DevBuf<uint8_t>  tmp_src, tmp_dst;  // constrained_values_src/dst (constraint_handler_gpu.h:51-52)
  size_t                n() const { return constrained.n; }
};

struct mfg_laplace
{
  mfg_ctx *ctx = nullptr;
  mfg_mf  *mf = nullptr;
  mfg_ch  *ch = nullptr;
  bool     owns_mf = false, owns_ch = false;
  mfg::DevBuf<uint8_t>  cw;     // [n_cells][npc] merged weights, operator dtype, color-sorted cell order
  mfg::DevBuf<uint32_t> cbits;  // bit i = DoF i constrained
  std::unique_ptr<mfg_vec> inv_diag;
  bool     diagonal_is_available = false;
  int      variant = 0;
  // options (mfg_laplace_set_option): fused CG loop, face-merge directions of the staged kernel's plan, per-item CTA barrier of the staged kernel
  bool     cg_fused = true;
  int      stage_merge_dirs = 7;
  bool     stage_sync = true;
  // slab kernels (kernels_slab3.cuh, kernels_stage.cuh): the index map and the merged weights in the order their threads
  // consume them, built on first use from idx / cw
  mfg::DevBuf<uint32_t> idxP;   // [n_groups][n^2 slots][32 lanes]
  mfg::DevBuf<uint8_t>  cwP;    // [n_groups][shared-memory image]
  mfg::DevBuf<uint32_t> mergeP; // [n_groups] face-merge mask
  int                   merge_dirs_built = -1;  // directions the mask was built for
  mfg::DevBuf<uint32_t> glist;  // multi-GPU work list: groups touching interface DoFs first (laplace_set_interface_dofs)
  uint32_t              n_iface_groups = 0;
  uint32_t              slab2_groups = 0;
  bool                  cwP_valid = false;
  // staged kernel (kernels_stage.cuh): plan built on first use from idx (stage_plan.cu)
  mfg::DevBuf<uint32_t> st_gdesc, st_halo;   // [n_groups][4] ; halo DoF lists
  mfg::DevBuf<uint16_t> st_ptab;             // deduplicated position tables
  mfg::DevBuf<uint32_t> st_fallback;         // groups the plan leaves to the slab2 kernel (interface groups first)
  uint32_t              st_fb_iface = 0;     // how many of them touch interface DoFs
  int                   st_pstride = 0;
  uint32_t              st_class_pat[8] = {0};  // tables the kernel keeps in shared memory, per class of groups
  bool                  st_built = false;
  uint32_t              st_stats[8] = {0};   // groups, staged, patterns, own, halo, plain, red, smem wavefronts (per staged group, x 16)
  mfg::DevBuf<double>  solver_dot;   // fused CG: per-warp partial sums of d . (A d) written by the cell kernel
  mfg::DevBuf<uint8_t> solver_work;  // mfg_solver_cg: residual, direction, A*direction, device-resident scalars (reused across solves)
  mfg::DevBuf<uint8_t> host_stage_src, host_stage_dst;  // device staging for vmult_host
  // pipelined host API (mfg_laplace_vmult_host_async): 2 slots x {src,dst} staging, copy streams, events
  struct HostSlot { mfg::DevBuf<uint8_t> src, dst; cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr; bool used = false; };
  HostSlot     slots[2];
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  // optional per-launch timing of the cell kernel (bench.py roofline figure)
  bool                     timing = false;
  int                      timing_stride = 1;   // bracket every timing_stride-th cell-kernel launch
  size_t                   timing_counter = 0;
  std::vector<cudaEvent_t> ev;        // pairs (start, stop)
  size_t                   ev_used = 0;
};

namespace mfg {
mfg_mesh *build_box_mesh(mfg_ctx *ctx, const mfg_box_desc &d);
void      mesh_lattice_to_dof(const mfg_mesh *m, size_t npts, const uint32_t *xyz_host, uint32_t *out_host);
void      mesh_cell_coords(const mfg_mesh *m, uint32_t *out_host);
void      mesh_support_points(const mfg_mesh *m, double *out_host);
void      mesh_parity_colors(const mfg_mesh *m, std::vector<uint32_t> &color_of_cell, uint32_t &n_colors);

mfg_mf *mf_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter);
mfg_mf *mf_from_desc(mfg_ctx *ctx, const mfg_mf_desc &d);
mfg_ch *ch_create(mfg_ctx *ctx, mfg_dtype dt, const uint32_t *constrained_host, size_t nc, const uint32_t *edge_host, size_t ne);
mfg_ch *ch_from_mesh(mfg_ctx *ctx, mfg_dtype dt, const mfg_mesh *mesh);
void    ch_set(mfg_ch *ch, mfg_vec *v, double val);
void    ch_save(mfg_ch *ch, mfg_vec *v);
void    ch_save2(mfg_ch *ch, const mfg_vec *v1, mfg_vec *v2);
void    ch_load(mfg_ch *ch, mfg_vec *v);
void    ch_load_and_add(mfg_ch *ch, mfg_vec *v1, mfg_vec *v2);

mfg_laplace *laplace_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter);
mfg_laplace *laplace_from_arrays(mfg_ctx *ctx, mfg_mf *mf, mfg_ch *ch, const double *coef_host);
void         laplace_set_coefficient_host(mfg_laplace *op, const double *coef_host);
void         laplace_vmult(mfg_laplace *op, void *dst, const void *src, bool add, int part = -1, void *cuda_stream = nullptr);
uint32_t     laplace_set_interface_dofs(mfg_laplace *op, const uint32_t *dofs_host, size_t n);
void         mf_get_gpu_data(mfg_mf *mf, mfg_gpu_data *out);
void         laplace_compute_diagonal(mfg_laplace *op);
int          laplace_launches_per_vmult(const mfg_laplace *op);
int          laplace_active_variant(const mfg_laplace *op);
void         laplace_kernel_time(mfg_laplace *op, double *total_ms, int *n_launches);
bool         laplace_cell_dot(mfg_laplace *op, void *dst, const void *src, double *dot_out, uint32_t *n_dot);
}  // namespace mfg
