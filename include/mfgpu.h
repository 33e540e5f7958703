/*
 * mfgpu.h -- C ABI of the B200-native matrix-free engine (libmfgpu.so).
 *
 * The reference (kalj/dealii-cuda) has no FFI boundary: its operator API is a
 * set of C++ templates on top of deal.II.  This header is the boundary a
 * maintainer binds instead; every entry point names the reference interface it
 * replaces (file:line relative to the reference tree).  The header-only C++
 * facade in include/dealii_cuda_b200/ re-creates the reference's class names
 * (GpuVector, MatrixFreeGpu, ConstraintHandlerGpu, LaplaceOperatorGpu, ...) on
 * top of these calls.
 *
 * Conventions: opaque handles; every call returns 0 on success or a negative
 * mfg_status; mfg_last_error() returns a thread-local message (the reference
 * throws dealii::ExcMessage from CUDA_CHECK_SUCCESS, cuda_utils.cuh:15-23).
 * All work is enqueued on the context's CUDA stream; only calls documented as
 * "blocking" synchronise.  There is NO CPU fallback: if no CUDA device is
 * usable, mfg_ctx_create fails.
 */
#ifndef MFGPU_H
#define MFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mfg_status
{
  MFG_OK             = 0,
  MFG_ERR_INVALID    = -1, /* bad argument / unsupported configuration */
  MFG_ERR_CUDA       = -2, /* CUDA runtime error (message has file:line) */
  MFG_ERR_NOMEM      = -3,
  MFG_ERR_UNSUPPORTED = -4
} mfg_status;

typedef enum mfg_dtype { MFG_F32 = 0, MFG_F64 = 1 } mfg_dtype;

/* scatter strategy of the cell loop (reference: -DMATRIX_FREE_COLOR vs atomics,
 * laplace_operator_gpu.h:132-136, fee_gpu.cuh:358-361) */
typedef enum mfg_scatter { MFG_SCATTER_ATOMIC = 0, MFG_SCATTER_COLOR = 1 } mfg_scatter;

typedef struct mfg_ctx     mfg_ctx;
typedef struct mfg_vec     mfg_vec;     /* GpuVector<Number>            gpu_vec.h:22-176 */
typedef struct mfg_mesh    mfg_mesh;    /* Triangulation+DoFHandler+ConstraintMatrix substitute */
typedef struct mfg_mf      mfg_mf;      /* MatrixFreeGpu<dim,Number>    matrix_free_gpu.h:81-229 */
typedef struct mfg_ch      mfg_ch;      /* ConstraintHandlerGpu<Number> constraint_handler_gpu.h:13-59 */
typedef struct mfg_laplace mfg_laplace; /* LaplaceOperatorGpu<dim,p,Number> laplace_operator_gpu.h:35-96 */

const char *mfg_last_error(void);
const char *mfg_version(void);

/* ---- context ------------------------------------------------------------ */
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL
 * for the legacy default stream (what the reference uses everywhere). */
int mfg_ctx_create(int device, void *stream, mfg_ctx **out);
int mfg_ctx_destroy(mfg_ctx *ctx);
int mfg_ctx_set_stream(mfg_ctx *ctx, void *stream);
int mfg_ctx_synchronize(mfg_ctx *ctx); /* blocking; timer() in timing.cu:10-17 */
int mfg_ctx_device_info(mfg_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes);

/* ---- GpuVector ---------------------------------------------------------- */
/* gpu_vec.h / gpu_vec.cu.  n elements of dtype; create zero-fills like
 * GpuVector(unsigned) (gpu_vec.cu:33-37); resize does not (gpu_vec.cu:185-196). */
int mfg_vec_create(mfg_ctx *ctx, mfg_dtype dt, size_t n, mfg_vec **out);
int mfg_vec_wrap(mfg_ctx *ctx, mfg_dtype dt, size_t n, void *device_ptr, mfg_vec **out); /* non-owning view */
int mfg_vec_destroy(mfg_vec *v);
int mfg_vec_resize(mfg_vec *v, size_t n);
size_t mfg_vec_size(const mfg_vec *v);
mfg_dtype mfg_vec_dtype(const mfg_vec *v);
void *mfg_vec_data(mfg_vec *v);                                   /* getData()/getDataRO() */
int mfg_vec_from_host(mfg_vec *v, const void *host, size_t n);    /* fromHost, blocking */
int mfg_vec_to_host(const mfg_vec *v, void *host, size_t n);      /* copyToHost, blocking */
int mfg_vec_copy(mfg_vec *dst, const mfg_vec *src);               /* operator=(GpuVector), converts f32<->f64 */
int mfg_vec_swap(mfg_vec *a, mfg_vec *b);                         /* swap, gpu_vec.h:164-172 */
int mfg_vec_fill(mfg_vec *v, double a);                           /* operator=(Number)  gpu_vec.cu:373-381 */
int mfg_vec_sadd(mfg_vec *v, double s, double a, const mfg_vec *x);  /* v = s*v + a*x   gpu_vec.cu:306-313 */
int mfg_vec_equ(mfg_vec *v, double a, const mfg_vec *x);          /* v = a*x            gpu_vec.cu:345-351 */
int mfg_vec_scale(mfg_vec *v, const mfg_vec *x);                  /* v *= x (pointwise) gpu_vec.cu:317-322 */
int mfg_vec_divide(mfg_vec *v, const mfg_vec *x);                 /* v /= x (pointwise) gpu_vec.cu:325-331 */
int mfg_vec_invert(mfg_vec *v);                                   /* v = 1/v            gpu_vec.cu:333-340 */
int mfg_vec_scal(mfg_vec *v, double a);                           /* v *= a             gpu_vec.cu:356-362 */
int mfg_vec_dot(const mfg_vec *a, const mfg_vec *b, double *out); /* operator*, blocking gpu_vec.cu:541-558 */
int mfg_vec_add_and_dot(mfg_vec *v, double a, const mfg_vec *x, const mfg_vec *w, double *out); /* v+=a*x; return v.w  gpu_vec.cu:567-617 */
int mfg_vec_l2_norm(const mfg_vec *v, double *out);               /* gpu_vec.cu:366-369 */
int mfg_vec_all_zero(const mfg_vec *v, int *out);                 /* gpu_vec.cu:511-527 */
/* dst[dst_idx[i]] = src[src_idx[i]]  (device index arrays) gpu_vec.cu:624-644 */
int mfg_vec_copy_with_indices(mfg_vec *dst, const mfg_vec *src, const uint32_t *dst_idx, const uint32_t *src_idx, size_t n);

/* ---- mesh / DoF substrate ------------------------------------------------ */
/* What the reference obtains from deal.II (GridGenerator::hyper_cube +
 * refine_global, DoFHandler::distribute_dofs, interpolate_boundary_values;
 * bmop.cu:111-132, poisson_common.h:58-72, bmop_common.h:108-120), built on
 * the device.  Cells are ordered along the Morton curve (deal.II's order after
 * global refinement), DoFs are numbered first-touch in hierarchic order
 * (vertices, lines, quads, hex) exactly like DoFHandler::distribute_dofs. */
typedef struct mfg_box_desc
{
  int      dim;            /* 2 or 3 */
  int      degree;         /* FE_Q degree 1..8 */
  int      log2_cells[3];  /* cells per direction = 2^log2_cells[d] */
  double   origin[3];      /* lower corner */
  double   h;              /* cell edge length */
  uint32_t dirichlet_faces;/* bit f set: face f (x0,x1,y0,y1,z0,z1) carries homogeneous Dirichlet data */
} mfg_box_desc;

int mfg_mesh_create_box(mfg_ctx *ctx, const mfg_box_desc *desc, mfg_mesh **out);
/* hyper_cube(left,right)^dim refined n_refine times, Dirichlet on the whole boundary */
int mfg_mesh_hyper_cube(mfg_ctx *ctx, int dim, int degree, int n_refine, double left, double right, mfg_mesh **out);
int mfg_mesh_destroy(mfg_mesh *m);
uint32_t mfg_mesh_n_cells(const mfg_mesh *m);
uint32_t mfg_mesh_n_dofs(const mfg_mesh *m);
uint32_t mfg_mesh_dofs_per_cell(const mfg_mesh *m);
uint32_t mfg_mesh_n_constrained(const mfg_mesh *m);
/* device pointers (valid until mfg_mesh_destroy) */
const uint32_t *mfg_mesh_loc2glob_device(const mfg_mesh *m);     /* [n_cells][(p+1)^dim] lexicographic */
const uint32_t *mfg_mesh_constrained_device(const mfg_mesh *m);  /* ascending */
/* blocking copies to host */
int mfg_mesh_get_loc2glob(const mfg_mesh *m, uint32_t *host);
int mfg_mesh_get_constrained(const mfg_mesh *m, uint32_t *host);
int mfg_mesh_get_cell_coords(const mfg_mesh *m, uint32_t *host /* [n_cells][3] */);
/* support point of every DoF (DoFTools::map_dofs_to_support_points; VectorTools::interpolate_boundary_values needs them,
   poisson.cu:155-158): host [n_dofs][dim] */
int mfg_mesh_get_support_points(const mfg_mesh *m, double *host);
/* global DoF index of lattice points (x,y,z in 0..p*N_d), host arrays, blocking */
int mfg_mesh_lattice_to_dof(const mfg_mesh *m, size_t n, const uint32_t *lattice_xyz, uint32_t *dof);
/* graph coloring of the cells (coloring.cc:20-33): color_of_cell[n_cells] to host; returns n_colors in *n_colors */
int mfg_mesh_color_cells(const mfg_mesh *m, uint32_t *color_of_cell, uint32_t *n_colors);
/* GraphColoringWrapper::make_graph_coloring (matrix_free_gpu/coloring.cc:8-33) for any mesh, on the host: deal.II's
 * algorithm restated (zones by breadth-first search over shared conflict indices, DSATUR inside a zone, colors of the even
 * and of the odd zones merged).  conflict_indices_host: [n_cells][dofs_per_cell] = the cell DoFs after resolve_indices
 * (bit 31 ignored).  Feed the result to mfg_mf_desc.n_colors / color_offsets (cells sorted by color). */
int mfg_graph_coloring(uint32_t n_cells, uint32_t dofs_per_cell, const uint32_t *conflict_indices_host, uint32_t n_indices,
                       uint32_t *color_of_cell, uint32_t *n_colors);

/* ---- adaptively refined meshes with hanging nodes: host substrate ------------------------------------------------------
 * What the reference takes from deal.II on such meshes (bmop.cu / poisson.cu with an adaptive grid), restated on the host:
 * a Triangulation on hyper_cube(left, right) with refine_global / set_refine_flag / execute_coarsening_and_refinement
 * (refinement flags are closed under deal.II's one-level rule across faces and, in 3D, edges; cells are kept per level in
 * creation order, so the active cells iterate like deal.II's), the reference's flagging helpers (bmop_common.h:9-105,
 * poisson_common.h:29-35), DoFHandler::distribute_dofs for FE_Q(p) and HangingNodes::setup_constraints
 * (matrix_free_gpu/hanging_nodes.cuh:209-454): the 9-bit mask per cell and loc2glob with the coarse neighbour's DoFs on
 * constrained faces / edges.  Pure host code, no device needed; mfg_laplace_create_from_amesh builds the device objects
 * through mfg_mf_reinit (constraint_mask) + mfg_ch_create + mfg_laplace_create_from_arrays. */
typedef struct mfg_amesh mfg_amesh;
int mfg_amesh_create(int dim, int degree, double left, double right, mfg_amesh **out);     /* GridGenerator::hyper_cube, one cell */
int mfg_amesh_destroy(mfg_amesh *am);
/* Triangulation::limit_level_difference_at_vertices (the MeshSmoothing flag of the reference's multigrid drivers, poisson_mg.cu:132,
 * bmop_mg.cu:130): cells that share only a vertex also differ by at most one level; required by the multigrid hierarchy */
int mfg_amesh_set_limit_level_difference_at_vertices(mfg_amesh *am, int on);
int mfg_amesh_refine_global(mfg_amesh *am, int times);                                     /* Triangulation::refine_global */
int mfg_amesh_set_refine_flags(mfg_amesh *am, const uint8_t *flags, size_t n_flags /* = n_active_cells */); /* cell->set_refine_flag() */
int mfg_amesh_mark_cells_in_annulus(mfg_amesh *am, double R, double r, const double *center /* [dim] or NULL = origin */); /* bmop_common.h:9-24 */
int mfg_amesh_mark_cells_on_shell(mfg_amesh *am, double R, const double *center);          /* bmop_common.h:27-47 */
int mfg_amesh_mark_octant(mfg_amesh *am);                                                  /* mark_cells(octant_criterion) poisson_common.h:29-35, 43-56 */
int mfg_amesh_execute_refinement(mfg_amesh *am);                                           /* execute_coarsening_and_refinement */
int mfg_amesh_pseudo_adaptive_refinement(mfg_amesh *am, int n_ref);                        /* bmop_common.h:49-105, domain CUBE */
int mfg_amesh_info(const mfg_amesh *am, int *dim, int *degree, double *left, double *right, int *n_levels, int *coarsest_active_level); /* any pointer may be NULL */
uint32_t mfg_amesh_n_active_cells(const mfg_amesh *am);
uint32_t mfg_amesh_n_levels(const mfg_amesh *am);
int mfg_amesh_get_active_cells(const mfg_amesh *am, uint32_t *level_xyz /* [n_active_cells][4]: level, x, y, z */);
/* all cells of a level (active or refined) in storage order = the level mesh of the multigrid hierarchy */
uint32_t mfg_amesh_n_level_cells(const mfg_amesh *am, int level);
int mfg_amesh_get_level_cells(const mfg_amesh *am, int level, uint32_t *xyz_children /* [n][4]: x, y, z, 1 if the cell has children */);
/* DoFHandler::distribute_dofs + HangingNodes::setup_constraints + the ConstraintHandlerGpu list (hanging and boundary DoFs) */
int mfg_amesh_distribute_dofs(mfg_amesh *am);
uint32_t mfg_amesh_n_dofs(const mfg_amesh *am);
uint32_t mfg_amesh_n_constrained(const mfg_amesh *am);
uint32_t mfg_amesh_n_hanging(const mfg_amesh *am);
uint32_t mfg_amesh_n_boundary(const mfg_amesh *am);
int mfg_amesh_get_boundary(const mfg_amesh *am, uint32_t *out);          /* DoFs on the domain boundary, ascending */
int mfg_amesh_get_support_points(const mfg_amesh *am, double *out);      /* DoFTools::map_dofs_to_support_points: [n_dofs][dim] */
/* any pointer may be NULL.  loc2glob: [n_cells][(p+1)^dim] after the rewrite; loc2glob_unconstrained: the DoFHandler's own map;
 * constraint_mask [n_cells]; constrained [n_constrained] ascending; hanging [n_hanging]; inv_jac [n_cells];
 * coefficient [n_cells][(p+1)^dim] = 1/(0.05+2|x_q|^2) (poisson_common.h:146-158); quadrature_points [n_cells][(p+1)^dim][dim] */
int mfg_amesh_get_arrays(const mfg_amesh *am, uint32_t *loc2glob, uint32_t *loc2glob_unconstrained, uint32_t *constraint_mask, uint32_t *constrained,
                         uint32_t *hanging, double *inv_jac, double *coefficient, double *quadrature_points);

/* The multigrid hierarchy on an adaptively refined mesh (local smoothing, poisson_mg.cu / bmop_mg.cu with an adaptive grid; the mesh
 * needs mfg_amesh_set_limit_level_difference_at_vertices like the reference's Triangulation): level meshes = all cells of a level,
 * DoFHandler::distribute_mg_dofs, MGConstrainedDoFs (boundary and refinement-edge indices), the transfer blocks of
 * MGTransferMatrixFreeGpu::build (mg_transfer_matrix_free_gpu.cu:173-257) and the copy indices of copy_to_mg / copy_from_mg
 * (.cu:109-146, 688-757).  Host code; min_level <= the coarsest active level. */
int mfg_amesh_build_mg(mfg_amesh *am, int min_level);
/* out[0..5] = cells, DoFs, boundary indices, refinement-edge indices, transfer blocks into this level (= refined cells of level-1),
 * copy-index pairs */
int mfg_amesh_mg_level_sizes(const mfg_amesh *am, int level, uint32_t out[6]);
/* any pointer may be NULL.  loc2glob [cells][(p+1)^dim] (level numbering, lexicographic); boundary / edge ascending; coefficient
 * [cells][(p+1)^dim]; copy_global / copy_level [pairs]: active DoF <-> level DoF; coarse_idx [blocks][(p+1)^dim] level-1 DoFs of the
 * refined cell (bit 31 = boundary DoF of level-1), fine_idx [blocks][(2p+1)^dim] level DoFs of its children (lexicographic lattice),
 * weights [blocks][3^dim] = 1 / multiplicity of the fine DoFs of a block region (low face, interior, high face per direction) */
int mfg_amesh_mg_level_get(const mfg_amesh *am, int level, uint32_t *loc2glob, uint32_t *boundary, uint32_t *edge, double *coefficient, uint32_t *copy_global,
                           uint32_t *copy_level, uint32_t *coarse_idx, uint32_t *fine_idx, double *weights);

/* ---- the reference's BALL_GRID (poisson_common.h:59-72: GridGenerator::hyper_ball + SphericalManifold on the boundary + refine_global)
 * as an unstructured quad / hex mesh with FE_Q DoFs and the tri-linear (MappingQ1) geometry per quadrature point (host code,
 * csrc/ball_mesh.cu).  The operator on it runs on the general-geometry path (MFG_GEOM_GENERAL). */
typedef struct mfg_umesh mfg_umesh;
int mfg_umesh_hyper_ball(int dim, int degree, double radius, mfg_umesh **out);
int mfg_umesh_destroy(mfg_umesh *um);
int mfg_umesh_refine_global(mfg_umesh *um, int times);
int mfg_umesh_distribute_dofs(mfg_umesh *um);               /* DoFHandler::distribute_dofs + the Dirichlet boundary DoFs */
uint32_t mfg_umesh_n_cells(const mfg_umesh *um);
uint32_t mfg_umesh_n_vertices(const mfg_umesh *um);
uint32_t mfg_umesh_n_dofs(const mfg_umesh *um);
uint32_t mfg_umesh_n_boundary(const mfg_umesh *um);
int mfg_umesh_get_support_points(const mfg_umesh *um, double *out);   /* DoFTools::map_dofs_to_support_points (MappingQ1): [n_dofs][dim] */
int mfg_umesh_get_mesh(const mfg_umesh *um, double *vertices /* [n_vertices][dim] */, uint32_t *cell_vertices /* [n_cells][2^dim], lexicographic */);
/* any pointer may be NULL.  loc2glob [n_cells][(p+1)^dim]; boundary [n_boundary] ascending; inv_jac [n_cells][(p+1)^dim][dim][dim]
 * (K[d1][d2] = d xi_d1 / d x_d2, FEValues::get_inverse_jacobians order); JxW, coefficient [n_cells][(p+1)^dim];
 * quadrature_points [n_cells][(p+1)^dim][dim] */
int mfg_umesh_get_arrays(const mfg_umesh *um, uint32_t *loc2glob, uint32_t *boundary, double *inv_jac, double *JxW, double *quadrature_points,
                         double *coefficient);

/* ---- MatrixFreeGpu ------------------------------------------------------- */
/* Explicit-array description: what ReinitHelper extracts from deal.II
 * (matrix_free_gpu.cu:283-339) -- this is the call a deal.II-based caller makes. */
typedef enum mfg_geometry
{
  MFG_GEOM_UNIFORM = 0, /* -DMATRIX_FREE_UNIFORM_MESH: scalar J^-1[0][0] per cell (matrix_free_gpu.cu:332-334) */
  MFG_GEOM_GENERAL = 1  /* full J^-1 per quadrature point (fee_gpu.cuh:236-240, 276-280; matrix_free_gpu.cu:326-338) */
} mfg_geometry;

typedef struct mfg_mf_desc
{
  int             dim, degree;
  mfg_dtype       dtype;
  uint32_t        n_cells, n_dofs;
  const uint32_t *loc2glob;        /* host, [n_cells][(p+1)^dim], lexicographic, unpadded */
  mfg_geometry    geometry;
  const double   *inv_jac;         /* host, UNIFORM: [n_cells]; GENERAL: [n_cells][(p+1)^dim][dim][dim], K[d1][d2] = d xi_d1 / d x_d2
                                      (FEValues::get_inverse_jacobians order; JxW is then required) */
  const double   *JxW;             /* host, [n_cells][(p+1)^dim] or NULL: then JxW = inv_jac^-dim * w_q */
  const double   *quadrature_points; /* host, [n_cells][(p+1)^dim][dim] (for evaluate_on_cells) or NULL */
  mfg_scatter     scatter;
  uint32_t        n_colors;        /* COLOR: cells must be sorted by color */
  const uint32_t *color_offsets;   /* COLOR: [n_colors+1] */
  /* hanging nodes (-DMATRIX_FREE_HANGING_NODES): per-cell 9-bit mask of HangingNodes::setup_constraints
   * (hanging_nodes.cuh:38-50, 209-454); loc2glob must already carry the coarse neighbour's DoFs on the
   * constrained faces / edges.  NULL: no hanging nodes.  Atomic scatter only. */
  const uint32_t *constraint_mask; /* host, [n_cells] or NULL */
} mfg_mf_desc;

int mfg_mf_reinit(mfg_ctx *ctx, const mfg_mf_desc *desc, mfg_mf **out);              /* MatrixFreeGpu::reinit matrix_free_gpu.cu:448-563 */
int mfg_mf_reinit_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter, mfg_mf **out);
/* the same on an adaptively refined mesh of the library (masks, rewritten loc2glob, quadrature points), for user-written cell loops */
int mfg_mf_reinit_from_amesh(mfg_ctx *ctx, const mfg_amesh *am, mfg_dtype dt, mfg_mf **out);
int mfg_mf_reinit_from_umesh(mfg_ctx *ctx, const mfg_umesh *um, mfg_dtype dt, mfg_mf **out);   /* the ball mesh: general geometry */
int mfg_mf_destroy(mfg_mf *mf);                                                       /* MatrixFreeGpu::free matrix_free_gpu.cu:566-596 */
uint32_t mfg_mf_n_dofs(const mfg_mf *mf);
uint32_t mfg_mf_n_cells(const mfg_mf *mf);
uint32_t mfg_mf_n_colors(const mfg_mf *mf);
size_t mfg_mf_memory_consumption(const mfg_mf *mf);                                   /* matrix_free_gpu.h:437-459 */
/* MatrixFreeGpu::get_gpu_data (matrix_free_gpu.h:261-278): the device arrays a user-written cell functor needs, for the
 * header-only generic path include/dealii_cuda_b200/fee_gpu.cuh (FEEvaluationGpu + cell_loop compiled with the user's
 * LocalOperator).  Arrays are in the cell order of the kernels (cells sorted by color; cells with hanging nodes last) and
 * stay valid until mfg_mf_destroy.  The first call builds JxW / inv_jac / quadrature points on the device. */
typedef struct mfg_gpu_data
{
  const uint32_t *loc2glob;          /* device [n_cells][(p+1)^dim], lexicographic; bit 31 may flag a constrained DoF
                                        (set once a LaplaceOperatorGpu was built on this mf): mask with 0x7fffffff */
  const void     *JxW;               /* device Number [n_cells][(p+1)^dim] */
  const void     *inv_jac;           /* device Number: uniform [n_cells]; general [n_cells][(p+1)^dim][dim][dim] */
  const void     *quadrature_points; /* device Number [n_cells][(p+1)^dim][dim], NULL if the description had none */
  const uint32_t *constraint_mask;   /* device [n_cells] hanging-node masks or NULL */
  const uint32_t *color_offsets;     /* host [n_colors+1] */
  uint32_t        n_cells, n_dofs, n_colors, n_plain_cells;
  int             dim, degree, general, use_coloring;
  mfg_dtype       dtype;
  void           *cuda_stream;       /* the context's stream */
  double          shape_values[81], shape_gradients[81], colloc_gradients[81]; /* [i*n+q], n = degree+1 */
} mfg_gpu_data;
int mfg_mf_get_gpu_data(mfg_mf *mf, mfg_gpu_data *out);
/* shape tables handed to the kernels: [i*n+q] = phi_i(x_q), phi_i'(x_q) (matrix_free_gpu.cu:502-513) */
int mfg_shape_info(int degree, double *shape_values, double *shape_gradients, double *q_points, double *q_weights);
/* 1-D hanging-node interpolation weights W[k*n+i] = phi_i(xi_k/2) (setup_constraint_weights, hanging_nodes.cuh:580-598) */
int mfg_hanging_node_weights(int degree, double *weights);

/* ---- ConstraintHandlerGpu ------------------------------------------------ */
int mfg_ch_create(mfg_ctx *ctx, mfg_dtype dt, const uint32_t *constrained_host, size_t n_constrained,
                  const uint32_t *edge_host, size_t n_edge, mfg_ch **out);          /* reinit constraint_handler_gpu.cu:69-123 */
int mfg_ch_create_from_mesh(mfg_ctx *ctx, mfg_dtype dt, const mfg_mesh *mesh, mfg_ch **out);
int mfg_ch_destroy(mfg_ch *ch);
size_t mfg_ch_n_constrained(const mfg_ch *ch);
int mfg_ch_set_constrained_values(mfg_ch *ch, mfg_vec *v, double val);               /* :127-137 */
int mfg_ch_save_constrained_values(mfg_ch *ch, mfg_vec *v);                          /* :140-150 */
int mfg_ch_save_constrained_values2(mfg_ch *ch, const mfg_vec *v1, mfg_vec *v2);     /* :153-166 */
int mfg_ch_load_constrained_values(mfg_ch *ch, mfg_vec *v);                          /* :168-178 */
int mfg_ch_load_and_add_constrained_values(mfg_ch *ch, mfg_vec *v1, mfg_vec *v2);    /* :181-194 */
int mfg_ch_copy_edge_values(mfg_ch *ch, mfg_vec *dst, const mfg_vec *src);           /* :196-200 */

/* ---- LaplaceOperatorGpu --------------------------------------------------- */
/* reinit(dof_handler, constraints) laplace_operator_gpu.h:120-151.  The
 * coefficient 1/(0.05+2|x|^2) (poisson_common.h:146-158) is evaluated on the
 * device at the Gauss points (evaluate_coefficient, :204-211).  */
int mfg_laplace_create(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter, mfg_laplace **out);
/* explicit arrays: mf (from mfg_mf_reinit), constraint handler, and coefficient
 * values at quadrature points [n_cells][(p+1)^dim] (host, double). The operator
 * takes ownership of neither mf nor ch. */
int mfg_laplace_create_from_arrays(mfg_ctx *ctx, mfg_mf *mf, mfg_ch *ch, const double *coefficient_host, mfg_laplace **out);
/* LaplaceOperatorGpu::reinit on an adaptively refined mesh (-DMATRIX_FREE_HANGING_NODES): MatrixFreeGpu with the masks,
 * ConstraintHandlerGpu with hanging + boundary DoFs, the reference coefficient; the operator owns both. */
int mfg_laplace_create_from_amesh(mfg_ctx *ctx, const mfg_amesh *am, mfg_dtype dt, mfg_laplace **out);
/* LaplaceOperatorGpu::reinit on the ball mesh (-DBALL_GRID): general geometry, Dirichlet boundary, the reference coefficient */
int mfg_laplace_create_from_umesh(mfg_ctx *ctx, const mfg_umesh *um, mfg_dtype dt, mfg_laplace **out);
/* replace the coefficient: a(x_q) at quadrature points, host, [n_cells][(p+1)^dim], cells in mesh / descriptor order */
int mfg_laplace_set_coefficient(mfg_laplace *op, const double *coefficient_host);
int mfg_laplace_destroy(mfg_laplace *op);                                            /* clear() :110-117 */
uint32_t mfg_laplace_m(const mfg_laplace *op);                                       /* m()/n() :52-53 */
/* kernel variant: 0 = auto, otherwise a variant id; for A/B measurements */
int mfg_laplace_set_variant(mfg_laplace *op, int variant);
/* measurement switches (A/B runs, tests): "cg_fused" 0/1 (mfg_solver_cg: d.(A d) from the cell kernel, default 1),
 * "stage_sync" 0/1 and "stage_merge_dirs" 0..7 (staged kernel, variant 40) */
int mfg_laplace_set_option(mfg_laplace *op, const char *name, int value);
int mfg_laplace_vmult(mfg_laplace *op, mfg_vec *dst, const mfg_vec *src);            /* :216-223 (Tvmult identical) */
int mfg_laplace_vmult_add(mfg_laplace *op, mfg_vec *dst, const mfg_vec *src);        /* :286-303 */
/* raw device pointers (dtype of the operator), for callers that own their memory */
int mfg_laplace_vmult_ptr(mfg_laplace *op, void *dst_dev, const void *src_dev);
int mfg_laplace_vmult_add_ptr(mfg_laplace *op, void *dst_dev, const void *src_dev);
/* host buffers: H2D copy of src, vmult, D2H copy of dst; blocking */
int mfg_laplace_vmult_host(mfg_laplace *op, void *dst_host, const void *src_host);
/* pipelined host-buffer apply: two slots (0/1), each with its own device staging; H2D, vmult and D2H of
 * consecutive calls overlap on three streams (PCIe is full duplex).  Host buffers should be pinned and must stay
 * valid until mfg_laplace_host_sync.  Non-blocking. */
int mfg_laplace_vmult_host_async(mfg_laplace *op, void *dst_host, const void *src_host, int slot);
int mfg_laplace_host_sync(mfg_laplace *op);   /* blocking: all pipelined applies have landed in their dst_host */
int mfg_laplace_compute_diagonal(mfg_laplace *op);                                   /* :405-421 */
int mfg_laplace_get_diagonal_inverse(mfg_laplace *op, mfg_vec **out);                /* :423-429, borrowed */
size_t mfg_laplace_memory_consumption(const mfg_laplace *op);                        /* :434-445 */
/* number of kernel launches one vmult enqueues (for bench.py's gpu_launches) */
int mfg_laplace_launches_per_vmult(const mfg_laplace *op);
int mfg_laplace_cell_launches_per_vmult(const mfg_laplace *op); /* cell-kernel launches among them */
/* per-kernel device timing for the roofline figure: on = k > 0 brackets every k-th cell-kernel launch by CUDA events on
 * the context stream (k = 1: every launch; a bracketed launch cannot overlap its neighbours, so a sample keeps the
 * measured loop honest); kernel_time_ms (blocking) returns and resets the accumulated time and the number of
 * bracketed launches. */
int mfg_laplace_enable_kernel_timing(mfg_laplace *op, int on);
int mfg_laplace_kernel_time_ms(mfg_laplace *op, double *total_ms, int *n_launches);
int mfg_laplace_active_variant(const mfg_laplace *op);   /* the kernel that will run (1 / 40 / 50); -1 if the requested variant cannot run on this operator (mfg_last_error) */
/* staged cell kernel (variant 40): numbers of its plan after the first apply: out[0..7] = groups, staged groups, patterns,
 * and per staged group x 16: own DoFs, halo entries, plain-stored DoFs, red.add DoFs, shared-memory wavefronts of the
 * staged reads + writes */
int mfg_laplace_stage_stats(const mfg_laplace *op, uint32_t out[8]);
/* The plan of the staged kernel for a host index array (pure host code, no device needed): what the kernel's copy lists,
 * position tables and merge masks are for [n_cells][(degree+1)^3] lexicographic indices with bit 31 = constrained DoF.
 * Used by the CPU tests, which replay the kernel's data movement with numpy and compare with a plain gather / scatter.
 * info[0..14] = groups, patterns, pattern stride (uint16), halo entries, fallback groups, cells per group, cells per
 * half warp, staging slots, points per direction, staged groups, load-list capacity, own-range capacity, and per staged
 * group the shared-memory wavefronts of the slab reads, of the staged writes and of one pass over the load list */
typedef struct mfg_stage_plan mfg_stage_plan;
int mfg_stage_plan_build(int degree, mfg_dtype dt, uint32_t n_plain, uint32_t n_cells, uint32_t n_dofs, const uint32_t *idx_host,
                         int merge_dirs, mfg_stage_plan **out);
int mfg_stage_plan_info(const mfg_stage_plan *p, uint32_t info[16]);
int mfg_stage_plan_get(const mfg_stage_plan *p, uint32_t *gdesc, uint32_t *halo, uint16_t *ptab, uint32_t *fallback);
int mfg_stage_plan_destroy(mfg_stage_plan *p);
/* bmop loop (bmop.cu:135-153): dst=init; k times {swap; vmult}; result left in *dst.
 * Returns device milliseconds measured with CUDA events on the context stream. */
int mfg_laplace_bmop(mfg_laplace *op, mfg_vec *dst, mfg_vec *src, int k, double init, float *elapsed_ms);

/* ---- MGTransferMatrixFreeGpu (matrix_free_gpu/mg_transfer_matrix_free_gpu.h:64-307) ----------------------------
 * One object per level pair of a globally refined mesh hierarchy (fine = global refinement of coarse).
 * prolongate: dst_fine = P src_coarse, coarse Dirichlet DoFs read as 0 (.cu:592-622).
 * restrict_and_add: dst_coarse += P^T src_fine on non-Dirichlet coarse DoFs (.cu:626-654).
 * copy_to_mg / copy_from_mg (.cu:688-757) are plain copies on globally refined meshes (the finest level IS the
 * active mesh with the same numbering): mfg_vec_copy. */
typedef struct mfg_mgt mfg_mgt;
int mfg_mgt_build(mfg_ctx *ctx, const mfg_mesh *coarse, const mfg_mesh *fine, mfg_dtype dt, mfg_mgt **out);
int mfg_mgt_destroy(mfg_mgt *t);
int mfg_mgt_prolongate(mfg_mgt *t, mfg_vec *dst_fine, const mfg_vec *src_coarse);
int mfg_mgt_restrict_and_add(mfg_mgt *t, mfg_vec *dst_coarse, const mfg_vec *src_fine);

/* transfer between two levels of an ADAPTIVE hierarchy from explicit blocks (mfg_amesh_mg_level_get: coarse_idx, fine_idx, weights):
 * one block per refined cell of the coarse level (internal::MGTransfer::setup_transfer, mg_transfer_matrix_free_gpu.cu:173-257) */
int mfg_mgt_build_from_blocks(mfg_ctx *ctx, mfg_dtype dt, int dim, int degree, uint32_t n_blocks, const uint32_t *coarse_idx_host, const uint32_t *fine_idx_host,
                              const double *weights_host, uint32_t n_coarse_dofs, uint32_t n_fine_dofs, mfg_mgt **out);

/* ---- assembled sparse matrix: the competitor row of the reference (CUDAWrappers::SparseMatrix, matrix_free_gpu/cuda_sparse_matrix.{h,cu};
 * bmop_spm.cu, poisson_spm.cu, test_spm.cu) ------------------------------------------------------------------------------
 * mfg_csr_assemble_laplace: host assembly (bmop_spm.cu:150-201) of the same operator the matrix-free path applies, from the same
 * arrays (uniform geometry, no hanging nodes): constrained rows / columns eliminated, diagonal 1.  mfg_spm_*: the device matrix
 * (reinit = upload, cuda_sparse_matrix.cu:60-120) and vmult (:414-429; here a hand-written CSR kernel, one warp per row). */
typedef struct mfg_csr mfg_csr;
typedef struct mfg_spm mfg_spm;
int mfg_csr_assemble_laplace(int dim, int degree, uint32_t n_cells, uint32_t n_dofs, const uint32_t *loc2glob_host, const double *inv_jac_host,
                             const double *coefficient_host, const uint32_t *constrained_host, size_t n_constrained, mfg_csr **out);
int mfg_csr_destroy(mfg_csr *A);
int mfg_csr_sizes(const mfg_csr *A, uint32_t *n_rows, size_t *nnz);
int mfg_csr_get(const mfg_csr *A, uint32_t *row_ptr, uint32_t *col, double *val);
int mfg_spm_create(mfg_ctx *ctx, mfg_dtype dt, const mfg_csr *A, mfg_spm **out);
int mfg_spm_create_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_spm **out);
int mfg_spm_destroy(mfg_spm *S);
uint32_t mfg_spm_m(const mfg_spm *S);
size_t mfg_spm_n_nonzero_elements(const mfg_spm *S);
size_t mfg_spm_memory_consumption(const mfg_spm *S);
int mfg_spm_vmult(mfg_spm *S, mfg_vec *dst, const mfg_vec *src);

/* ---- solver ----------------------------------------------------------------------------------------------
 * Conjugate gradients with the control flow of deal.II's SolverCG as instantiated on GpuVector by the reference
 * (poisson.cu:233-260): stops when |r| <= abs_tol (the reference uses 1e-12*|b|) or after max_iter iterations.
 * use_jacobi: precondition with the operator's inverse diagonal (PreconditionChebyshev's default degree 0 is a
 * scaled Jacobi step).  residual_history (optional) receives |r| after every iteration: max_iter+1 entries. */
int mfg_solver_cg(mfg_laplace *op, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int use_jacobi, int *iters,
                  double *last_residual, double *residual_history);

/* ---- PreconditionChebyshev / Multigrid / PreconditionMG as the reference instantiates them on GpuVector ----------------
 * (poisson_mg.cu:430-552, bmop_mg.cu:300-340; in the reference this is deal.II template code, here C++ host code of the
 * library: level operators, matrix-free transfer, one fused vector kernel per Chebyshev product).
 * mfg_chebyshev_create: PreconditionChebyshev::initialize with AdditionalData{degree, smoothing_range, eig_cg_n_iterations,
 * preconditioner = the operator's inverse diagonal}: lambda_max from eig_cg_n_iterations CG/Lanczos steps on Dinv A (deal.II's
 * procedure); eig_cg_n_iterations = 0 takes lambda_max = 1 (deal.II's default max_eigenvalue). */
typedef struct mfg_cheb mfg_cheb;
typedef struct mfg_mg mfg_mg;
int mfg_chebyshev_create(mfg_laplace *op, int degree, double smoothing_range, int eig_cg_n_iterations, mfg_cheb **out);
int mfg_chebyshev_destroy(mfg_cheb *c);
int mfg_chebyshev_vmult(mfg_cheb *c, mfg_vec *dst, const mfg_vec *src);   /* dst = p(Dinv A) Dinv src from a zero guess */
int mfg_chebyshev_step(mfg_cheb *c, mfg_vec *dst, const mfg_vec *src);    /* the same sweep starting from dst */
int mfg_chebyshev_info(const mfg_cheb *c, double *lambda_max, double *lambda_min, double *theta, double *delta, int *eig_iterations);
/* One Chebyshev product's vector update on its own (PreconditionChebyshev::vector_updates, precondition_chebyshev in deal.II):
 *   zero_start: d = f2 Dinv b, x = d;   else: d = (first ? 0 : f1 d) + f2 Dinv (b - ax), x += d      with ax = A x from the caller
 * (ax may be NULL with zero_start).  For callers whose operator product is not an mfg_laplace, e.g. the partitioned operator with
 * its interface exchange (dealii_cuda_b200/partitioned_mg.py). */
int mfg_vec_chebyshev_update(mfg_ctx *ctx, mfg_vec *x, mfg_vec *d, const mfg_vec *ax, const mfg_vec *b, const mfg_vec *dinv, double f1, double f2,
                             int zero_start, int first);
/* Geometric multigrid on hyper_cube(left,right) refine_global(min_level .. max_level): level LaplaceOperatorGpu (level_mg_handler,
 * laplace_operator_gpu.h:156-186; on a globally refined mesh a level IS a uniform mesh and its edge matrices are zero),
 * MGTransferMatrixFreeGpu, Chebyshev smoothers (poisson_mg.cu:461-470: degree 5, range 15, 15 eigenvalue iterations), coarse
 * solve = unpreconditioned CG to a 1e-10 reduction (poisson_mg.cu:73-80).  mfg_mg_vcycle = PreconditionMG::vmult;
 * mfg_mg_solve_cg = SolverCG with that preconditioner on the finest level (poisson_mg.cu:504-518). */
int mfg_mg_create(mfg_ctx *ctx, int dim, int degree, int min_level, int max_level, mfg_dtype dt, double left, double right, int smoother_degree,
                  double smoothing_range, int eig_cg_n_iterations, mfg_mg **out);
int mfg_mg_destroy(mfg_mg *mg);
int mfg_mg_vcycle(mfg_mg *mg, mfg_vec *dst, const mfg_vec *src);
int mfg_mg_level_operator(mfg_mg *mg, int level, mfg_laplace **op);   /* borrowed */
int mfg_mg_info(const mfg_mg *mg, int level, double *lambda_max, long *coarse_iterations, size_t *n_dofs);
int mfg_mg_solve_cg(mfg_mg *mg, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int *iters, double *last_residual, double *history);

/* ---- multi-GPU: interface-DoF exchange (new capability; the reference is single-GPU) -----------
 * The mesh is partitioned into boxes of cells, one per GPU.  Every rank stores all DoFs its cells touch;
 * DoFs on partition interfaces are replicated and kept consistent.  After the local cell loop the partial
 * sums of the interface DoFs are exchanged and added in ascending rank order on every replica
 * (bit-identical replicas).  The transport is the caller's (NCCL all_to_all / P2P copies); this object owns
 * the index lists and the pack / ordered-accumulate kernels.
 *   pack_idx[n_send]            local DoF of every send-buffer entry (send buffer = concatenation over
 *                               destination ranks in ascending rank order)
 *   shared_dofs[n_shared]       local DoFs that receive contributions
 *   offsets[n_shared+1], slots  CSR: contributions of shared DoF u in ascending rank order; slot >= 0 is an
 *                               index into the receive buffer, slot == -1 is this rank's own partial sum */
/* Box partition and exchange plan on the host (csrc/partition.cu; pure host code, no device needed).  Ranks sit on the grid
 * 1 -> 1x1x1, 2 -> 1x1x2, 4 -> 1x2x2, 8 -> 2x2x2 (2D: 2 -> 1x2, 4 -> 2x2), x fastest.  weak scaling (strong = 0): a 2^r cube of
 * cells per rank; strong = 1: the refine_global(r) cube [left,right]^dim cut into the rank grid. */
typedef struct mfg_partition_plan mfg_partition_plan;
int mfg_partition_rank_coords(int rank, int world, int dim, int coords[3], int grid[3]);
int mfg_partition_box(int rank, int world, int dim, int degree, int r, double left, double right, int strong, mfg_box_desc *out);
int mfg_partition_global_n_dofs(int world, int dim, int degree, int r, int strong, uint64_t *out);
/* lattice points (local coordinates 0 .. degree*2^lg_d, lexicographic, x fastest: the same order on both sides) this rank shares
 * with the neighbour at grid offset delta[d] in {-1,0,1}; drop_dirichlet leaves out points on the global boundary (constrained on
 * every replica, not exchanged).  xyz NULL: count only.  *neighbor_rank = -1 when there is no rank at that offset. */
int mfg_partition_interface_points(int rank, int world, int dim, int degree, int r, int strong, const int delta[3], int drop_dirichlet,
                                   int *neighbor_rank, size_t *n_points, uint32_t *xyz);
/* the plan of mfg_exchange_create from the DoF lists (mfg_mesh_lattice_to_dof of the points above): list i holds
 * dofs[list_start[i] .. list_start[i+1]) shared with rank list_rank[i]; the `repl` lists (all shared points incl. Dirichlet
 * ones) only decide ownership: owner = lowest rank that touches a DoF */
int mfg_partition_plan_create(int rank, int world, uint32_t n_local, int n_lists, const int *list_rank, const size_t *list_start, const uint32_t *dofs,
                              int n_repl, const int *repl_rank, const size_t *repl_start, const uint32_t *repl_dofs, mfg_partition_plan **out);
int mfg_partition_plan_destroy(mfg_partition_plan *p);
int mfg_partition_plan_sizes(const mfg_partition_plan *p, size_t out[5]); /* n_send, n_shared, n_slots, n_neighbors, n_local */
int mfg_partition_plan_get(const mfg_partition_plan *p, int *neighbors, uint32_t *splits /* [world] */, uint32_t *recv_off /* [world] */,
                           uint32_t *pack_idx, uint32_t *shared_dofs, uint32_t *offsets, int32_t *slots, uint8_t *owned_mask);
typedef struct mfg_exchange mfg_exchange;
int mfg_exchange_create(mfg_ctx *ctx, mfg_dtype dt, const uint32_t *pack_idx_host, size_t n_send, const uint32_t *shared_dofs_host,
                        size_t n_shared, const uint32_t *offsets_host, const int32_t *slots_host, size_t n_slots, mfg_exchange **out);
int mfg_exchange_destroy(mfg_exchange *ex);
int mfg_exchange_pack(mfg_exchange *ex, const void *vec_dev, void *send_dev);           /* send[k] = vec[pack_idx[k]] */
int mfg_exchange_accumulate(mfg_exchange *ex, void *vec_dev, const void *recv_dev);     /* ordered sum into vec */
/* the same on another CUDA stream (overlap with the interior cells of mfg_laplace_vmult_part_ptr) */
int mfg_exchange_accumulate_stream(mfg_exchange *ex, void *vec_dev, const void *recv_dev, void *cuda_stream);
/* Overlap of the interface exchange with interior cells: mark the DoFs whose partial sums are exchanged;
   part 0 of an apply = zero/constraint pass, part 1 = the cell groups contributing to marked DoFs, part 2 = all other
   groups (touches no marked DoF), part -1 = everything.  Parts 1 and 2 may be enqueued on different streams after
   part 0 (cuda_stream, NULL = the context's stream) and then run side by side.  Kernels without a work list run all
   cells in part 2. */
int mfg_laplace_set_interface_dofs(mfg_laplace *op, const uint32_t *dofs_host, size_t n, uint32_t *n_interface_groups);
int mfg_laplace_vmult_part_ptr(mfg_laplace *op, void *dst_dev, const void *src_dev, int part, void *cuda_stream);
/* pack + send over peer memory (NVLink P2P stores): entry k of the send order lies in chunk c = the c-th neighbour,
   chunk_start[c] <= k < chunk_start[c+1], and is written to ((Number *)peer_dst[c])[k - chunk_start[c]], a pointer into that
   neighbour's receive buffer mapped into this process (CUDA IPC / symmetric memory) */
int mfg_exchange_push_stream(mfg_exchange *ex, const void *vec_dev, const uint64_t *peer_dst, const uint32_t *chunk_start, int n_chunks,
                             void *cuda_stream);
/* send[k] = vec[pack_idx[k]] on another CUDA stream */
int mfg_exchange_pack_stream(mfg_exchange *ex, const void *vec_dev, void *send_dev, void *cuda_stream);
/* owned-DoF dot product support: mask[i] = 1 if this rank owns DoF i (lowest rank touching it) */
int mfg_vec_dot_masked(const mfg_vec *a, const mfg_vec *b, const uint8_t *owned_mask_dev, double *out);
/* CG over the partition (SolverCG control flow, poisson.cu:233-260, one process per GPU): the three vector kernels of
 * mfg_solver_cg with sums over the OWNED DoFs, and the scalars in a device block `scal` of 9 doubles ([0] d.h, [1] g.g,
 * [2] g.z, [3] |g|, [4] previous g.z, [5] alpha, [6] beta, [7] iteration of convergence or -1, [8] scratch) that the caller
 * all-reduces in stream order (NCCL) between the kernels -- entries [0] after mfg_cgd_dot, [1..2] after mfg_cgd_residual --
 * so that no rank reads a scalar on the host inside the loop.  scal needs 10 doubles ([9] = iteration counter on the device: with
 * it < 0 mfg_cgd_beta / mfg_cgd_advance take the iteration number from it, so one captured CUDA graph serves every iteration).
 * All kernels run on the context stream. */
int mfg_cgd_init(mfg_ctx *ctx, double *scal_dev);
int mfg_cgd_dot(mfg_ctx *ctx, mfg_dtype dt, const void *d, const void *h, const uint8_t *owned_dev, size_t n, double *scal_dev);
int mfg_cgd_alpha(mfg_ctx *ctx, double *scal_dev);                 /* alpha = g.z / d.h */
int mfg_cgd_residual(mfg_ctx *ctx, mfg_dtype dt, void *g, void *h, const void *minv, const uint8_t *owned_dev, size_t n, double *scal_dev, int first);
int mfg_cgd_beta(mfg_ctx *ctx, double *scal_dev, double tol, int it);  /* |g|, convergence flag, beta */
int mfg_cgd_advance(mfg_ctx *ctx, mfg_dtype dt, void *x, void *d, const void *z, size_t n, const double *scal_dev, int it);

/* ---- multigrid with local smoothing on an adaptively refined mesh (poisson_mg.cu / bmop_mg.cu with an adaptive grid) -------------
 * Built from the host hierarchy of mfg_amesh_build_mg: per level a LaplaceOperatorGpu::reinit(dof_handler, mg_constrained_dofs, level)
 * (laplace_operator_gpu.h:153-186; constraint handler = boundary + refinement-edge DoFs, constraint_handler_gpu.cu:99-123), its
 * interface operators vmult_interface_down / vmult_interface_up (:306-352), MGTransferMatrixFreeGpu with copy_to_mg / copy_from_mg
 * through index pairs (mg_transfer_matrix_free_gpu.cu:688-757), Chebyshev smoothers, Multigrid::level_v_step with
 * set_edge_matrices(interface, interface) (poisson_mg.cu:365-375), coarse solve = CG to a 1e-10 reduction.
 * mfg_amg_vcycle = PreconditionMG::vmult on vectors of the ACTIVE mesh; mfg_amg_solve_cg = SolverCG on the active operator
 * (hanging nodes resolved in its gather / scatter) preconditioned by it. */
typedef struct mfg_amg mfg_amg;
int mfg_amg_create(mfg_ctx *ctx, mfg_amesh *am, int min_level, mfg_dtype dt, int smoother_degree, double smoothing_range, int eig_cg_n_iterations,
                   mfg_amg **out);
int mfg_amg_destroy(mfg_amg *mg);
int mfg_amg_vcycle(mfg_amg *mg, mfg_vec *dst, const mfg_vec *src);
int mfg_amg_active_operator(mfg_amg *mg, mfg_laplace **op);             /* borrowed: the operator on the active mesh */
int mfg_amg_level_operator(mfg_amg *mg, int level, mfg_laplace **op);   /* borrowed */
/* interface operators of a level (laplace_operator_gpu.h:306-352) on level vectors */
int mfg_amg_vmult_interface_down(mfg_amg *mg, int level, mfg_vec *dst, const mfg_vec *src);
int mfg_amg_vmult_interface_up(mfg_amg *mg, int level, mfg_vec *dst, const mfg_vec *src);
/* level transfer and copies, for tests and drivers: prolongate to `level` from level-1, restrict_and_add from `level` into level-1 */
int mfg_amg_prolongate(mfg_amg *mg, int level, mfg_vec *dst_fine, const mfg_vec *src_coarse);
int mfg_amg_restrict_and_add(mfg_amg *mg, int level, mfg_vec *dst_coarse, const mfg_vec *src_fine);
int mfg_amg_copy_to_level(mfg_amg *mg, int level, mfg_vec *dst_level, const mfg_vec *src_active);     /* one level of copy_to_mg (dst zeroed first) */
int mfg_amg_copy_from_level(mfg_amg *mg, int level, mfg_vec *dst_active, const mfg_vec *src_level);   /* one level of copy_from_mg (dst not zeroed) */
int mfg_amg_info(const mfg_amg *mg, int level, double *lambda_max, long *coarse_iterations, size_t *n_dofs, size_t *n_edge);
int mfg_amg_solve_cg(mfg_amg *mg, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int *iters, double *last_residual, double *history);

#ifdef __cplusplus
}
#endif
#endif /* MFGPU_H */
