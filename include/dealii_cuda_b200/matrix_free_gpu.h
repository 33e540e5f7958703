// matrix_free_gpu.h -- C++ facade: HyperCubeMesh (stands in for Triangulation + DoFHandler + ConstraintMatrix),
// ConstraintHandlerGpu<Number> (constraint_handler_gpu.h:13-59), MatrixFreeGpu<dim,Number>
// (matrix_free_gpu.h:81-229) and LaplaceOperatorGpu<dim,fe_degree,Number> (laplace_operator_gpu.h:35-96).
#pragma once
#include <memory>
#include <vector>
#include "gpu_vec.h"

namespace dealii_cuda_b200 {

// GridGenerator::hyper_cube(left,right) + refine_global(n_refine) + FE_Q(fe_degree) + distribute_dofs +
// interpolate_boundary_values(0, ZeroFunction)   (bmop.cu:111-132, poisson_common.h:58-72)
template <int dim> class HyperCubeMesh
{
public:
  HyperCubeMesh(unsigned int fe_degree, unsigned int n_refine, double left = -1., double right = 1.)
  {
    check(mfg_mesh_hyper_cube(default_context(), dim, (int)fe_degree, (int)n_refine, left, right, &m_));
  }
  // a box of 2^k cells per direction with Dirichlet conditions on the faces named in the descriptor (partitions of the cube:
  // BoxPartition::box() of distributed.h)
  HyperCubeMesh(unsigned int fe_degree, const mfg_box_desc &box)
  {
    mfg_box_desc d = box;
    d.dim = dim; d.degree = (int)fe_degree;
    check(mfg_mesh_create_box(default_context(), &d, &m_));
  }
  ~HyperCubeMesh() { if (m_) mfg_mesh_destroy(m_); }
  HyperCubeMesh(const HyperCubeMesh &) = delete;
  unsigned int n_dofs() const { return mfg_mesh_n_dofs(m_); }
  unsigned int n_active_cells() const { return mfg_mesh_n_cells(m_); }
  unsigned int n_constraints() const { return mfg_mesh_n_constrained(m_); }
  std::vector<unsigned int> get_loc2glob() const
  {
    std::vector<unsigned int> h((size_t)mfg_mesh_n_cells(m_) * mfg_mesh_dofs_per_cell(m_));
    check(mfg_mesh_get_loc2glob(m_, h.data()));
    return h;
  }
  std::vector<unsigned int> constrained_dofs() const
  {
    std::vector<unsigned int> h(mfg_mesh_n_constrained(m_));
    check(mfg_mesh_get_constrained(m_, h.data()));
    return h;
  }
  mfg_mesh *handle() const { return m_; }

private:
  mfg_mesh *m_ = nullptr;
};

// Triangulation with local refinement + DoFHandler + HangingNodes (the reference's adaptive-grid runs: bmop_common.h:9-120,
// matrix_free_gpu/hanging_nodes.cuh:209-454), host substrate of the library (mfg_amesh_*).  Usage mirrors deal.II:
//   AdaptiveMesh<3> mesh(4); mesh.pseudo_adaptive_refinement(6); mesh.distribute_dofs(); op.reinit(mesh);
template <int dim> class AdaptiveMesh
{
public:
  enum MeshSmoothing { none = 0, limit_level_difference_at_vertices = 1 };  // Triangulation::MeshSmoothing (poisson_mg.cu:132)
  explicit AdaptiveMesh(unsigned int fe_degree, MeshSmoothing smoothing = none, double left = -1., double right = 1.)
  {
    check(mfg_amesh_create(dim, (int)fe_degree, left, right, &m_));
    if (smoothing == limit_level_difference_at_vertices) check(mfg_amesh_set_limit_level_difference_at_vertices(m_, 1));
  }
  ~AdaptiveMesh() { if (m_) mfg_amesh_destroy(m_); }
  AdaptiveMesh(const AdaptiveMesh &) = delete;
  void refine_global(unsigned int times = 1) { check(mfg_amesh_refine_global(m_, (int)times)); }
  void set_refine_flags(const std::vector<unsigned char> &flags) { check(mfg_amesh_set_refine_flags(m_, flags.data(), flags.size())); }
  void mark_cells_in_annulus(double R, double r = 0.0, const double *center = nullptr) { check(mfg_amesh_mark_cells_in_annulus(m_, R, r, center)); }
  void mark_cells_on_shell(double R, const double *center = nullptr) { check(mfg_amesh_mark_cells_on_shell(m_, R, center)); }
  void mark_octant() { check(mfg_amesh_mark_octant(m_)); }
  void execute_coarsening_and_refinement() { check(mfg_amesh_execute_refinement(m_)); }
  void pseudo_adaptive_refinement(int n_ref) { check(mfg_amesh_pseudo_adaptive_refinement(m_, n_ref)); }
  void distribute_dofs() { check(mfg_amesh_distribute_dofs(m_)); }
  unsigned int n_active_cells() const { return mfg_amesh_n_active_cells(m_); }
  unsigned int n_levels() const { return mfg_amesh_n_levels(m_); }
  unsigned int n_dofs() const { return mfg_amesh_n_dofs(m_); }
  unsigned int n_constraints() const { return mfg_amesh_n_constrained(m_); }
  // hanging + Dirichlet boundary DoFs, ascending (the ConstraintHandlerGpu list, constraint_handler_gpu.cu:77-83)
  std::vector<unsigned int> constrained_dofs() const
  {
    std::vector<unsigned int> c(mfg_amesh_n_constrained(m_));
    check(mfg_amesh_get_arrays(m_, nullptr, nullptr, nullptr, c.data(), nullptr, nullptr, nullptr, nullptr));
    return c;
  }
  // DoFs on the boundary of the domain, ascending (what VectorTools::interpolate_boundary_values visits, poisson.cu:155-158)
  std::vector<unsigned int> boundary_dofs() const
  {
    std::vector<unsigned int> b(mfg_amesh_n_boundary(m_));
    check(mfg_amesh_get_boundary(m_, b.data()));
    return b;
  }
  // DoFTools::map_dofs_to_support_points: [n_dofs][dim]
  std::vector<double> support_points() const
  {
    std::vector<double> x((size_t)mfg_amesh_n_dofs(m_) * dim);
    check(mfg_amesh_get_support_points(m_, x.data()));
    return x;
  }
  mfg_amesh *handle() const { return m_; }

private:
  mfg_amesh *m_ = nullptr;
};

// GridGenerator::hyper_ball + SphericalManifold on the boundary + refine_global (the reference's -DBALL_GRID, poisson_common.h:59-72):
// unstructured mesh with FE_Q DoFs and MappingQ1 geometry, host substrate of the library (mfg_umesh_*)
template <int dim> class BallMesh
{
public:
  BallMesh(unsigned int fe_degree, unsigned int n_refine, double radius = 1.)
  {
    check(mfg_umesh_hyper_ball(dim, (int)fe_degree, radius, &m_));
    check(mfg_umesh_refine_global(m_, (int)n_refine));
    check(mfg_umesh_distribute_dofs(m_));
  }
  ~BallMesh() { if (m_) mfg_umesh_destroy(m_); }
  BallMesh(const BallMesh &) = delete;
  unsigned int n_dofs() const { return mfg_umesh_n_dofs(m_); }
  unsigned int n_active_cells() const { return mfg_umesh_n_cells(m_); }
  unsigned int n_constraints() const { return mfg_umesh_n_boundary(m_); }
  std::vector<unsigned int> boundary_dofs() const
  {
    std::vector<unsigned int> b(mfg_umesh_n_boundary(m_));
    check(mfg_umesh_get_arrays(m_, nullptr, b.data(), nullptr, nullptr, nullptr, nullptr));
    return b;
  }
  std::vector<double> support_points() const   // DoFTools::map_dofs_to_support_points: [n_dofs][dim]
  {
    std::vector<double> x((size_t)mfg_umesh_n_dofs(m_) * dim);
    check(mfg_umesh_get_support_points(m_, x.data()));
    return x;
  }
  mfg_umesh *handle() const { return m_; }

private:
  mfg_umesh *m_ = nullptr;
};

template <typename Number> class ConstraintHandlerGpu
{
public:
  ~ConstraintHandlerGpu() { if (ch_) mfg_ch_destroy(ch_); }
  // reinit(ConstraintMatrix, n_dofs): ascending list of constrained DoFs (constraint_handler_gpu.cu:69-95)
  void reinit(const std::vector<unsigned int> &constrained, unsigned int /*n_dofs*/,
              const std::vector<unsigned int> &edge = std::vector<unsigned int>())
  {
    if (ch_) mfg_ch_destroy(ch_);
    ch_ = nullptr;
    check(mfg_ch_create(default_context(), dtype_of<Number>(), constrained.data(), constrained.size(), edge.data(), edge.size(), &ch_));
  }
  template <int dim> void reinit(const HyperCubeMesh<dim> &mesh)
  {
    if (ch_) mfg_ch_destroy(ch_);
    ch_ = nullptr;
    check(mfg_ch_create_from_mesh(default_context(), dtype_of<Number>(), mesh.handle(), &ch_));
  }
  void set_constrained_values(GpuVector<Number> &v, Number val) const { check(mfg_ch_set_constrained_values(ch_, v.handle(), (double)val)); }
  void save_constrained_values(GpuVector<Number> &v) { check(mfg_ch_save_constrained_values(ch_, v.handle())); }
  void save_constrained_values(const GpuVector<Number> &v1, GpuVector<Number> &v2) { check(mfg_ch_save_constrained_values2(ch_, v1.handle(), v2.handle())); }
  void load_constrained_values(GpuVector<Number> &v) const { check(mfg_ch_load_constrained_values(ch_, v.handle())); }
  void load_and_add_constrained_values(GpuVector<Number> &v1, GpuVector<Number> &v2) const { check(mfg_ch_load_and_add_constrained_values(ch_, v1.handle(), v2.handle())); }
  void copy_edge_values(GpuVector<Number> &dst, const GpuVector<Number> &src) const { check(mfg_ch_copy_edge_values(ch_, dst.handle(), src.handle())); }
  mfg_ch *handle() const { return ch_; }

private:
  mfg_ch *ch_ = nullptr;
};

template <int dim, typename Number> class MatrixFreeGpu
{
public:
  enum ParallelizationScheme { scheme_par_in_elem, scheme_par_over_elems };
  struct AdditionalData
  {
    AdditionalData(ParallelizationScheme s = scheme_par_in_elem, bool use_coloring = false) : parallelization_scheme(s), use_coloring(use_coloring) {}
    ParallelizationScheme parallelization_scheme;
    bool                  use_coloring;
  };
  unsigned int n_cells_tot = 0, n_dofs = 0, fe_degree = 0, dofs_per_cell = 0, qpts_per_cell = 0, num_colors = 0;
  bool         use_coloring = false;

  ~MatrixFreeGpu() { free(); }
  // reinit(dof_handler, constraints, quad, additional_data)  (matrix_free_gpu.cu:448-563)
  void reinit(const HyperCubeMesh<dim> &mesh, const AdditionalData ad = AdditionalData())
  {
    free();
    check(mfg_mf_reinit_from_mesh(default_context(), mesh.handle(), dtype_of<Number>(), ad.use_coloring ? MFG_SCATTER_COLOR : MFG_SCATTER_ATOMIC, &mf_));
    fill_counters(ad.use_coloring);
  }
  // the ball mesh: full J^-1 and JxW per quadrature point (general geometry), quadrature points
  void reinit(const BallMesh<dim> &mesh)
  {
    free();
    check(mfg_mf_reinit_from_umesh(default_context(), mesh.handle(), dtype_of<Number>(), &mf_));
    fill_counters(false);
  }
  // adaptively refined mesh of the library: constraint masks, rewritten loc2glob, quadrature points (atomic scatter)
  void reinit(const AdaptiveMesh<dim> &mesh)
  {
    free();
    check(mfg_mf_reinit_from_amesh(default_context(), mesh.handle(), dtype_of<Number>(), &mf_));
    fill_counters(false);
  }
  // explicit arrays, as ReinitHelper extracts them from deal.II (matrix_free_gpu.cu:283-339)
  void reinit(const mfg_mf_desc &desc)
  {
    free();
    check(mfg_mf_reinit(default_context(), &desc, &mf_));
    fe_degree = desc.degree;
    fill_counters(desc.scatter == MFG_SCATTER_COLOR);
  }
  void        free() { if (mf_) mfg_mf_destroy(mf_); mf_ = nullptr; }
  std::size_t memory_consumption() const { return mfg_mf_memory_consumption(mf_); }
  mfg_mf     *handle() const { return mf_; }

private:
  void fill_counters(bool coloring)
  {
    n_cells_tot = mfg_mf_n_cells(mf_); n_dofs = mfg_mf_n_dofs(mf_); num_colors = mfg_mf_n_colors(mf_); use_coloring = coloring;
  }
  mfg_mf *mf_ = nullptr;
};

template <int dim, int fe_degree, typename Number> class LaplaceOperatorGpu
{
public:
  typedef Number            value_type;
  typedef GpuVector<Number> VectorType;

  explicit LaplaceOperatorGpu(bool use_coloring = false) : use_coloring_(use_coloring) {}
  ~LaplaceOperatorGpu() { clear(); }
  void clear() { if (op_) mfg_laplace_destroy(op_); op_ = nullptr; }
  // reinit(dof_handler, constraints)  (laplace_operator_gpu.h:120-151)
  void reinit(const HyperCubeMesh<dim> &mesh)
  {
    clear();
    check(mfg_laplace_create(default_context(), mesh.handle(), dtype_of<Number>(), use_coloring_ ? MFG_SCATTER_COLOR : MFG_SCATTER_ATOMIC, &op_));
  }
  // explicit arrays: MatrixFreeGpu built from mfg_mf_desc, constraint handler, coefficient at the quadrature points
  void reinit(const MatrixFreeGpu<dim, Number> &data, const ConstraintHandlerGpu<Number> &ch, const std::vector<double> &coefficient)
  {
    clear();
    check(mfg_laplace_create_from_arrays(default_context(), data.handle(), ch.handle(), coefficient.data(), &op_));
  }
  // adaptively refined mesh with hanging nodes (-DMATRIX_FREE_HANGING_NODES): masks, rewritten loc2glob, hanging + boundary
  // constraints and the coefficient come from the mesh object
  void reinit(const AdaptiveMesh<dim> &mesh)
  {
    clear();
    check(mfg_laplace_create_from_amesh(default_context(), mesh.handle(), dtype_of<Number>(), &op_));
  }
  // the ball mesh (-DBALL_GRID): non-affine cells, full J^-1 per quadrature point (fee_gpu.cuh:236-240, 276-280)
  void reinit(const BallMesh<dim> &mesh)
  {
    clear();
    check(mfg_laplace_create_from_umesh(default_context(), mesh.handle(), dtype_of<Number>(), &op_));
  }
  unsigned int m() const { return mfg_laplace_m(op_); }
  unsigned int n() const { return mfg_laplace_m(op_); }
  // "we cannot access matrix elements of a matrix free operator directly" (laplace_operator_gpu.h:69-74)
  Number el(const unsigned int, const unsigned int) const { throw std::runtime_error("LaplaceOperatorGpu::el: not implemented (matrix-free operator)"); }
  // edge matrices of the level operators (laplace_operator_gpu.h:306-352): they couple a level to the refinement edge of an
  // adaptively refined hierarchy.  The level operators of this library live on globally refined meshes, which have no
  // refinement edges: both products are zero.
  void vmult_interface_down(VectorType &dst, const VectorType &) const { dst = Number(0); }
  void vmult_interface_up(VectorType &dst, const VectorType &) const { dst = Number(0); }
  void vmult(VectorType &dst, const VectorType &src) const { check(mfg_laplace_vmult(op_, dst.handle(), src.handle())); }
  void Tvmult(VectorType &dst, const VectorType &src) const { vmult(dst, src); }
  void vmult_add(VectorType &dst, const VectorType &src) const { check(mfg_laplace_vmult_add(op_, dst.handle(), src.handle())); }
  void Tvmult_add(VectorType &dst, const VectorType &src) const { vmult_add(dst, src); }
  void compute_diagonal() { check(mfg_laplace_compute_diagonal(op_)); }
  // get_diagonal_inverse (laplace_operator_gpu.h:423-429): DiagonalMatrix over a view of the operator's inverse diagonal
  // (valid while the operator lives); Assert(diagonal_is_available) -> throws before compute_diagonal()
  std::shared_ptr<DiagonalMatrix<VectorType>> get_diagonal_inverse() const
  {
    mfg_vec *v = nullptr;
    check(mfg_laplace_get_diagonal_inverse(op_, &v));
    return std::make_shared<DiagonalMatrix<VectorType>>(VectorType::borrowed(v));
  }
  std::size_t  memory_consumption() const { return mfg_laplace_memory_consumption(op_); }
  mfg_laplace *handle() const { return op_; }

private:
  bool         use_coloring_;
  mfg_laplace *op_ = nullptr;
};

// MGTransferMatrixFreeGpu<dim,Number> (mg_transfer_matrix_free_gpu.h:64-307) for a globally refined hierarchy:
// build(meshes of consecutive levels, coarsest first), prolongate(to_level, dst, src), restrict_and_add(from_level, dst, src);
// copy_to_mg / copy_from_mg are plain copies there (the finest level is the active mesh, .cu:688-757).
template <int dim, typename Number> class MGTransferMatrixFreeGpu
{
public:
  MGTransferMatrixFreeGpu() {}
  ~MGTransferMatrixFreeGpu() { clear(); }
  MGTransferMatrixFreeGpu(const MGTransferMatrixFreeGpu &) = delete;
  void clear()
  {
    for (mfg_mgt *t : t_) mfg_mgt_destroy(t);
    t_.clear();
  }
  // levels[l] = mesh of level min_level + l
  void build(const std::vector<const HyperCubeMesh<dim> *> &levels, unsigned int min_level = 0)
  {
    clear();
    min_level_ = min_level;
    for (std::size_t l = 0; l + 1 < levels.size(); ++l)
      {
        mfg_mgt *t = nullptr;
        check(mfg_mgt_build(default_context(), levels[l]->handle(), levels[l + 1]->handle(), dtype_of<Number>(), &t));
        t_.push_back(t);
      }
  }
  void prolongate(const unsigned int to_level, GpuVector<Number> &dst, const GpuVector<Number> &src) const
  {
    check(mfg_mgt_prolongate(t_.at(to_level - min_level_ - 1), dst.handle(), src.handle()));
  }
  void restrict_and_add(const unsigned int from_level, GpuVector<Number> &dst, const GpuVector<Number> &src) const
  {
    check(mfg_mgt_restrict_and_add(t_.at(from_level - min_level_ - 1), dst.handle(), src.handle()));
  }
  void copy_to_mg(std::vector<GpuVector<Number>> &dst_levels, const GpuVector<Number> &src) const { dst_levels.back() = src; }
  void copy_from_mg(GpuVector<Number> &dst, const std::vector<GpuVector<Number>> &src_levels) const { dst = src_levels.back(); }

private:
  std::vector<mfg_mgt *> t_;
  unsigned int           min_level_ = 0;
};

// Multigrid + PreconditionMG + MGTransferMatrixFreeGpu + level LaplaceOperatorGpu objects of the reference's poisson_mg.cu /
// bmop_mg.cu on an ADAPTIVELY refined mesh (local smoothing: level operators with MGConstrainedDoFs, edge matrices
// vmult_interface_down / up, index-based copy_to_mg / copy_from_mg), owned by the library (mfg_amg_*).  The mesh must be built
// with AdaptiveMesh<dim>::limit_level_difference_at_vertices like the reference's Triangulation (poisson_mg.cu:132).
//   AdaptiveMesh<3> mesh(4, AdaptiveMesh<3>::limit_level_difference_at_vertices); ...refine...; mesh.distribute_dofs();
//   AdaptiveMultigrid<3, double> mg(mesh);  mg.vmult(dst, src) /* PreconditionMG::vmult */;  mg.solve_cg(x, b, tol);
template <int dim, typename Number> class AdaptiveMultigrid
{
public:
  explicit AdaptiveMultigrid(const AdaptiveMesh<dim> &mesh, unsigned int min_level = 0, unsigned int smoother_degree = 5, double smoothing_range = 15.,
                             unsigned int eig_cg_n_iterations = 15)
  {
    check(mfg_amg_create(default_context(), mesh.handle(), (int)min_level, dtype_of<Number>(), (int)smoother_degree, smoothing_range,
                         (int)eig_cg_n_iterations, &mg_));
  }
  ~AdaptiveMultigrid() { if (mg_) mfg_amg_destroy(mg_); }
  AdaptiveMultigrid(const AdaptiveMultigrid &) = delete;
  unsigned int m() const { mfg_laplace *op = nullptr; check(mfg_amg_active_operator(mg_, &op)); return mfg_laplace_m(op); }
  // the operator on the active mesh (hanging nodes resolved in gather / scatter)
  void vmult_active(GpuVector<Number> &dst, const GpuVector<Number> &src) const
  {
    mfg_laplace *op = nullptr;
    check(mfg_amg_active_operator(mg_, &op));
    check(mfg_laplace_vmult(op, dst.handle(), src.handle()));
  }
  void vmult(GpuVector<Number> &dst, const GpuVector<Number> &src) const { check(mfg_amg_vcycle(mg_, dst.handle(), src.handle())); }  // PreconditionMG::vmult
  void vmult_interface_down(unsigned int level, GpuVector<Number> &dst, const GpuVector<Number> &src) const { check(mfg_amg_vmult_interface_down(mg_, (int)level, dst.handle(), src.handle())); }
  void vmult_interface_up(unsigned int level, GpuVector<Number> &dst, const GpuVector<Number> &src) const { check(mfg_amg_vmult_interface_up(mg_, (int)level, dst.handle(), src.handle())); }
  void prolongate(unsigned int to_level, GpuVector<Number> &dst, const GpuVector<Number> &src) const { check(mfg_amg_prolongate(mg_, (int)to_level, dst.handle(), src.handle())); }
  void restrict_and_add(unsigned int from_level, GpuVector<Number> &dst, const GpuVector<Number> &src) const { check(mfg_amg_restrict_and_add(mg_, (int)from_level, dst.handle(), src.handle())); }
  // SolverCG + PreconditionMG (poisson_mg.cu:504-518); returns the number of iterations
  int solve_cg(GpuVector<Number> &x, const GpuVector<Number> &b, double abs_tol, int max_iter = 1000) const
  {
    int its = 0;
    check(mfg_amg_solve_cg(mg_, x.handle(), b.handle(), abs_tol, max_iter, &its, nullptr, nullptr));
    return its;
  }
  mfg_amg *handle() const { return mg_; }

private:
  mfg_amg *mg_ = nullptr;
};

}  // namespace dealii_cuda_b200
