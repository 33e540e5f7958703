// emu_generic.cc -- TEST INFRASTRUCTURE: the generic FEEvaluationGpu path (device code of include/dealii_cuda_b200/fee_gpu.cuh,
// cut in front of its host launch functions by tests/test_generic_path_emulated.py) compiled for the CPU on tests/emu/cuda_emu.h,
// with the user functors of examples/generic_ops.cu (mass operator, the reference's Laplace LocalOperator).  The launch logic below
// mirrors cell_loop: one launch over the cells without a constraint mask, one over the cells that carry one.
#include "cuda_emu.h"
// the dynamic shared memory of the running block: the kernels declare `extern __shared__ ... fee_smem_raw[]` inside the namespace
namespace dealii_cuda_b200 { alignas(16) unsigned char fee_smem_raw[256 * 1024]; }
#include "fee_gpu_device_part.h"   // generated: the header up to (not including) its host functions

using namespace dealii_cuda_b200;

template <int dim, int fe_degree, typename Number> struct MassOp
{
  typedef FEEvaluationGpu<dim, fe_degree, Number> FEE;
  void cell_apply(Number *dst, const Number *src, const typename FEE::data_type *gpu_data, const unsigned int cell, SharedData<dim, Number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.read_dof_values(src);
    phi.evaluate(true, false);
    phi.apply_quad_point_operations(this);
    phi.integrate(true, false);
    phi.distribute_local_to_global(dst);
  }
  void quad_operation(FEE *phi, const unsigned int q) const { phi->submit_value(phi->get_value(q), q); }
};

template <int dim, int fe_degree, typename Number> struct LaplaceOp
{
  typedef FEEvaluationGpu<dim, fe_degree, Number> FEE;
  const Number *coefficient;
  void cell_apply(Number *dst, const Number *src, const typename FEE::data_type *gpu_data, const unsigned int cell, SharedData<dim, Number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.read_dof_values(src);
    phi.evaluate(false, true);
    phi.apply_quad_point_operations(this);
    phi.integrate(false, true);
    phi.distribute_local_to_global(dst);
  }
  void quad_operation(FEE *phi, const unsigned int q) const
  {
    typename FEE::gradient_type g = phi->get_gradient(q);
    const Number a = coefficient[phi->get_global_q(q)];
    for (int d = 0; d < dim; ++d) g[d] *= a;
    phi->submit_gradient(g, q);
  }
};

template <int dim, int p, typename LocOp>
void run(const LocOp &op, uint32_t n_cells, uint32_t n_plain, const uint32_t *l2g, const double *jxw, const double *inv_jac, const uint32_t *mask,
         const double *val, const double *grad, const double *hang, const double *src, double *dst)
{
  constexpr unsigned n = p + 1, npc = dim == 2 ? n * n : n * n * n;
  GpuData<dim, double> gd;
  gd.loc2glob = l2g; gd.JxW = jxw; gd.inv_jac = inv_jac; gd.quadrature_points = nullptr;
  gd.shape_values = gd.shape_gradients = nullptr;
  gd.general = 0; gd.use_coloring = 0; gd.constraint_mask = nullptr; gd.hanging_weights = nullptr;
  ShapeTables<double, n> tab;
  for (unsigned i = 0; i < n * n; ++i) { tab.val[i] = val[i]; tab.grad[i] = grad[i]; tab.hang[i] = hang[i]; }
  const unsigned cpb = npc >= 128 ? 1 : 128 / npc;
  auto kern = apply_kernel_shmem<LocOp, dim, p, double>;
  if (n_plain)
    {
      gd.n_cells = n_plain;
      emu_launch((n_plain + cpb - 1) / cpb, cpb * npc, kern, dst, src, op, gd, tab, 0u);
    }
  if (mask && n_plain < n_cells)
    {
      gd.constraint_mask = mask; gd.n_cells = n_cells;
      emu_launch((n_cells - n_plain + cpb - 1) / cpb, cpb * npc, kern, dst, src, op, gd, tab, n_plain);
    }
}

// which: 0 mass, 1 laplace (coefficient [n_cells][npc] in kernel cell order).  Arrays in kernel cell order (cells without a mask first).
extern "C" int emu_generic_apply(int which, int dim, int degree, uint32_t n_cells, uint32_t n_plain, const uint32_t *l2g, const double *jxw,
                                 const double *inv_jac, const uint32_t *mask, const double *val, const double *grad, const double *hang,
                                 const double *coefficient, const double *src, double *dst)
{
#define CASE(D, P)                                                                                                     \
  if (dim == D && degree == P)                                                                                         \
    {                                                                                                                  \
      if (which == 0) run<D, P>(MassOp<D, P, double>(), n_cells, n_plain, l2g, jxw, inv_jac, mask, val, grad, hang, src, dst); \
      else                                                                                                             \
        {                                                                                                              \
          LaplaceOp<D, P, double> op;                                                                                  \
          op.coefficient = coefficient;                                                                                \
          run<D, P>(op, n_cells, n_plain, l2g, jxw, inv_jac, mask, val, grad, hang, src, dst);                         \
        }                                                                                                              \
      return 0;                                                                                                        \
    }
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(3, 1) CASE(3, 2)
#undef CASE
  return -1;
}
