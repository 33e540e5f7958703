// generic_ops.cu -- user-written cell functors on the header-only generic path (include/dealii_cuda_b200/fee_gpu.cuh),
// compiled with nvcc exactly like a driver of the reference compiles its LocalOperator into apply_kernel_shmem:
//   * MassOp           : (phi_i, u)                evaluate(true,false) / submit_value / integrate(true,false)
//   * LaplaceOp        : (a grad phi_i, grad u)    the reference's LocalOperator (laplace_operator_gpu.h:247-282) with the
//                                                  coefficient array indexed by get_global_q
//   * RhsOp            : (phi_i, f),  f(x) = 1 + x_0 + 2 x_1 [+ 3 x_2]   get_quadrature_point / submit_value / integrate
// Exported with C linkage for tests/test_gpu_generic_path.py (ctypes); tests compare against numpy restatements.
#include "../include/dealii_cuda_b200/fee_gpu.cuh"

using namespace dealii_cuda_b200;

template <int dim, int fe_degree, typename Number> struct MassOp
{
  typedef FEEvaluationGpu<dim, fe_degree, Number> FEE;
  __device__ void cell_apply(Number *dst, const Number *src, const typename FEE::data_type *gpu_data, const unsigned int cell,
                             SharedData<dim, Number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.read_dof_values(src);
    phi.evaluate(true, false);
    phi.apply_quad_point_operations(this);
    phi.integrate(true, false);
    phi.distribute_local_to_global(dst);
  }
  __device__ void quad_operation(FEE *phi, const unsigned int q) const { phi->submit_value(phi->get_value(q), q); }
};

template <int dim, int fe_degree, typename Number> struct LaplaceOp
{
  typedef FEEvaluationGpu<dim, fe_degree, Number> FEE;
  const Number *coefficient;  // [n_cells][n_q_points], kernel cell order
  __device__ void cell_apply(Number *dst, const Number *src, const typename FEE::data_type *gpu_data, const unsigned int cell,
                             SharedData<dim, Number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.read_dof_values(src);
    phi.evaluate(false, true);
    phi.apply_quad_point_operations(this);
    phi.integrate(false, true);
    phi.distribute_local_to_global(dst);
  }
  __device__ void quad_operation(FEE *phi, const unsigned int q) const
  {
    typename FEE::gradient_type g = phi->get_gradient(q);
    const Number a = coefficient[phi->get_global_q(q)];
    for (int d = 0; d < dim; ++d) g[d] *= a;
    phi->submit_gradient(g, q);
  }
};

template <int dim, int fe_degree, typename Number> struct RhsOp
{
  typedef FEEvaluationGpu<dim, fe_degree, Number> FEE;
  __device__ void cell_apply(Number *dst, const Number *, const typename FEE::data_type *gpu_data, const unsigned int cell,
                             SharedData<dim, Number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.apply_quad_point_operations(this);
    phi.integrate(true, false);
    phi.distribute_local_to_global(dst);
  }
  __device__ void quad_operation(FEE *phi, const unsigned int q) const
  {
    const typename FEE::gradient_type x = phi->get_quadrature_point(q);
    Number f = 1;
    for (int d = 0; d < dim; ++d) f += Number(d + 1) * x[d];
    phi->submit_value(f, q);
  }
};

// the right-hand side integral without a source vector: MatrixFreeGpu::cell_loop(dst, loc_op) (matrix_free_gpu.h:382-393)
template <int dim, int fe_degree, typename Number> struct RhsOpDst
{
  typedef FEEvaluationGpu<dim, fe_degree, Number> FEE;
  __device__ void cell_apply(Number *dst, const typename FEE::data_type *gpu_data, const unsigned int cell, SharedData<dim, Number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.apply_quad_point_operations(this);
    phi.integrate(true, false);
    phi.distribute_local_to_global(dst);
  }
  __device__ void quad_operation(FEE *phi, const unsigned int q) const
  {
    const typename FEE::gradient_type x = phi->get_quadrature_point(q);
    Number f = 1;
    for (int d = 0; d < dim; ++d) f += Number(d + 1) * x[d];
    phi->submit_value(f, q);
  }
};

// the reference's LocalCoeffOp (laplace_operator_gpu.h:191-203) with Coefficient::value = 1 / (0.05 + 2 |x|^2)
// (poisson_common.h:146-158), for MatrixFreeGpu::evaluate_on_cells<Op> (matrix_free_gpu.h:415-435)
template <int dim, int fe_degree, typename Number> struct LocalCoeffOp
{
  static constexpr unsigned int n_q_points = FEEvaluationGpu<dim, fe_degree, Number>::n_q_points;
  static __device__ void eval(Number *coefficient, const Number *qpts)
  {
    for (unsigned int q = 0; q < n_q_points; ++q)
      {
        Number r2 = 0;
        for (int d = 0; d < dim; ++d) r2 += qpts[q * dim + d] * qpts[q * dim + d];
        coefficient[q] = Number(1) / (Number(0.05) + Number(2) * r2);
      }
  }
};

template <int dim, int p, typename Number> static void run(mfg_mf *mf, int which, void *dst, const void *src, const void *coef)
{
  Number *d = static_cast<Number *>(dst);
  const Number *s = static_cast<const Number *>(src);
  if (which == 0) cell_loop<dim, p, Number>(mf, d, s, MassOp<dim, p, Number>());
  else if (which == 1)
    {
      LaplaceOp<dim, p, Number> op;
      op.coefficient = static_cast<const Number *>(coef);
      cell_loop<dim, p, Number>(mf, d, s, op);
    }
  else if (which == 2) cell_loop<dim, p, Number>(mf, d, s, RhsOp<dim, p, Number>());
  else if (which == 3) cell_loop<dim, p, Number>(mf, d, RhsOpDst<dim, p, Number>());
  else evaluate_on_cells<dim, p, Number, LocalCoeffOp<dim, p, Number>>(mf, d);
}

// the facade form, as a driver of the reference would write it (compiled, not run by the tests)
template <int dim, int p, typename Number>
void mass_apply_on_facade(const MatrixFreeGpu<dim, Number> &data, GpuVector<Number> &dst, const GpuVector<Number> &src)
{
  dst = Number(0);
  cell_loop<dim, p>(data, dst, src, MassOp<dim, p, Number>());
}
template void mass_apply_on_facade<3, 4, double>(const MatrixFreeGpu<3, double> &, GpuVector<double> &, const GpuVector<double> &);
template <int dim, int p, typename Number> void coefficient_on_facade(const MatrixFreeGpu<dim, Number> &data, GpuVector<Number> &coefficient, GpuVector<Number> &rhs)
{
  evaluate_on_cells<dim, p, Number, LocalCoeffOp<dim, p, Number>>(data, coefficient);   // data.template evaluate_on_cells<LocalCoeffOp>(coefficient)
  rhs = Number(0);
  cell_loop<dim, p>(data, rhs, RhsOpDst<dim, p, Number>());                            // data.cell_loop(rhs, loc_op)
}
template void coefficient_on_facade<3, 4, double>(const MatrixFreeGpu<3, double> &, GpuVector<double> &, GpuVector<double> &);

// which: 0 mass, 1 laplace (coef_dev required), 2 right-hand side, 3 right-hand side through the dst-only cell_loop,
// 4 evaluate_on_cells<LocalCoeffOp> (dst_dev: [n_cells][n_q_points]).  Returns 0, or -1 with the message in generic_last_error().
static std::string g_err;
extern "C" const char *generic_last_error() { return g_err.c_str(); }
extern "C" int generic_apply(mfg_mf *mf, int which, int dim, int degree, int f64, void *dst_dev, const void *src_dev, const void *coef_dev)
{
  try
    {
#define CASE(D, P)                                                                                          \
  if (dim == D && degree == P)                                                                              \
    {                                                                                                       \
      if (f64) run<D, P, double>(mf, which, dst_dev, src_dev, coef_dev);                                     \
      else run<D, P, float>(mf, which, dst_dev, src_dev, coef_dev);                                          \
      return 0;                                                                                             \
    }
      CASE(2, 1) CASE(2, 2) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
      g_err = "generic_apply: (dim, degree) not instantiated in this example";
      return -1;
    }
  catch (const std::exception &e)
    {
      g_err = e.what();
      return -1;
    }
}
