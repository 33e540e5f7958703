// poisson.cu -- the reference's Poisson driver (poisson.cu:60-300) on the drop-in:
//   -div(a grad u) = f on [-1,1]^dim,  u = u_exact on the boundary,  a(x) = 1 / (0.05 + 2 |x|^2)   (poisson_common.h:146-158)
//   u_exact = sum of three Gaussians                                                              (poisson_common.cc:5-60)
// * operator: LaplaceOperatorGpu (precompiled sm_100a kernels of libmfgpu.so)
// * right-hand side with Dirichlet lifting, rhs_i = sum_q (phi_i f - a grad phi_i . grad g~) JxW   (poisson.cu:153-229):
//   ONE user-written functor on the header-only generic path (FEEvaluationGpu: evaluate gradients of the boundary-value
//   vector g~, submit_value(f), submit_gradient(-a grad g~), integrate(true, true))
// * solver: SolverCG with the inverse diagonal as preconditioner (mfg_solver_cg; poisson.cu:233-260)
// * error: || u_h - u_exact ||_L2 through a second functor (sum of (phi_i, e^2) over i = integral of e^2)
// usage: poisson <dim> <degree> <min_refinement> <max_refinement> [nonuniform | ball]     prints one line per refinement
// (nonuniform: the reference's grid_refinement = NONUNIFORM -- locally refined mesh with hanging nodes, poisson_common.h:76-92;
//  ball: domain = BALL -- hyper_ball with non-affine cells, poisson_common.h:65-70):
//        dim degree refinement n_dofs cg_iterations l2_error setup_seconds solve_seconds
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "../include/dealii_cuda_b200/fee_gpu.cuh"

using namespace dealii_cuda_b200;
typedef double number;

// ---- analytic data (host and device) -----------------------------------------------------------
template <int dim> struct Centers;
template <> struct Centers<2> { __host__ __device__ static double c(int i, int d) { const double v[3][2] = {{-0.5, 0.5}, {-0.5, -0.5}, {0.5, -0.5}}; return v[i][d]; } };
template <> struct Centers<3> { __host__ __device__ static double c(int i, int d) { const double v[3][3] = {{-0.5, 0.5, 0.25}, {-0.6, -0.5, -0.125}, {0.5, -0.5, 0.5}}; return v[i][d]; } };

template <int dim> __host__ __device__ inline double gauss_norm()
{
  const double w = 1.0 / 3.0, s = sqrt(2.0 * 3.14159265358979323846) * w;
  return dim == 2 ? s * s : s * s * s;
}
// u, grad u, laplace u of Solution<dim> (poisson_common.cc:31-174)
template <int dim> __host__ __device__ inline void solution(const double *x, double &u, double *gu, double &lu)
{
  const double w2 = 1.0 / 9.0;
  u = 0; lu = 0;
  for (int d = 0; d < dim; ++d) gu[d] = 0;
  for (int i = 0; i < 3; ++i)
    {
      double r2 = 0;
      for (int d = 0; d < dim; ++d) { const double t = x[d] - Centers<dim>::c(i, d); r2 += t * t; }
      const double e = exp(-r2 / w2) / gauss_norm<dim>();
      u += e;
      for (int d = 0; d < dim; ++d) gu[d] += -2.0 * (x[d] - Centers<dim>::c(i, d)) / w2 * e;
      lu += (-2.0 * dim / w2 + 4.0 * r2 / (w2 * w2)) * e;
    }
}
// a and grad a of Coefficient<dim> (poisson_common.h:146-175)
template <int dim> __host__ __device__ inline void coefficient(const double *x, double &a, double *ga)
{
  double r2 = 0;
  for (int d = 0; d < dim; ++d) r2 += x[d] * x[d];
  a = 1.0 / (0.05 + 2.0 * r2);
  for (int d = 0; d < dim; ++d) ga[d] = -4.0 * x[d] * a * a;
}

// ---- user functors on the generic path -----------------------------------------------------------
template <int dim, int fe_degree> struct RhsWithLifting
{
  typedef FEEvaluationGpu<dim, fe_degree, number> FEE;
  __device__ void cell_apply(number *dst, const number *lift, const typename FEE::data_type *gpu_data, const unsigned int cell,
                             SharedData<dim, number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.read_dof_values(lift);
    phi.evaluate(false, true);
    phi.apply_quad_point_operations(this);
    phi.integrate(true, true);
    phi.distribute_local_to_global(dst);
  }
  __device__ void quad_operation(FEE *phi, const unsigned int q) const
  {
    const typename FEE::gradient_type xq = phi->get_quadrature_point(q);
    double x[dim], u, gu[dim], lu, a, ga[dim];
    for (int d = 0; d < dim; ++d) x[d] = xq[d];
    solution<dim>(x, u, gu, lu);
    coefficient<dim>(x, a, ga);
    double f = -a * lu;  // f = -div(a grad u) = -a laplace u - grad a . grad u   (RightHandSide, poisson_common.h)
    for (int d = 0; d < dim; ++d) f -= ga[d] * gu[d];
    typename FEE::gradient_type g = phi->get_gradient(q);
    for (int d = 0; d < dim; ++d) g[d] *= -a;
    phi->submit_value(f, q);
    phi->submit_gradient(g, q);
  }
};

template <int dim, int fe_degree> struct SquaredError
{
  typedef FEEvaluationGpu<dim, fe_degree, number> FEE;
  __device__ void cell_apply(number *dst, const number *uh, const typename FEE::data_type *gpu_data, const unsigned int cell,
                             SharedData<dim, number> *shdata) const
  {
    FEE phi(cell, gpu_data, shdata);
    phi.read_dof_values(uh);
    phi.evaluate(true, false);
    phi.apply_quad_point_operations(this);
    phi.integrate(true, false);
    phi.distribute_local_to_global(dst);
  }
  __device__ void quad_operation(FEE *phi, const unsigned int q) const
  {
    const typename FEE::gradient_type xq = phi->get_quadrature_point(q);
    double x[dim], u, gu[dim], lu;
    for (int d = 0; d < dim; ++d) x[d] = xq[d];
    solution<dim>(x, u, gu, lu);
    const double e = phi->get_value(q) - u;
    phi->submit_value(e * e, q);
  }
};

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

enum GridCase { UNIFORM, NONUNIFORM, BALL };

template <int dim, int fe_degree> void run(int min_ref, int max_ref, GridCase grid)
{
  const bool nonuniform = grid == NONUNIFORM;
  for (int r = min_ref; r <= max_ref; ++r)
    {
      const double t0 = now();
      // make_grid + setup_system (poisson.cu:96-148); the mesh outlives the operator objects (they keep a pointer to it)
      std::unique_ptr<HyperCubeMesh<dim>> mesh;
      std::unique_ptr<AdaptiveMesh<dim>>  amesh;
      std::unique_ptr<BallMesh<dim>>      ball;
      LaplaceOperatorGpu<dim, fe_degree, number> system_matrix;
      MatrixFreeGpu<dim, number> data;                             // for the user-written cell loops
      ConstraintHandlerGpu<number> ch;
      std::vector<double> pts;
      std::vector<unsigned int> boundary;
      unsigned int n = 0;
      if (grid == BALL)
        {
          // domain = BALL (poisson_common.h:65-70): hyper_ball + SphericalManifold on the boundary, refine_global(r); non-affine cells
          ball.reset(new BallMesh<dim>(fe_degree, r));
          system_matrix.reinit(*ball);
          data.reinit(*ball);
          n = ball->n_dofs();
          pts = ball->support_points();
          boundary = ball->boundary_dofs();
          ch.reinit(boundary, n);
        }
      else if (!nonuniform)
        {
          mesh.reset(new HyperCubeMesh<dim>(fe_degree, r));
          system_matrix.reinit(*mesh);
          data.reinit(*mesh);
          n = mesh->n_dofs();
          pts.resize((size_t)n * dim);
          check(mfg_mesh_get_support_points(mesh->handle(), pts.data()));
          boundary = mesh->constrained_dofs();
          ch.reinit(*mesh);
        }
      else
        {
          // grid_refinement = NONUNIFORM (poisson_common.h:76-92): refine_global(r), then twice the cells of the octant x_d > 0.2;
          // hanging-node constraints join the Dirichlet ones (poisson.cu:139-146)
          amesh.reset(new AdaptiveMesh<dim>(fe_degree));
          amesh->refine_global(r);
          for (int k = 0; k < 2; ++k) { amesh->mark_octant(); amesh->execute_coarsening_and_refinement(); }
          amesh->distribute_dofs();
          system_matrix.reinit(*amesh);
          data.reinit(*amesh);
          n = amesh->n_dofs();
          pts = amesh->support_points();
          boundary = amesh->boundary_dofs();
          ch.reinit(amesh->constrained_dofs(), n);
        }
      // interpolate_boundary_values(Solution): g~ = u_exact at the boundary DoFs, 0 elsewhere (poisson.cu:155-158)
      std::vector<number> lift_host(n, 0.0);
      for (unsigned int c : boundary)
        {
          double u, gu[dim], lu;
          solution<dim>(&pts[(size_t)c * dim], u, gu, lu);
          lift_host[c] = u;
        }
      GpuVector<number> lift(lift_host), rhs(n), x(n), err(n);
      // assemble_system (poisson.cu:153-229); on the adaptive mesh read_dof_values / distribute_local_to_global interpolate
      rhs = number(0);
      cell_loop<dim, fe_degree>(data, rhs, lift, RhsWithLifting<dim, fe_degree>());
      // constrained rows of the operator are the identity: their right-hand side is the boundary value (hanging rows: 0)
      ch.set_constrained_values(rhs, 0);
      rhs += lift;  // lift is zero away from the boundary
      system_matrix.compute_diagonal();
      check(mfg_ctx_synchronize(default_context()));
      const double t1 = now();
      // solve (poisson.cu:233-293)
      x = number(0);
      int iters = 0;
      double res = 0;
      check(mfg_solver_cg(system_matrix.handle(), x.handle(), rhs.handle(), 1e-12 * (double)rhs.l2_norm(), 20000, 1, &iters, &res, nullptr));
      check(mfg_ctx_synchronize(default_context()));
      const double t2 = now();
      // L2 error: sum_i (phi_i, e^2) = integral of e^2
      err = number(0);
      cell_loop<dim, fe_degree>(data, err, x, SquaredError<dim, fe_degree>());
      GpuVector<number> ones(n);
      ones = number(1);
      const double l2 = std::sqrt((double)(err * ones));
      std::printf("%d %d %d %u %d %.6e %.3f %.3f\n", dim, fe_degree, r, n, iters, l2, t1 - t0, t2 - t1);
      std::fflush(stdout);
    }
}

int main(int argc, char **argv)
{
  const int dim = argc > 1 ? std::atoi(argv[1]) : 3, degree = argc > 2 ? std::atoi(argv[2]) : 4;
  const int min_ref = argc > 3 ? std::atoi(argv[3]) : 2, max_ref = argc > 4 ? std::atoi(argv[4]) : 4;
  const GridCase grid = argc > 5 && std::string(argv[5]) == "nonuniform" ? NONUNIFORM : argc > 5 && std::string(argv[5]) == "ball" ? BALL : UNIFORM;
  try
    {
      if (dim == 2 && degree == 1) run<2, 1>(min_ref, max_ref, grid);
      else if (dim == 2 && degree == 2) run<2, 2>(min_ref, max_ref, grid);
      else if (dim == 2 && degree == 4) run<2, 4>(min_ref, max_ref, grid);
      else if (dim == 3 && degree == 1) run<3, 1>(min_ref, max_ref, grid);
      else if (dim == 3 && degree == 2) run<3, 2>(min_ref, max_ref, grid);
      else if (dim == 3 && degree == 3) run<3, 3>(min_ref, max_ref, grid);
      else if (dim == 3 && degree == 4) run<3, 4>(min_ref, max_ref, grid);
      else { std::fprintf(stderr, "poisson: (dim, degree) not instantiated\n"); return 2; }
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "poisson: %s\n", e.what());
      return 1;
    }
  return 0;
}
