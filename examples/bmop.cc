// bmop.cc -- the reference's benchmark driver (bmop.cu:66-230) on the C++ facade: host code is plain C++
// (g++), all device work happens inside libmfgpu.so.   usage: bmop <max_refinement> [min_refinement]
// -DADAPTIVE_GRID: the reference's pseudo-adaptive mesh with hanging nodes (BASELINE configs[3]); -DBALL_GRID: hyper_ball with non-affine
// cells; -DDEGREE_FE, -DDIMENSION, -DBMOP_USE_FLOATS as there
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "../include/dealii_cuda_b200/matrix_free_gpu.h"

using namespace dealii_cuda_b200;

#ifndef DEGREE_FE
#define DEGREE_FE 4
#endif
#ifndef DIMENSION
#define DIMENSION 3
#endif
#ifdef BMOP_USE_FLOATS
typedef float number;
#else
typedef double number;
#endif
#define N_ITERATIONS 100

template <int dim, int fe_degree> void run(int n_ref)
{
#if defined(BALL_GRID)
  BallMesh<dim> mesh(fe_degree, n_ref);        // bmop_setup_mesh(domain = BALL) (bmop.cu:164-168): hyper_ball + SphericalManifold + refine_global
#elif defined(ADAPTIVE_GRID)
  AdaptiveMesh<dim> mesh(fe_degree);           // bmop_setup_mesh(..., pseudo_adaptive_grid = true, n_ref) (bmop.cu:170-181)
  mesh.pseudo_adaptive_refinement(n_ref);      // bmop_common.h:49-105
  mesh.distribute_dofs();                      // + make_hanging_node_constraints (bmop.cu:116-126)
#else
  HyperCubeMesh<dim> mesh(fe_degree, n_ref);  // bmop_setup_mesh + setup_system (bmop.cu:111-132)
#endif
  LaplaceOperatorGpu<dim, fe_degree, number> system_matrix;
  system_matrix.reinit(mesh);
  GpuVector<number> src(system_matrix.n()), dst(system_matrix.n());
  check(mfg_ctx_synchronize(default_context()));
  const auto t0 = std::chrono::steady_clock::now();
  dst = number(0.1);  // IC (bmop.cu:140)
  for (int i = 0; i < N_ITERATIONS; ++i)
    {
      dst.swap(src);
      system_matrix.vmult(dst, src);
    }
  check(mfg_ctx_synchronize(default_context()));
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("%d\t%d\t%u\t%g\n", dim, fe_degree, mesh.n_dofs(), sec / N_ITERATIONS);
}

// bmop_mg's transfer (bmop_mg.cu): prolongate / restrict_and_add between two levels through the facade; checks <P u, v> = <u, R v>
template <int dim, int fe_degree> void mg_transfer_smoke()
{
  HyperCubeMesh<dim> coarse(fe_degree, 1), fine(fe_degree, 2);
  MGTransferMatrixFreeGpu<dim, number> transfer;
  transfer.build({&coarse, &fine}, 1);
  GpuVector<number> uc(coarse.n_dofs()), vf(fine.n_dofs()), pu(fine.n_dofs()), rv(coarse.n_dofs());
  uc = number(1);
  vf = number(0.5);
  transfer.prolongate(2, pu, uc);
  rv = number(0);
  transfer.restrict_and_add(2, rv, vf);
  const double a = (double)(pu * vf), b = (double)(uc * rv);
  std::printf("mg transfer adjointness: <Pu,v> = %.12g, <u,Rv> = %.12g\n", a, b);
}

#ifdef ADAPTIVE_GRID
// poisson_mg.cu on the pseudo-adaptive grid: SolverCG preconditioned by the multigrid V-cycle with local smoothing
// (Triangulation::limit_level_difference_at_vertices like poisson_mg.cu:132); right-hand side b = A u for u = 1 on the free DoFs
template <int dim, int fe_degree> void adaptive_mg_solve(int n_ref)
{
  AdaptiveMesh<dim> mesh(fe_degree, AdaptiveMesh<dim>::limit_level_difference_at_vertices);
  mesh.pseudo_adaptive_refinement(n_ref);
  mesh.distribute_dofs();
  AdaptiveMultigrid<dim, number> mg(mesh);
  // u = 1 on the free DoFs, 0 on the constrained ones (hanging + boundary): b = A u is then zero on the constrained rows, as the
  // reference's right-hand sides are after ConstraintMatrix::condense (the V-cycle returns 0 on hanging DoFs)
  std::vector<number> uh(mg.m(), number(1));
  for (unsigned int c : mesh.constrained_dofs()) uh[c] = number(0);
  GpuVector<number> u(uh), b(mg.m()), x(mg.m());
  mg.vmult_active(b, u);
  const double bnorm = (double)b.l2_norm();
  check(mfg_ctx_synchronize(default_context()));
  const auto t0 = std::chrono::steady_clock::now();
  const int its = mg.solve_cg(x, b, 1e-10 * bnorm, 100);
  check(mfg_ctx_synchronize(default_context()));
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  x -= u;
  std::printf("adaptive mg: %d\t%d\t%u cells\t%u dofs\t%u levels\t%d iterations\t%g s\terror %.3e\n", dim, fe_degree, mesh.n_active_cells(), mesh.n_dofs(),
              mesh.n_levels(), its, sec, (double)x.l2_norm() / (double)u.l2_norm());
}
#endif

int main(int argc, char **argv)
{
  try
    {
      const int max_refinement = argc > 1 ? std::atoi(argv[1]) : 1;
      const int min_refinement = argc > 2 ? std::atoi(argv[2]) : 0;
#ifdef ADAPTIVE_GRID
      if (argc > 3 && std::string(argv[3]) == "mg") { adaptive_mg_solve<DIMENSION, DEGREE_FE>(max_refinement); return 0; }
#else
      if (argc > 3 && std::string(argv[3]) == "mg") { mg_transfer_smoke<DIMENSION, DEGREE_FE>(); return 0; }
#endif
      for (int r = min_refinement; r <= max_refinement; ++r) run<DIMENSION, DEGREE_FE>(r);
    }
  catch (std::exception &exc)
    {
      std::fprintf(stderr, "Exception on processing:\n%s\nAborting!\n", exc.what());
      return 1;
    }
  return 0;
}
