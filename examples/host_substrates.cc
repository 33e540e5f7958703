// host_substrates.cc -- the mesh substrates of the library through the C++ facade, WITHOUT a device: the adaptive triangulation
// (pseudo_adaptive_refinement, bmop_common.h:49-105), its multigrid hierarchy, and the BALL_GRID mesh (poisson_common.h:59-72).
// Prints the counts tests/test_adaptive_mesh.py compares with the Python binding.   usage: host_substrates <dim> <degree> <n_ref>
#include <cstdio>
#include <cstdlib>
#include "../include/dealii_cuda_b200/matrix_free_gpu.h"

using namespace dealii_cuda_b200;

template <int dim> void run(int degree, int n_ref)
{
  AdaptiveMesh<dim> mesh(degree, AdaptiveMesh<dim>::limit_level_difference_at_vertices);
  mesh.pseudo_adaptive_refinement(n_ref);
  mesh.distribute_dofs();
  std::printf("adaptive %d %d %d cells %u levels %u dofs %u constrained %u boundary %zu\n", dim, degree, n_ref, mesh.n_active_cells(), mesh.n_levels(),
              mesh.n_dofs(), mesh.n_constraints(), mesh.boundary_dofs().size());
  check(mfg_amesh_build_mg(mesh.handle(), 0));
  for (unsigned int l = 0; l < mesh.n_levels(); ++l)
    {
      uint32_t sz[6];
      check(mfg_amesh_mg_level_sizes(mesh.handle(), (int)l, sz));
      std::printf("level %u cells %u dofs %u boundary %u edge %u blocks %u copy %u\n", l, sz[0], sz[1], sz[2], sz[3], sz[4], sz[5]);
    }
  BallMesh<dim> ball(degree, n_ref > 2 ? n_ref - 2 : 0);
  std::printf("ball %d %d cells %u dofs %u boundary %u\n", dim, degree, ball.n_active_cells(), ball.n_dofs(), ball.n_constraints());
}

int main(int argc, char **argv)
{
  try
    {
      const int dim = argc > 1 ? std::atoi(argv[1]) : 3, degree = argc > 2 ? std::atoi(argv[2]) : 2, n_ref = argc > 3 ? std::atoi(argv[3]) : 4;
      if (dim == 2) run<2>(degree, n_ref);
      else run<3>(degree, n_ref);
    }
  catch (std::exception &exc)
    {
      std::fprintf(stderr, "Exception: %s\n", exc.what());
      return 1;
    }
  return 0;
}
