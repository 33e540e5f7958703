// partitioned_mg.cc -- multigrid-preconditioned CG over the box partition on the C++ facade (include/dealii_cuda_b200/partitioned_mg.h):
// all boxes of the partition in this process, on one device (LocalWorldLevel).   usage: partitioned_mg <boxes> <refinement> [strong]
// weak (default): a refine_global(r) cube of cells per box; strong: the refine_global(r) cube [-1,1]^dim cut into the boxes.
// -DDEGREE_FE, -DDIMENSION as in bmop.cc.  Prints: boxes, dim, degree, global DoFs, MG-CG iterations, coarse CG iterations, error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../include/dealii_cuda_b200/partitioned_mg.h"

using namespace dealii_cuda_b200;

#ifndef DEGREE_FE
#define DEGREE_FE 4
#endif
#ifndef DIMENSION
#define DIMENSION 3
#endif
typedef double number;

int main(int argc, char **argv)
{
  try
    {
      const int  world = argc > 1 ? std::atoi(argv[1]) : 2, r = argc > 2 ? std::atoi(argv[2]) : 2;
      const bool strong = argc > 3 && !std::strcmp(argv[3], "strong");
      PartitionedMultigrid<DIMENSION, DEGREE_FE, number> mg(world, 1, r, strong);
      const auto &L = mg.finest();
      // u = 1 on the free DoFs, 0 on the Dirichlet boundary; b = A u
      Field<number> u = L.new_field(), b = L.new_field(), x = L.new_field();
      for (int a = 0; a < world; ++a)
        {
          std::vector<number> h(L.parts()[a].n, 1.0);
          for (unsigned int c : L.parts()[a].mesh->constrained_dofs()) h[c] = 0.0;
          u[a] = h;
          x[a] = number(0);
        }
      L.vmult(b, u);
      const double bn = std::sqrt(L.dot(b, b));
      double       res = 0;
      const int    its = mg.solve_cg(x, b, 1e-10 * bn, 50, &res);
      for (int a = 0; a < world; ++a) x[a].add(number(-1), u[a]);
      const double err = std::sqrt(L.dot(x, x) / L.dot(u, u));
      std::printf("partitioned mg: %d boxes (%s)\t%d\t%d\t%llu dofs\t%d iterations\t%ld coarse iterations\tlambda_max %.6f\terror %.3e\n", world,
                  strong ? "strong" : "weak", DIMENSION, DEGREE_FE, L.n_global(), its, mg.coarse_iterations(), r > 1 ? mg.lambda_max(r) : 0.0, err);
      return its <= 20 && err <= 1e-7 ? 0 : 2;
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "Exception on processing:\n  %s\nAborting!\n", e.what());
      return 1;
    }
}
