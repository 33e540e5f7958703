// partition_plan.cc -- host-only use of the multi-GPU facade (include/dealii_cuda_b200/distributed.h): the box partition and the
// exchange plan of every rank, without a device.  The lattice -> DoF map of a real run is mfg_mesh_lattice_to_dof; here a plain
// lexicographic numbering of the local lattice stands in, which is enough to exercise the plan (tests/test_partition.py compares the
// printed numbers with the Python binding).   usage: partition_plan <world> <dim> <degree> <refine> [strong]
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include "../include/dealii_cuda_b200/distributed.h"

using namespace dealii_cuda_b200;

int main(int argc, char **argv)
{
  try
    {
      const int  world = argc > 1 ? std::atoi(argv[1]) : 8, dim = argc > 2 ? std::atoi(argv[2]) : 3, degree = argc > 3 ? std::atoi(argv[3]) : 2;
      const int  r = argc > 4 ? std::atoi(argv[4]) : 2;
      const bool strong = argc > 5 && std::atoi(argv[5]) != 0;
      for (int rank = 0; rank < world; ++rank)
        {
          BoxPartition part(rank, world, dim, degree, r, strong);
          const mfg_box_desc &b = part.box();
          uint32_t            M[3] = {1, 1, 1};
          for (int d = 0; d < dim; ++d) M[d] = (uint32_t)degree * (1u << b.log2_cells[d]) + 1;
          const uint32_t n_local = M[0] * M[1] * M[2];
          ExchangePlan   plan(part, n_local, [&](const std::vector<uint32_t> &xyz) {
            std::vector<uint32_t> dofs(xyz.size() / 3);
            for (size_t k = 0; k < dofs.size(); ++k) dofs[k] = xyz[3 * k] + M[0] * (xyz[3 * k + 1] + M[1] * xyz[3 * k + 2]);
            return dofs;
          });
          const unsigned long long owned = std::accumulate(plan.owned_mask.begin(), plan.owned_mask.end(), 0ull);
          std::printf("rank %d coords %d %d %d n_local %u neighbors %zu n_send %zu shared %zu slots %zu owned %llu global %llu\n", rank, part.coords()[0],
                      part.coords()[1], part.coords()[2], n_local, plan.neighbors.size(), plan.n_send(), plan.shared_dofs.size(), plan.slots.size(), owned,
                      part.global_n_dofs());
        }
    }
  catch (std::exception &exc)
    {
      std::fprintf(stderr, "Exception: %s\n", exc.what());
      return 1;
    }
  return 0;
}
