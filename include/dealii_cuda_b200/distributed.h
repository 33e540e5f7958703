// distributed.h -- C++ facade of the multi-GPU pieces of the library (new capability: the reference is single-GPU, gpu_vec.h:174-175):
//   BoxPartition      rank grid, local box, interface lattice points            (mfg_partition_*, host code)
//   ExchangePlan      send lists, ordered-sum CSR, owned-DoF mask of one rank   (mfg_partition_plan_*, host code)
//   InterfaceExchange pack / push / ordered accumulate kernels on that plan     (mfg_exchange_*, device)
// The transport between the processes (NCCL send/recv, CUDA IPC or symmetric-memory peer pointers, the device-side barrier)
// is the caller's: dealii_cuda_b200/distributed.py does it with torch.distributed; an MPI + NCCL host would do the same calls.
#pragma once
#include <array>
#include <map>
#include <vector>
#include "matrix_free_gpu.h"

namespace dealii_cuda_b200 {

class BoxPartition
{
public:
  // weak scaling (strong = false): a 2^r cube of cells per rank; strong: the refine_global(r) cube [left,right]^dim cut into the grid
  BoxPartition(int rank, int world, int dim, int degree, int r, bool strong = false, double left = -1., double right = 1.)
    : rank_(rank), world_(world), dim_(dim), degree_(degree), r_(r), strong_(strong)
  {
    check(mfg_partition_rank_coords(rank, world, dim, me_.data(), grid_.data()));
    check(mfg_partition_box(rank, world, dim, degree, r, left, right, strong ? 1 : 0, &box_));
  }
  const mfg_box_desc       &box() const { return box_; }          // feed to mfg_mesh_create_box
  const std::array<int, 3> &coords() const { return me_; }
  const std::array<int, 3> &grid() const { return grid_; }
  unsigned long long        global_n_dofs() const
  {
    uint64_t n = 0;
    check(mfg_partition_global_n_dofs(world_, dim_, degree_, r_, strong_ ? 1 : 0, &n));
    return n;
  }
  // lattice points shared with the neighbour at grid offset delta (x, y, z in -1..1): (neighbour rank or -1, [m][3] points)
  int interface_points(const std::array<int, 3> &delta, bool drop_dirichlet, std::vector<uint32_t> &xyz) const
  {
    int    nb = -1;
    size_t cnt = 0;
    check(mfg_partition_interface_points(rank_, world_, dim_, degree_, r_, strong_ ? 1 : 0, delta.data(), drop_dirichlet ? 1 : 0, &nb, &cnt, nullptr));
    xyz.assign(3 * cnt, 0u);
    if (cnt) check(mfg_partition_interface_points(rank_, world_, dim_, degree_, r_, strong_ ? 1 : 0, delta.data(), drop_dirichlet ? 1 : 0, &nb, &cnt, xyz.data()));
    return nb;
  }
  int rank() const { return rank_; }
  int world() const { return world_; }
  int dim() const { return dim_; }

private:
  int                rank_, world_, dim_, degree_, r_;
  bool               strong_;
  std::array<int, 3> me_{}, grid_{};
  mfg_box_desc       box_{};
};

class ExchangePlan
{
public:
  std::vector<int>      neighbors;
  std::vector<uint32_t> splits, recv_off, pack_idx, shared_dofs, offsets;
  std::vector<int32_t>  slots;
  std::vector<uint8_t>  owned_mask;

  // lattice_to_dof: callable (const std::vector<uint32_t> &xyz) -> std::vector<uint32_t> of local DoFs (mfg_mesh_lattice_to_dof)
  template <typename LatticeToDof> ExchangePlan(const BoxPartition &part, uint32_t n_local, LatticeToDof &&lattice_to_dof)
  {
    std::map<int, std::vector<uint32_t>> lists, repl;
    std::vector<uint32_t>                pts;
    const int                            dim = part.dim();
    for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx)
          {
            if (!dx && !dy && !dz) continue;
            const std::array<int, 3> delta{dx, dy, dz};
            const int                nb = part.interface_points(delta, false, pts);
            if (nb < 0) continue;
            repl[nb] = lattice_to_dof(pts);
            part.interface_points(delta, true, pts);
            if (!pts.empty()) lists[nb] = lattice_to_dof(pts);
          }
    auto flatten = [](const std::map<int, std::vector<uint32_t>> &m, std::vector<int> &ranks, std::vector<size_t> &start, std::vector<uint32_t> &flat) {
      start.assign(1, 0);
      for (const auto &kv : m)
        {
          ranks.push_back(kv.first);
          flat.insert(flat.end(), kv.second.begin(), kv.second.end());
          start.push_back(flat.size());
        }
    };
    std::vector<int>      lr, rr;
    std::vector<size_t>   ls, rs;
    std::vector<uint32_t> lf, rf;
    flatten(lists, lr, ls, lf);
    flatten(repl, rr, rs, rf);
    mfg_partition_plan *p = nullptr;
    check(mfg_partition_plan_create(part.rank(), part.world(), n_local, (int)lr.size(), lr.data(), ls.data(), lf.data(), (int)rr.size(), rr.data(), rs.data(),
                                    rf.data(), &p));
    size_t sz[5];
    int    rc = mfg_partition_plan_sizes(p, sz);
    if (rc == MFG_OK)
      {
        neighbors.resize(sz[3]); splits.resize(part.world()); recv_off.resize(part.world()); pack_idx.resize(sz[0]); shared_dofs.resize(sz[1]);
        offsets.resize(sz[1] + 1); slots.resize(sz[2]); owned_mask.resize(sz[4]);
        rc = mfg_partition_plan_get(p, neighbors.data(), splits.data(), recv_off.data(), pack_idx.data(), shared_dofs.data(), offsets.data(), slots.data(),
                                    owned_mask.data());
      }
    mfg_partition_plan_destroy(p);
    check(rc);
  }
  size_t n_send() const { return pack_idx.size(); }
};

// the pack / push / accumulate kernels on a plan (device object)
template <typename Number> class InterfaceExchange
{
public:
  explicit InterfaceExchange(const ExchangePlan &plan)
  {
    check(mfg_exchange_create(default_context(), dtype_of<Number>(), plan.pack_idx.data(), plan.pack_idx.size(), plan.shared_dofs.data(), plan.shared_dofs.size(),
                              plan.offsets.data(), plan.slots.data(), plan.slots.size(), &ex_));
  }
  ~InterfaceExchange() { if (ex_) mfg_exchange_destroy(ex_); }
  InterfaceExchange(const InterfaceExchange &) = delete;
  void pack(const GpuVector<Number> &vec, Number *send_dev) const { check(mfg_exchange_pack(ex_, vec.getDataRO(), send_dev)); }                 // send[k] = vec[pack_idx[k]]
  void accumulate(GpuVector<Number> &vec, const Number *recv_dev) const { check(mfg_exchange_accumulate(ex_, vec.getData(), recv_dev)); }        // ordered sum
  void push(const GpuVector<Number> &vec, const std::vector<uint64_t> &peer_dst, const std::vector<uint32_t> &chunk_start, void *cuda_stream) const
  {
    check(mfg_exchange_push_stream(ex_, vec.getDataRO(), peer_dst.data(), chunk_start.data(), (int)peer_dst.size(), cuda_stream));
  }
  mfg_exchange *handle() const { return ex_; }

private:
  mfg_exchange *ex_ = nullptr;
};

}  // namespace dealii_cuda_b200
