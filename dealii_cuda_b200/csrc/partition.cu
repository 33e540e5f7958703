// partition.cu -- box partition of the uniform mesh over the GPUs of one node and the interface-DoF exchange plan
// (host code; new capability: the reference is single-GPU, GpuVector::compress is a no-op and locally_owned_elements() is
// the complete index set, gpu_vec.h:174-175).
//
// Layout: ranks ordered x fastest on the grid 1 -> 1x1x1, 2 -> 1x1x2, 4 -> 1x2x2, 8 -> 2x2x2 (2D: 2 -> 1x2, 4 -> 2x2).
// Weak scaling: every rank owns a 2^r cube of cells (the domain grows with the ranks); strong scaling: the refine_global(r)
// cube [left,right]^dim is cut into the rank grid.  Every rank numbers its box like a standalone mesh and stores all DoFs its
// cells touch; DoFs on partition interfaces are replicated.  For every neighbour (faces, edges, vertices of the box) the
// shared lattice points are listed in lexicographic order (x fastest) -- the same order on both sides, so the k-th entry a
// rank sends to a neighbour is the k-th entry that neighbour expects from it.  Points on the global Dirichlet boundary are
// constrained on every replica (dst[c] = src[c]) and take no part in the exchange.
#include <algorithm>
#include <map>
#include <memory>
#include "common.cuh"

using namespace mfg;

struct mfg_partition_plan
{
  int rank = 0, world = 1;
  std::vector<int>      neighbors;   // ascending ranks that share unconstrained DoFs with this rank
  std::vector<uint32_t> splits;      // [world] entries sent to / received from every rank
  std::vector<uint32_t> recv_off;    // [world] where a neighbour's block starts in this rank's receive buffer
  std::vector<uint32_t> pack_idx;    // [n_send] local DoF of every send-buffer entry
  std::vector<uint32_t> shared_dofs; // ascending local DoFs that receive contributions
  std::vector<uint32_t> offsets;     // [n_shared+1] CSR over the contributions of a shared DoF, ascending rank order
  std::vector<int32_t>  slots;       // receive-buffer index of a contribution, -1 = this rank's own partial sum
  std::vector<uint8_t>  owned;       // [n_local] 1 = this rank is the lowest rank touching the DoF
};

namespace {

void rank_grid(int world, int dim, int g[3])
{
  MFG_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  g[0] = g[1] = g[2] = 1;
  if (dim == 3)
    {
      MFG_REQUIRE(world == 1 || world == 2 || world == 4 || world == 8, "3D: 1, 2, 4 or 8 ranks");
      if (world == 2) g[2] = 2;
      else if (world == 4) g[1] = g[2] = 2;
      else if (world == 8) g[0] = g[1] = g[2] = 2;
    }
  else
    {
      MFG_REQUIRE(world == 1 || world == 2 || world == 4, "2D: 1, 2 or 4 ranks");
      if (world == 2) g[1] = 2;
      else if (world == 4) g[0] = g[1] = 2;
    }
}

void rank_coords(int rank, int world, int dim, int me[3], int g[3])
{
  rank_grid(world, dim, g);
  MFG_REQUIRE(rank >= 0 && rank < world, "rank out of range");
  me[0] = rank % g[0]; me[1] = (rank / g[0]) % g[1]; me[2] = rank / (g[0] * g[1]);
}

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

void local_log2(int world, int dim, int r, bool strong, int lg[3])
{
  int g[3];
  rank_grid(world, dim, g);
  lg[0] = lg[1] = lg[2] = 0;
  for (int d = 0; d < dim; ++d)
    {
      lg[d] = r - (strong ? ilog2(g[d]) : 0);
      MFG_REQUIRE(lg[d] >= 0, "more ranks than cells in a direction");
    }
}

}  // namespace

extern "C" {

int mfg_partition_rank_coords(int rank, int world, int dim, int coords[3], int grid[3])
{
  return guarded([&] { MFG_REQUIRE(coords && grid, "null argument"); rank_coords(rank, world, dim, coords, grid); });
}

int mfg_partition_box(int rank, int world, int dim, int degree, int r, double left, double right, int strong, mfg_box_desc *out)
{
  return guarded([&] {
    MFG_REQUIRE(out, "null argument");
    int me[3], g[3], lg[3];
    rank_coords(rank, world, dim, me, g);
    local_log2(world, dim, r, strong != 0, lg);
    std::memset(out, 0, sizeof(*out));
    out->dim = dim; out->degree = degree;
    out->h = (right - left) / (double)(1 << r);
    uint32_t faces = 0;
    for (int d = 0; d < dim; ++d)
      {
        if (me[d] == 0) faces |= 1u << (2 * d);
        if (me[d] == g[d] - 1) faces |= 1u << (2 * d + 1);
        out->log2_cells[d] = lg[d];
        out->origin[d] = left + me[d] * out->h * (double)(1 << lg[d]);
      }
    out->dirichlet_faces = faces;
  });
}

int mfg_partition_global_n_dofs(int world, int dim, int degree, int r, int strong, uint64_t *out)
{
  return guarded([&] {
    MFG_REQUIRE(out, "null argument");
    int g[3], lg[3];
    rank_grid(world, dim, g);
    local_log2(world, dim, r, strong != 0, lg);
    uint64_t n = 1;
    for (int d = 0; d < dim; ++d) n *= (uint64_t)degree * ((uint64_t)1 << lg[d]) * g[d] + 1;
    *out = n;
  });
}

// Shared lattice points of `rank` with the neighbour at grid offset delta (each in -1, 0, 1): local lattice coordinates
// (0 .. degree * 2^lg_d per direction), lexicographic, x fastest.  drop_dirichlet: leave out the points on the global boundary.
// xyz == NULL: only count.  *neighbor_rank = -1 (and *n_points = 0) if there is no rank at that offset.
int mfg_partition_interface_points(int rank, int world, int dim, int degree, int r, int strong, const int delta[3], int drop_dirichlet,
                                   int *neighbor_rank, size_t *n_points, uint32_t *xyz)
{
  return guarded([&] {
    MFG_REQUIRE(delta && neighbor_rank && n_points, "null argument");
    MFG_REQUIRE(degree >= 1, "degree must be positive");
    int me[3], g[3], lg[3];
    rank_coords(rank, world, dim, me, g);
    local_log2(world, dim, r, strong != 0, lg);
    *neighbor_rank = -1; *n_points = 0;
    bool any = false;
    int  nb[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d)
      {
        const int dl = d < dim ? delta[d] : 0;
        MFG_REQUIRE(dl >= -1 && dl <= 1, "delta entries must be -1, 0 or 1");
        any = any || dl != 0;
        nb[d] = me[d] + dl;
        if (nb[d] < 0 || nb[d] >= g[d]) return;
      }
    if (!any) return;
    uint32_t lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};  // [lo, hi) per direction
    for (int d = 0; d < dim; ++d)
      {
        const uint32_t M = (uint32_t)degree << lg[d];
        if (delta[d] == 1) { lo[d] = M; hi[d] = M + 1; }
        else if (delta[d] == -1) { lo[d] = 0; hi[d] = 1; }
        else
          {
            lo[d] = (drop_dirichlet && me[d] == 0) ? 1 : 0;
            hi[d] = (drop_dirichlet && me[d] == g[d] - 1) ? M : M + 1;
            if (hi[d] <= lo[d]) { hi[d] = lo[d]; }
          }
      }
    *neighbor_rank = nb[0] + g[0] * (nb[1] + g[1] * nb[2]);
    size_t cnt = 1;
    for (int d = 0; d < 3; ++d) cnt *= hi[d] - lo[d];
    *n_points = cnt;
    if (!xyz || cnt == 0) return;
    size_t k = 0;
    for (uint32_t z = lo[2]; z < hi[2]; ++z)
      for (uint32_t y = lo[1]; y < hi[1]; ++y)
        for (uint32_t x = lo[0]; x < hi[0]; ++x) { xyz[3 * k] = x; xyz[3 * k + 1] = y; xyz[3 * k + 2] = z; ++k; }
  });
}

// The exchange plan of one rank from the DoF lists it shares with its neighbours:
//   n_lists lists; list i = dofs[list_start[i] .. list_start[i+1]) shared with rank list_rank[i] (exchanged, Dirichlet points
//   dropped); replicated lists likewise (all shared points: they decide ownership only).
int mfg_partition_plan_create(int rank, int world, uint32_t n_local, int n_lists, const int *list_rank, const size_t *list_start, const uint32_t *dofs,
                              int n_repl, const int *repl_rank, const size_t *repl_start, const uint32_t *repl_dofs, mfg_partition_plan **out)
{
  return guarded([&] {
    MFG_REQUIRE(out && rank >= 0 && rank < world, "bad argument");
    MFG_REQUIRE(n_lists == 0 || (list_rank && list_start && dofs), "null list argument");
    MFG_REQUIRE(n_repl == 0 || (repl_rank && repl_start && repl_dofs), "null list argument");
    std::unique_ptr<mfg_partition_plan> p(new mfg_partition_plan);
    p->rank = rank; p->world = world;
    std::map<int, std::pair<size_t, size_t>> lists;  // neighbour rank -> [begin, end) in dofs (ascending ranks)
    for (int i = 0; i < n_lists; ++i)
      {
        MFG_REQUIRE(list_rank[i] >= 0 && list_rank[i] < world && list_rank[i] != rank, "bad neighbour rank");
        MFG_REQUIRE(!lists.count(list_rank[i]), "two lists for one neighbour");
        lists[list_rank[i]] = {list_start[i], list_start[i + 1]};
      }
    p->splits.assign(world, 0); p->recv_off.assign(world, 0);
    uint32_t o = 0;
    for (auto &kv : lists)
      {
        p->neighbors.push_back(kv.first);
        const size_t cnt = kv.second.second - kv.second.first;
        p->splits[kv.first] = (uint32_t)cnt;
        p->recv_off[kv.first] = o;
        for (size_t k = kv.second.first; k < kv.second.second; ++k)
          {
            MFG_REQUIRE(dofs[k] < n_local, "shared DoF out of range");
            p->pack_idx.push_back(dofs[k]);
          }
        o += (uint32_t)cnt;
      }
    // contributions per shared DoF: (rank, slot) in ascending rank order, the own partial sum as (rank, -1)
    std::vector<std::pair<uint32_t, std::pair<int, int32_t>>> contrib;  // (dof, (rank, slot))
    contrib.reserve(p->pack_idx.size() * 2);
    for (auto &kv : lists)
      for (size_t k = kv.second.first; k < kv.second.second; ++k)
        contrib.push_back({dofs[k], {kv.first, (int32_t)(p->recv_off[kv.first] + (k - kv.second.first))}});
    std::sort(contrib.begin(), contrib.end());
    p->owned.assign(n_local, 1);
    p->offsets.push_back(0);
    for (size_t i = 0; i < contrib.size();)
      {
        const uint32_t d = contrib[i].first;
        size_t         j = i;
        bool           own_done = false;
        if (contrib[i].second.first < rank) p->owned[d] = 0;  // owner = lowest rank touching the DoF
        for (; j < contrib.size() && contrib[j].first == d; ++j)
          {
            if (!own_done && contrib[j].second.first > rank) { p->slots.push_back(-1); own_done = true; }
            p->slots.push_back(contrib[j].second.second);
          }
        if (!own_done) p->slots.push_back(-1);
        p->shared_dofs.push_back(d);
        p->offsets.push_back((uint32_t)p->slots.size());
        i = j;
      }
    // constrained interface DoFs are replicated too but take no part in the exchange
    for (int i = 0; i < n_repl; ++i)
      if (repl_rank[i] < rank)
        for (size_t k = repl_start[i]; k < repl_start[i + 1]; ++k)
          {
            MFG_REQUIRE(repl_dofs[k] < n_local, "replicated DoF out of range");
            p->owned[repl_dofs[k]] = 0;
          }
    *out = p.release();
  });
}
int mfg_partition_plan_destroy(mfg_partition_plan *p) { return guarded([&] { delete p; }); }
// sizes: out[0] = n_send, [1] = n_shared, [2] = n_slots, [3] = n_neighbors, [4] = n_local
int mfg_partition_plan_sizes(const mfg_partition_plan *p, size_t out[5])
{
  return guarded([&] {
    MFG_REQUIRE(p && out, "null argument");
    out[0] = p->pack_idx.size(); out[1] = p->shared_dofs.size(); out[2] = p->slots.size(); out[3] = p->neighbors.size(); out[4] = p->owned.size();
  });
}
int mfg_partition_plan_get(const mfg_partition_plan *p, int *neighbors, uint32_t *splits, uint32_t *recv_off, uint32_t *pack_idx, uint32_t *shared_dofs,
                           uint32_t *offsets, int32_t *slots, uint8_t *owned_mask)
{
  return guarded([&] {
    MFG_REQUIRE(p, "null argument");
    if (neighbors) std::copy(p->neighbors.begin(), p->neighbors.end(), neighbors);
    if (splits) std::copy(p->splits.begin(), p->splits.end(), splits);
    if (recv_off) std::copy(p->recv_off.begin(), p->recv_off.end(), recv_off);
    if (pack_idx) std::copy(p->pack_idx.begin(), p->pack_idx.end(), pack_idx);
    if (shared_dofs) std::copy(p->shared_dofs.begin(), p->shared_dofs.end(), shared_dofs);
    if (offsets) std::copy(p->offsets.begin(), p->offsets.end(), offsets);
    if (slots) std::copy(p->slots.begin(), p->slots.end(), slots);
    if (owned_mask) std::copy(p->owned.begin(), p->owned.end(), owned_mask);
  });
}

}  // extern "C"
