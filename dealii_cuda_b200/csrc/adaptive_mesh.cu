// adaptive_mesh.cu -- host substrate for adaptively refined meshes with hanging nodes (BASELINE configs[3]).
//
// What the reference takes from deal.II on such meshes, restated without deal.II:
//   * Triangulation<dim> on hyper_cube(left, right): refine_global, set_refine_flag + execute_coarsening_and_refinement
//     (flags closed under deal.II's rule that neighbouring cells differ by at most one level -- across faces, in 3D also
//     across edges, NOT across vertices), cells stored per level in creation order, so that the active cells come out in
//     deal.II's order (level by level, children behind the cells that existed before);
//   * the reference's flagging helpers mark_cells_in_annulus / mark_cells_on_shell / pseudo_adaptive_refinement
//     (bmop_common.h:9-105) and octant_criterion (poisson_common.h:29-35);
//   * DoFHandler::distribute_dofs for FE_Q(p): first touch over the active cells, hierarchic order inside a cell, hanging
//     faces / edges carry their own DoFs;
//   * HangingNodes::setup_constraints (matrix_free_gpu/hanging_nodes.cuh:209-454): per active cell the 9-bit mask (:38-50)
//     and the rewrite of loc2glob (constrained faces / edges point at the coarse neighbour's DoFs);
//   * the constraint list of ConstraintHandlerGpu (constraint_handler_gpu.cu:77-83): hanging DoFs and Dirichlet boundary
//     DoFs, ascending; J^-1 per cell (uniform-mesh form, matrix_free_gpu.cu:332-334) and Coefficient::value at the Gauss
//     points (poisson_common.h:146-158).
// Everything here is host code; mfg_laplace_create_from_amesh hands the arrays to the same device path a deal.II based
// caller uses (mfg_mf_reinit with constraint_mask, mfg_ch_create, mfg_laplace_create_from_arrays).
// The checker is oracle/adaptive.py (tests/test_adaptive_mesh.py compares every array bit for bit).
#include <algorithm>
#include <cmath>
#include <memory>
#include <unordered_map>
#include "mesh.cuh"
#include "operators.cuh"

using namespace mfg;

namespace {

constexpr uint32_t CONSTR_TYPE[3] = {1u << 0, 1u << 1, 1u << 2};   // hanging_nodes.cuh:38-40
constexpr uint32_t CONSTR_FACE[3] = {1u << 3, 1u << 4, 1u << 5};   // :43-45
// :48-50 name an edge bit by its two normal directions (XY, YZ, ZX = bits 6, 7, 8); indexed here by the direction the
// edge runs along: x -> YZ, y -> ZX, z -> XY
constexpr uint32_t EDGE_BIT_ALONG[3] = {1u << 7, 1u << 8, 1u << 6};

struct ACell
{
  uint32_t x[3];
  int32_t  child0;  // index of the first child on the next level, -1: active
  uint8_t  flag;
};

inline uint64_t pack(const uint32_t x[3]) { return (uint64_t)x[0] | ((uint64_t)x[1] << 21) | ((uint64_t)x[2] << 42); }

struct EntityKey
{
  uint32_t b[3];    // lower corner in units of the finest active level
  uint32_t shape;   // free-direction mask | size << 3   (vertices: 0)
  bool     operator==(const EntityKey &o) const { return b[0] == o.b[0] && b[1] == o.b[1] && b[2] == o.b[2] && shape == o.shape; }
};
struct EntityHash
{
  size_t operator()(const EntityKey &k) const
  {
    uint64_t h = k.b[0] * 0x9E3779B97F4A7C15ull;
    h ^= (k.b[1] + 0x7F4A7C15u) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.b[2] + 0x165667B1u) * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= (k.shape + 0x27D4EB2Fu) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

}  // namespace

struct mfg_amesh
{
  int    dim = 0, p = 0, n = 0;
  double left = 0, right = 0;
  bool   limit_level_difference_at_vertices = false;  // Triangulation::MeshSmoothing flag of the MG drivers (poisson_mg.cu:132)
  std::vector<std::vector<ACell>>                        levels;
  std::vector<std::unordered_map<uint64_t, uint32_t>>    index;  // per level: packed coordinates -> position in the level
  // active cells in deal.II's iteration order (rebuilt after every refinement)
  std::vector<uint32_t> act_level, act_pos;
  // after distribute_dofs
  bool                  dofs_ready = false;
  uint32_t              n_dofs = 0, npc = 0, lmax = 0;
  std::vector<uint32_t> l2g_own, l2g, mask, hanging, boundary, constrained;
  std::vector<double>   inv_jac;
  FEData1D              fe;
  // multigrid hierarchy (build_mg): level meshes = ALL cells of a level, level DoFs, MGConstrainedDoFs sets, transfer blocks
  struct MgLevel
  {
    uint32_t              n_cells = 0, n_dofs = 0, n_parents = 0;
    std::vector<uint32_t> l2g;                     // [n_cells][npc] lexicographic, level numbering
    std::vector<uint32_t> boundary, edge;          // ascending: boundary_indices / refinement_edge_indices of the level
    std::vector<uint32_t> coarse_idx, fine_idx;    // transfer INTO this level: [n_parents][npc] (bit 31: boundary DoF of level-1), [n_parents][(2p+1)^dim]
    std::vector<double>   weights;                 // [n_parents][3^dim]: 1 / number of refined parents that hold the fine DoF
    std::vector<uint32_t> copy_global, copy_level; // copy_to_mg / copy_from_mg pairs (active DoF, level DoF)
  };
  bool                 mg_ready = false;
  int                  mg_min_level = 0;
  std::vector<MgLevel> mg;                         // [level - mg_min_level]

  uint32_t n_active() const { return (uint32_t)act_level.size(); }
  ACell       &cell(uint32_t a) { return levels[act_level[a]][act_pos[a]]; }
  const ACell &cell(uint32_t a) const { return levels[act_level[a]][act_pos[a]]; }

  void rebuild_active()
  {
    act_level.clear(); act_pos.clear();
    for (uint32_t l = 0; l < levels.size(); ++l)
      for (uint32_t i = 0; i < levels[l].size(); ++i)
        if (levels[l][i].child0 < 0) { act_level.push_back(l); act_pos.push_back(i); }
    dofs_ready = false; mg_ready = false;
  }

  // the active cell covering position xs of level `level`: 1 = found (level / position returned), 0 = outside the domain,
  // -1 = the place is covered by finer cells
  int find_active(uint32_t level, const int64_t xs[3], uint32_t &fl, uint32_t &fp) const
  {
    for (int d = 0; d < dim; ++d)
      if (xs[d] < 0 || xs[d] >= ((int64_t)1 << level)) return 0;
    for (int l = (int)level; l >= 0; --l)
      {
        uint32_t c[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) c[d] = (uint32_t)(xs[d] >> (level - l));
        auto it = index[l].find(pack(c));
        if (it == index[l].end()) continue;
        if (levels[l][it->second].child0 >= 0) return -1;
        fl = (uint32_t)l; fp = it->second;
        return 1;
      }
    return -1;
  }

  void refine_cell(uint32_t l, uint32_t i)
  {
    if (levels.size() <= l + 1) { levels.emplace_back(); index.emplace_back(); }
    MFG_REQUIRE(l + 1 <= 20, "more than 20 refinement levels");
    const uint32_t first = (uint32_t)levels[l + 1].size();
    ACell          parent = levels[l][i];
    for (uint32_t k = 0; k < (1u << dim); ++k)  // deal.II child order: x fastest
      {
        ACell ch;
        for (int d = 0; d < 3; ++d) ch.x[d] = d < dim ? 2 * parent.x[d] + ((k >> d) & 1u) : 0u;
        ch.child0 = -1; ch.flag = 0;
        index[l + 1][pack(ch.x)] = (uint32_t)levels[l + 1].size();
        levels[l + 1].push_back(ch);
      }
    levels[l][i].child0 = (int32_t)first;
    levels[l][i].flag = 0;
  }

  // Triangulation::execute_coarsening_and_refinement (refinement only): close the flags under the one-level rule, then refine
  // level by level in storage order
  void execute_refinement()
  {
    std::vector<std::pair<uint32_t, uint32_t>> stack;
    for (uint32_t l = 0; l < levels.size(); ++l)
      for (uint32_t i = 0; i < levels[l].size(); ++i)
        if (levels[l][i].child0 < 0 && levels[l][i].flag) stack.emplace_back(l, i);
    const int max_nonzero = limit_level_difference_at_vertices ? dim : (dim == 3 ? 2 : 1);  // faces; in 3D also edges; with the flag also vertices
    while (!stack.empty())
      {
        const auto [l, i] = stack.back();
        stack.pop_back();
        const ACell c = levels[l][i];
        int delta[3];
        for (delta[2] = (dim == 3 ? -1 : 0); delta[2] <= (dim == 3 ? 1 : 0); ++delta[2])
          for (delta[1] = -1; delta[1] <= 1; ++delta[1])
            for (delta[0] = -1; delta[0] <= 1; ++delta[0])
              {
                const int nz = (delta[0] != 0) + (delta[1] != 0) + (delta[2] != 0);
                if (nz == 0 || nz > max_nonzero) continue;
                int64_t xs[3] = {0, 0, 0};
                for (int d = 0; d < dim; ++d) xs[d] = (int64_t)c.x[d] + delta[d];
                uint32_t fl, fp;
                if (find_active(l, xs, fl, fp) == 1 && fl < l && !levels[fl][fp].flag)
                  {
                    levels[fl][fp].flag = 1;
                    stack.emplace_back(fl, fp);
                  }
              }
      }
    const uint32_t n_levels = (uint32_t)levels.size();
    for (uint32_t l = 0; l < n_levels; ++l)
      {
        const uint32_t cnt = (uint32_t)levels[l].size();  // (children created in this pass are never flagged)
        for (uint32_t i = 0; i < cnt; ++i)
          if (levels[l][i].child0 < 0 && levels[l][i].flag) refine_cell(l, i);
      }
    rebuild_active();
  }

  double cell_h(uint32_t level) const { return (right - left) / (double)((uint64_t)1 << level); }
};

namespace {

// dealii::Point::distance: sqrt of the sum of squares, d = 0 .. dim-1
inline double distance(const double *a, const double *b, int dim)
{
  double s = 0;
  for (int d = 0; d < dim; ++d) s += (a[d] - b[d]) * (a[d] - b[d]);
  return std::sqrt(s);
}

void mark_in_annulus(mfg_amesh *am, double R, double r, const double *center)
{
  for (uint32_t a = 0; a < am->n_active(); ++a)
    {
      ACell       &c = am->cell(a);
      const double h = am->cell_h(am->act_level[a]);
      double       ctr[3] = {0, 0, 0};
      for (int d = 0; d < am->dim; ++d) ctr[d] = am->left + h * ((double)c.x[d] + 0.5);
      const double dist = distance(ctr, center, am->dim);
      if (dist > r && dist < R) c.flag = 1;
    }
}

void mark_on_shell(mfg_amesh *am, double R, const double *center)
{
  const int nverts = 1 << am->dim;
  for (uint32_t a = 0; a < am->n_active(); ++a)
    {
      ACell       &c = am->cell(a);
      const double h = am->cell_h(am->act_level[a]);
      int          ninside = 0;
      for (int v = 0; v < nverts; ++v)
        {
          double x[3] = {0, 0, 0};
          for (int d = 0; d < am->dim; ++d) x[d] = am->left + h * (double)(c.x[d] + ((v >> d) & 1));
          ninside += distance(x, center, am->dim) < R;
        }
      if (ninside != 0 && ninside != nverts) c.flag = 1;
    }
}

void refine_global(mfg_amesh *am, int times)
{
  for (int t = 0; t < times; ++t)
    {
      for (uint32_t a = 0; a < am->n_active(); ++a) am->cell(a).flag = 1;
      am->execute_refinement();
    }
}

// bmop_common.h:49-105 for domain == CUBE (the float `reduction` and the double radii as written there)
void pseudo_adaptive_refinement(mfg_amesh *am, int n_ref)
{
  n_ref = std::max(n_ref - 2, 0);
  refine_global(am, n_ref);
  const float  reduction = am->dim == 2 ? 0.005f : 0.015f;
  const double zero[3] = {0, 0, 0};
  mark_in_annulus(am, 0.55 - reduction, 0.0, zero);
  am->execute_refinement();
  mark_in_annulus(am, 0.42 - reduction, 0.3 + reduction, zero);
  am->execute_refinement();
  mark_in_annulus(am, 0.41 - reduction, 0.32 + reduction, zero);
  am->execute_refinement();
  double offset[3] = {0, 0, 0};
  for (int d = 0; d < am->dim; ++d) offset[d] = -0.1 * (d + 1);
  mark_in_annulus(am, 0.33 - reduction, 0.17 + reduction, offset);
  am->execute_refinement();
  mark_in_annulus(am, 0.31 - reduction, 0.21 + reduction, offset);
  am->execute_refinement();
  if (am->dim == 2)
    for (int s = 0; s < 4; ++s)
      {
        mark_on_shell(am, 0.25, offset);
        am->execute_refinement();
      }
}

// DoFHandler::distribute_dofs / distribute_mg_dofs for FE_Q(p): first touch over the given cells in order, hierarchic order
// inside a cell.  The DoFs of one mesh entity (vertex, line, quad, hex) are consecutive in the hierarchic order and every
// cell meets them in the same order, so an entity gets a block of numbers when its first cell arrives.  sizes[c] = edge length
// of cell c in the units of the entity keys, S = edge length of the domain in those units.  Returns the number of DoFs.
uint32_t number_cells(int dim, int p, const std::vector<const ACell *> &cells, const std::vector<uint32_t> &sizes, uint32_t S,
                      std::vector<uint32_t> &l2g, std::vector<uint8_t> &on_boundary)
{
  const int      n = p + 1;
  const uint32_t npc = ipow(n, dim), nc = (uint32_t)cells.size();
  const std::vector<uint32_t> h2l = hierarchic_to_lexicographic(dim, p);
  std::unordered_map<EntityKey, uint32_t, EntityHash> first_dof;
  first_dof.reserve((size_t)nc * (dim == 3 ? 8 : 4));
  l2g.assign((size_t)nc * npc, 0);
  on_boundary.clear();
  uint32_t nxt = 0;
  for (uint32_t a = 0; a < nc; ++a)
    {
      const ACell   &c = *cells[a];
      const uint32_t s = sizes[a];
      EntityKey      prev{{0, 0, 0}, 0xffffffffu};
      uint32_t       base_no = 0, rank = 0;
      for (uint32_t hI = 0; hI < npc; ++hI)
        {
          const uint32_t li = h2l[hI];
          uint32_t       idx[3] = {0, 0, 0}, t = li;
          for (int d = 0; d < dim; ++d) { idx[d] = t % n; t /= n; }
          EntityKey key{{0, 0, 0}, 0};
          uint32_t  freemask = 0;
          bool      bnd = false;
          for (int d = 0; d < dim; ++d)
            {
              const uint32_t o = c.x[d] * s;
              if (idx[d] > 0 && idx[d] < (uint32_t)p) { freemask |= 1u << d; key.b[d] = o; }
              else
                {
                  key.b[d] = o + (idx[d] == (uint32_t)p ? s : 0u);
                  bnd = bnd || key.b[d] == 0 || key.b[d] == S;
                }
            }
          key.shape = freemask ? (freemask | (s << 3)) : 0u;
          if (key == prev) ++rank;
          else
            {
              rank = 0; prev = key;
              auto it = first_dof.find(key);
              if (it == first_dof.end())
                {
                  int nfree = 0;
                  for (int d = 0; d < dim; ++d) nfree += (freemask >> d) & 1u;
                  const uint32_t cnt = ipow(p - 1, nfree);
                  base_no = nxt;
                  first_dof.emplace(key, nxt);
                  MFG_REQUIRE((uint64_t)nxt + cnt < 0x80000000ull, "more than 2^31 DoFs");
                  nxt += cnt;
                  on_boundary.resize(nxt, 0);
                }
              else base_no = it->second;
            }
          const uint32_t g = base_no + rank;
          l2g[(size_t)a * npc + li] = g;
          if (bnd) on_boundary[g] = 1;
        }
    }
  return nxt;
}

void distribute_dofs(mfg_amesh *am)
{
  const int      dim = am->dim, p = am->p, n = am->n;
  const uint32_t npc = ipow(n, dim), nact = am->n_active();
  MFG_REQUIRE((uint64_t)nact * npc < (1ull << 32), "too many cells for 32-bit local-to-global offsets");
  am->npc = npc;
  am->lmax = 0;
  for (uint32_t a = 0; a < nact; ++a) am->lmax = std::max(am->lmax, am->act_level[a]);
  const std::vector<uint32_t> h2l = hierarchic_to_lexicographic(dim, p);
  auto size_of = [&](uint32_t a) { return 1u << (am->lmax - am->act_level[a]); };

  std::vector<const ACell *> cells(nact);
  std::vector<uint32_t>      sizes(nact);
  for (uint32_t a = 0; a < nact; ++a) { cells[a] = &am->cell(a); sizes[a] = size_of(a); }
  std::vector<uint8_t> on_boundary;
  am->n_dofs = number_cells(dim, p, cells, sizes, 1u << am->lmax, am->l2g_own, on_boundary);

  // ---- HangingNodes::setup_constraints (hanging_nodes.cuh:209-454): masks and the loc2glob rewrite
  am->l2g = am->l2g_own;
  am->mask.assign(nact, 0);
  // (level, position) -> active index
  std::vector<std::vector<uint32_t>> act_of(am->levels.size());
  for (uint32_t l = 0; l < am->levels.size(); ++l) act_of[l].assign(am->levels[l].size(), 0xffffffffu);
  for (uint32_t a = 0; a < nact; ++a) act_of[am->act_level[a]][am->act_pos[a]] = a;
  auto lat = [&](const uint32_t i[3]) { return i[0] + n * (i[1] + n * i[2]); };
  std::vector<uint32_t> candidates;
  for (uint32_t a = 0; a < nact; ++a)
    {
      const ACell   &c = am->cell(a);
      const uint32_t l = am->act_level[a];
      uint32_t       mask = 0;
      for (int d = 0; d < dim; ++d)
        for (int side = 0; side < 2; ++side)
          {
            int64_t xs[3] = {c.x[0], c.x[1], c.x[2]};
            xs[d] += side == 0 ? -1 : 1;
            uint32_t fl, fp;
            if (am->find_active(l, xs, fl, fp) != 1 || fl >= l) continue;
            // only the outer face of a child can see a coarser neighbour
            MFG_REQUIRE((int)(c.x[d] & 1u) == side, "mesh is not one-irregular");
            mask |= CONSTR_FACE[d];
            const uint32_t nb = act_of[fl][fp];
            const uint32_t nt = ipow(n, dim - 1);
            for (uint32_t t = 0; t < nt; ++t)
              {
                uint32_t mine[3] = {0, 0, 0}, theirs[3] = {0, 0, 0}, tt = t;
                mine[d] = side == 0 ? 0 : p;
                theirs[d] = side == 0 ? p : 0;  // the neighbour's opposite face
                for (int e = 0; e < dim; ++e)
                  if (e != d) { mine[e] = theirs[e] = tt % n; tt /= n; }
                candidates.push_back(am->l2g_own[(size_t)a * npc + lat(mine)]);
                am->l2g[(size_t)a * npc + lat(mine)] = am->l2g_own[(size_t)nb * npc + lat(theirs)];
              }
          }
      if (dim == 3)
        for (int along = 0; along < 3; ++along)  // edges running along `along` that are not part of a constrained face (:371)
          {
            const int a1 = (along + 1) % 3, a2 = (along + 2) % 3;
            if (mask & (CONSTR_FACE[a1] | CONSTR_FACE[a2])) continue;
            const int s1 = c.x[a1] & 1u, s2 = c.x[a2] & 1u;  // the outer edge of this child
            static const int dd[3][2] = {{-1, -1}, {-1, 0}, {0, -1}};
            bool     found = false;
            uint32_t fl = 0, fp = 0;
            for (int k = 0; k < 3 && !found; ++k)
              {
                int64_t xs[3] = {c.x[0], c.x[1], c.x[2]};
                xs[a1] += s1 == 0 ? dd[k][0] : -dd[k][0];
                xs[a2] += s2 == 0 ? dd[k][1] : -dd[k][1];
                uint32_t gl, gp;
                if (am->find_active(l, xs, gl, gp) == 1 && gl < l) { found = true; fl = gl; fp = gp; }
              }
            if (!found) continue;
            mask |= EDGE_BIT_ALONG[along];
            const uint32_t nb = act_of[fl][fp];
            const ACell   &cc = am->levels[fl][fp];
            const uint32_t ss = 1u << (am->lmax - fl), ms = size_of(a);
            for (uint32_t t = 0; t < (uint32_t)n; ++t)
              {
                uint32_t mine[3] = {0, 0, 0}, theirs[3] = {0, 0, 0};
                mine[along] = theirs[along] = t;
                const int ax[2] = {a1, a2}, sd[2] = {s1, s2};
                for (int k = 0; k < 2; ++k)
                  {
                    mine[ax[k]] = sd[k] == 0 ? 0 : p;
                    const uint32_t pos = c.x[ax[k]] * ms + (sd[k] == 0 ? 0u : ms), so = cc.x[ax[k]] * ss;
                    MFG_REQUIRE(pos == so || pos == so + ss, "edge of a fine cell is not an edge of its coarse neighbour");
                    theirs[ax[k]] = pos == so ? 0 : p;
                  }
                candidates.push_back(am->l2g_own[(size_t)a * npc + lat(mine)]);
                am->l2g[(size_t)a * npc + lat(mine)] = am->l2g_own[(size_t)nb * npc + lat(theirs)];
              }
          }
      if (mask)
        for (int d = 0; d < dim; ++d)
          if ((c.x[d] & 1u) == 0) mask |= CONSTR_TYPE[d];
      am->mask[a] = mask;
    }
  // a candidate that some cell still reads through its rewritten map is a real (coarse) DoF
  std::vector<uint8_t> referenced(am->n_dofs, 0);
  for (uint32_t g : am->l2g) referenced[g] = 1;
  std::sort(candidates.begin(), candidates.end());
  candidates.erase(std::unique(candidates.begin(), candidates.end()), candidates.end());
  am->hanging.clear();
  for (uint32_t g : candidates)
    if (!referenced[g]) am->hanging.push_back(g);
  am->boundary.clear();
  for (uint32_t g = 0; g < am->n_dofs; ++g)
    if (on_boundary[g]) am->boundary.push_back(g);
  am->constrained.resize(am->hanging.size() + am->boundary.size());
  std::merge(am->hanging.begin(), am->hanging.end(), am->boundary.begin(), am->boundary.end(), am->constrained.begin());
  am->constrained.erase(std::unique(am->constrained.begin(), am->constrained.end()), am->constrained.end());
  am->inv_jac.resize(nact);
  for (uint32_t a = 0; a < nact; ++a) am->inv_jac[a] = 1.0 / am->cell_h(am->act_level[a]);
  am->dofs_ready = true;
}

// The multigrid hierarchy of the reference's poisson_mg.cu / bmop_mg.cu on an adaptively refined mesh, as deal.II hands it to
// MGTransferMatrixFreeGpu, LaplaceOperatorGpu::reinit(dof_handler, mg_constrained_dofs, level) and ConstraintHandlerGpu::reinit
// (mg_constrained_dofs, level):
//   * DoFHandler::distribute_mg_dofs: first touch over ALL cells of a level in storage order;
//   * MGConstrainedDoFs: boundary_indices(level) and refinement_edge_indices(level) = the DoFs on faces of level cells whose
//     neighbour inside the domain is not refined to that level (MGTools::extract_inner_interface_dofs);
//   * the transfer blocks of internal::MGTransfer::setup_transfer (mg_transfer_matrix_free_gpu.cu:173-257): per refined
//     cell of level l-1 its level DoFs and the (2p+1)^dim level-l DoFs of its children, weights = 1 / multiplicity;
//   * the copy indices of MGTransfer::fill_copy_indices (.cu:109-146): (active DoF, level DoF) on the active cells of a level,
//     without the level's refinement-edge DoFs.
void build_mg(mfg_amesh *am, int min_level)
{
  MFG_REQUIRE(am->dofs_ready, "call mfg_amesh_distribute_dofs first");
  const int      dim = am->dim, p = am->p, n = am->n;
  const uint32_t npc = am->npc, nact = am->n_active();
  uint32_t       min_active = 0xffffffffu;
  for (uint32_t a = 0; a < nact; ++a) min_active = std::min(min_active, am->act_level[a]);
  MFG_REQUIRE(min_level >= 0 && (uint32_t)min_level <= min_active, "min_level must not exceed the coarsest active level");
  const int n_levels = (int)am->levels.size();
  am->mg.clear();
  am->mg.resize(n_levels - min_level);
  am->mg_min_level = min_level;
  auto lat = [&](const uint32_t i[3]) { return i[0] + n * (i[1] + n * i[2]); };
  std::vector<std::vector<uint8_t>> is_edge(n_levels - min_level);
  for (int l = min_level; l < n_levels; ++l)
    {
      mfg_amesh::MgLevel &L = am->mg[l - min_level];
      const auto         &C = am->levels[l];
      L.n_cells = (uint32_t)C.size();
      MFG_REQUIRE((uint64_t)L.n_cells * npc < (1ull << 32), "too many cells on a level");
      std::vector<const ACell *> cells(C.size());
      std::vector<uint32_t>      sizes(C.size(), 1u);
      for (size_t i = 0; i < C.size(); ++i) cells[i] = &C[i];
      std::vector<uint8_t> on_boundary;
      L.n_dofs = number_cells(dim, p, cells, sizes, 1u << l, L.l2g, on_boundary);
      for (uint32_t g = 0; g < L.n_dofs; ++g)
        if (on_boundary[g]) L.boundary.push_back(g);
      std::vector<uint8_t> &E = is_edge[l - min_level];
      E.assign(L.n_dofs, 0);
      for (size_t ci = 0; ci < C.size(); ++ci)
        for (int d = 0; d < dim; ++d)
          for (int side = 0; side < 2; ++side)
            {
              int64_t nb = (int64_t)C[ci].x[d] + (side == 0 ? -1 : 1);
              if (nb < 0 || nb >= ((int64_t)1 << l)) continue;  // domain boundary
              uint32_t x[3] = {C[ci].x[0], C[ci].x[1], C[ci].x[2]};
              x[d] = (uint32_t)nb;
              if (am->index[l].count(pack(x))) continue;       // the neighbour is refined to this level as well
              const uint32_t nt = ipow(n, dim - 1);
              for (uint32_t t = 0; t < nt; ++t)
                {
                  uint32_t idx[3] = {0, 0, 0}, tt = t;
                  idx[d] = side == 0 ? 0 : p;
                  for (int e = 0; e < dim; ++e)
                    if (e != d) { idx[e] = tt % n; tt /= n; }
                  E[L.l2g[ci * npc + lat(idx)]] = 1;
                }
            }
      for (uint32_t g = 0; g < L.n_dofs; ++g)
        if (E[g]) L.edge.push_back(g);
    }
  // transfer blocks into level l from level l-1
  const int      nf = 2 * p + 1;
  const uint32_t nF = ipow(nf, dim), n3 = ipow(3, dim);
  for (int l = min_level + 1; l < n_levels; ++l)
    {
      mfg_amesh::MgLevel       &L = am->mg[l - min_level];
      const mfg_amesh::MgLevel &Lc = am->mg[l - 1 - min_level];
      const auto               &P = am->levels[l - 1];
      std::vector<uint8_t>      cb(Lc.n_dofs, 0);
      for (uint32_t g : Lc.boundary) cb[g] = 1;
      std::vector<uint32_t> count(L.n_dofs, 0);
      for (size_t pi = 0; pi < P.size(); ++pi)
        {
          if (P[pi].child0 < 0) continue;
          ++L.n_parents;
          for (uint32_t i = 0; i < npc; ++i)
            {
              const uint32_t g = Lc.l2g[pi * npc + i];
              L.coarse_idx.push_back(g | (cb[g] ? 0x80000000u : 0u));
            }
          for (uint32_t f = 0; f < nF; ++f)
            {
              uint32_t a[3] = {0, 0, 0}, t = f, k = 0, li[3] = {0, 0, 0};
              for (int d = 0; d < dim; ++d)
                {
                  a[d] = t % nf; t /= nf;
                  const uint32_t kd = a[d] > (uint32_t)p ? 1u : 0u;   // the children share the points a_d == p: take the lower one
                  k |= kd << d;
                  li[d] = a[d] - kd * p;
                }
              const uint32_t g = L.l2g[((size_t)P[pi].child0 + k) * npc + lat(li)];
              L.fine_idx.push_back(g);
              ++count[g];
            }
        }
      L.weights.resize((size_t)L.n_parents * n3);
      for (uint32_t q = 0; q < L.n_parents; ++q)
        for (uint32_t r = 0; r < n3; ++r)
          {
            uint32_t t = r, f = 0, stride = 1;
            for (int d = 0; d < dim; ++d)
              {
                const uint32_t rd = t % 3; t /= 3;
                f += (rd == 0 ? 0u : rd == 1 ? 1u : (uint32_t)(2 * p)) * stride;   // a representative point of the region
                stride *= nf;
              }
            L.weights[(size_t)q * n3 + r] = 1.0 / (double)count[L.fine_idx[(size_t)q * nF + f]];
          }
    }
  // copy indices
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> pairs(n_levels - min_level);
  for (uint32_t a = 0; a < nact; ++a)
    {
      const uint32_t            l = am->act_level[a], pos = am->act_pos[a];
      const mfg_amesh::MgLevel &L = am->mg[l - min_level];
      for (uint32_t i = 0; i < npc; ++i)
        {
          const uint32_t lv = L.l2g[(size_t)pos * npc + i];
          if (!is_edge[l - min_level][lv]) pairs[l - min_level].push_back({am->l2g_own[(size_t)a * npc + i], lv});
        }
    }
  for (int l = min_level; l < n_levels; ++l)
    {
      auto &pr = pairs[l - min_level];
      std::sort(pr.begin(), pr.end());
      pr.erase(std::unique(pr.begin(), pr.end()), pr.end());
      mfg_amesh::MgLevel &L = am->mg[l - min_level];
      for (auto &q : pr) { L.copy_global.push_back(q.first); L.copy_level.push_back(q.second); }
    }
  am->mg_ready = true;
}

// inverse Jacobian and Coefficient::value at the Gauss points of ALL cells of a level (the level operator's data)
void level_coefficient(const mfg_amesh *am, int level, double *coef)
{
  const int    dim = am->dim, n = am->n;
  const auto  &C = am->levels[level];
  const double h = am->cell_h(level);
  for (size_t ci = 0; ci < C.size(); ++ci)
    for (uint32_t q = 0; q < am->npc; ++q)
      {
        uint32_t t = q;
        double   r2 = 0;
        for (int d = 0; d < dim; ++d)
          {
            const double x = am->left + h * ((double)C[ci].x[d] + am->fe.qpts[t % n]);
            t /= n;
            r2 += x * x;
          }
        coef[ci * am->npc + q] = 1.0 / (0.05 + 2.0 * r2);
      }
}

// Coefficient::value = 1 / (0.05 + 2 |x|^2) (poisson_common.h:146-158) at the Gauss points of every active cell
void coefficient_at_qpoints(const mfg_amesh *am, double *out, double *qpoints)
{
  const int dim = am->dim, n = am->n;
  for (uint32_t a = 0; a < am->n_active(); ++a)
    {
      const ACell &c = am->cell(a);
      const double h = am->cell_h(am->act_level[a]);
      for (uint32_t q = 0; q < am->npc; ++q)
        {
          uint32_t t = q;
          double   r2 = 0;
          for (int d = 0; d < dim; ++d)
            {
              const double x = am->left + h * ((double)c.x[d] + am->fe.qpts[t % n]);
              t /= n;
              r2 += x * x;
              if (qpoints) qpoints[((size_t)a * am->npc + q) * dim + d] = x;
            }
          if (out) out[(size_t)a * am->npc + q] = 1.0 / (0.05 + 2.0 * r2);
        }
    }
}

}  // namespace

extern "C" {

int mfg_amesh_create(int dim, int degree, double left, double right, mfg_amesh **out)
{
  return guarded([&] {
    MFG_REQUIRE(out, "null argument");
    MFG_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    MFG_REQUIRE(degree >= 1 && degree <= 8, "degree must be in 1..8");
    MFG_REQUIRE(right > left, "empty interval");
    std::unique_ptr<mfg_amesh> am(new mfg_amesh);
    am->dim = dim; am->p = degree; am->n = degree + 1; am->left = left; am->right = right;
    am->fe = make_fe_data(degree);
    am->levels.emplace_back(); am->index.emplace_back();
    ACell root{{0, 0, 0}, -1, 0};
    am->levels[0].push_back(root);
    am->index[0][pack(root.x)] = 0;
    am->rebuild_active();
    *out = am.release();
  });
}
int mfg_amesh_set_limit_level_difference_at_vertices(mfg_amesh *am, int on)
{
  return guarded([&] { MFG_REQUIRE(am, "null argument"); am->limit_level_difference_at_vertices = on != 0; });
}
int mfg_amesh_destroy(mfg_amesh *am) { return guarded([&] { delete am; }); }
int mfg_amesh_refine_global(mfg_amesh *am, int times)
{
  return guarded([&] { MFG_REQUIRE(am && times >= 0, "bad argument"); refine_global(am, times); });
}
int mfg_amesh_set_refine_flags(mfg_amesh *am, const uint8_t *flags, size_t n_flags)
{
  return guarded([&] {
    MFG_REQUIRE(am && flags, "null argument");
    MFG_REQUIRE(n_flags == am->n_active(), "one flag per active cell");
    for (uint32_t a = 0; a < am->n_active(); ++a)
      if (flags[a]) am->cell(a).flag = 1;
  });
}
int mfg_amesh_mark_cells_in_annulus(mfg_amesh *am, double R, double r, const double *center)
{
  return guarded([&] {
    MFG_REQUIRE(am, "null argument");
    const double zero[3] = {0, 0, 0};
    mark_in_annulus(am, R, r, center ? center : zero);
  });
}
int mfg_amesh_mark_cells_on_shell(mfg_amesh *am, double R, const double *center)
{
  return guarded([&] {
    MFG_REQUIRE(am, "null argument");
    const double zero[3] = {0, 0, 0};
    mark_on_shell(am, R, center ? center : zero);
  });
}
int mfg_amesh_mark_octant(mfg_amesh *am)
{
  return guarded([&] {
    MFG_REQUIRE(am, "null argument");
    for (uint32_t a = 0; a < am->n_active(); ++a)
      {
        ACell       &c = am->cell(a);
        const double h = am->cell_h(am->act_level[a]);
        bool         ref = true;
        for (int d = 0; d < am->dim; ++d) ref = ref && (am->left + h * ((double)c.x[d] + 0.5)) > 0.2;
        if (ref) c.flag = 1;
      }
  });
}
int mfg_amesh_execute_refinement(mfg_amesh *am) { return guarded([&] { MFG_REQUIRE(am, "null argument"); am->execute_refinement(); }); }
int mfg_amesh_pseudo_adaptive_refinement(mfg_amesh *am, int n_ref)
{
  return guarded([&] {
    MFG_REQUIRE(am && n_ref >= 0, "bad argument");
    MFG_REQUIRE(am->levels.size() == 1, "pseudo_adaptive_refinement starts from the unrefined hyper_cube");
    pseudo_adaptive_refinement(am, n_ref);
  });
}
int mfg_amesh_info(const mfg_amesh *am, int *dim, int *degree, double *left, double *right, int *n_levels, int *coarsest_active_level)
{
  return guarded([&] {
    MFG_REQUIRE(am, "null argument");
    if (dim) *dim = am->dim;
    if (degree) *degree = am->p;
    if (left) *left = am->left;
    if (right) *right = am->right;
    if (n_levels) *n_levels = (int)am->levels.size();
    if (coarsest_active_level)
      {
        uint32_t m = 0xffffffffu;
        for (uint32_t a = 0; a < am->n_active(); ++a) m = std::min(m, am->act_level[a]);
        *coarsest_active_level = (int)m;
      }
  });
}
uint32_t mfg_amesh_n_active_cells(const mfg_amesh *am) { return am ? am->n_active() : 0; }
uint32_t mfg_amesh_n_levels(const mfg_amesh *am) { return am ? (uint32_t)am->levels.size() : 0; }
int mfg_amesh_get_active_cells(const mfg_amesh *am, uint32_t *level_xyz)
{
  return guarded([&] {
    MFG_REQUIRE(am && level_xyz, "null argument");
    for (uint32_t a = 0; a < am->n_active(); ++a)
      {
        const ACell &c = am->cell(a);
        level_xyz[4 * a] = am->act_level[a];
        for (int d = 0; d < 3; ++d) level_xyz[4 * a + 1 + d] = c.x[d];
      }
  });
}
uint32_t mfg_amesh_n_level_cells(const mfg_amesh *am, int level) { return am && level >= 0 && level < (int)am->levels.size() ? (uint32_t)am->levels[level].size() : 0; }
int mfg_amesh_get_level_cells(const mfg_amesh *am, int level, uint32_t *xyz_children)
{
  return guarded([&] {
    MFG_REQUIRE(am && xyz_children && level >= 0 && level < (int)am->levels.size(), "bad argument");
    const auto &L = am->levels[level];
    for (size_t i = 0; i < L.size(); ++i)
      {
        for (int d = 0; d < 3; ++d) xyz_children[4 * i + d] = L[i].x[d];
        xyz_children[4 * i + 3] = L[i].child0 >= 0 ? 1u : 0u;
      }
  });
}
int mfg_amesh_distribute_dofs(mfg_amesh *am) { return guarded([&] { MFG_REQUIRE(am, "null argument"); distribute_dofs(am); }); }
uint32_t mfg_amesh_n_dofs(const mfg_amesh *am) { return am && am->dofs_ready ? am->n_dofs : 0; }
uint32_t mfg_amesh_n_constrained(const mfg_amesh *am) { return am && am->dofs_ready ? (uint32_t)am->constrained.size() : 0; }
uint32_t mfg_amesh_n_hanging(const mfg_amesh *am) { return am && am->dofs_ready ? (uint32_t)am->hanging.size() : 0; }
int mfg_amesh_get_arrays(const mfg_amesh *am, uint32_t *loc2glob, uint32_t *loc2glob_unconstrained, uint32_t *constraint_mask, uint32_t *constrained,
                         uint32_t *hanging, double *inv_jac, double *coefficient, double *quadrature_points)
{
  return guarded([&] {
    MFG_REQUIRE(am && am->dofs_ready, "call mfg_amesh_distribute_dofs first");
    if (loc2glob) std::copy(am->l2g.begin(), am->l2g.end(), loc2glob);
    if (loc2glob_unconstrained) std::copy(am->l2g_own.begin(), am->l2g_own.end(), loc2glob_unconstrained);
    if (constraint_mask) std::copy(am->mask.begin(), am->mask.end(), constraint_mask);
    if (constrained) std::copy(am->constrained.begin(), am->constrained.end(), constrained);
    if (hanging) std::copy(am->hanging.begin(), am->hanging.end(), hanging);
    if (inv_jac) std::copy(am->inv_jac.begin(), am->inv_jac.end(), inv_jac);
    if (coefficient || quadrature_points) coefficient_at_qpoints(am, coefficient, quadrature_points);
  });
}

int mfg_amesh_build_mg(mfg_amesh *am, int min_level) { return guarded([&] { MFG_REQUIRE(am, "null argument"); build_mg(am, min_level); }); }
// sizes of a level of the hierarchy: out[0..5] = cells, DoFs, boundary indices, refinement-edge indices, refined cells of level-1
// (transfer blocks into this level), copy-index pairs
int mfg_amesh_mg_level_sizes(const mfg_amesh *am, int level, uint32_t out[6])
{
  return guarded([&] {
    MFG_REQUIRE(am && out && am->mg_ready, "call mfg_amesh_build_mg first");
    MFG_REQUIRE(level >= am->mg_min_level && level < (int)am->levels.size(), "bad level");
    const mfg_amesh::MgLevel &L = am->mg[level - am->mg_min_level];
    out[0] = L.n_cells; out[1] = L.n_dofs; out[2] = (uint32_t)L.boundary.size(); out[3] = (uint32_t)L.edge.size(); out[4] = L.n_parents;
    out[5] = (uint32_t)L.copy_global.size();
  });
}
int mfg_amesh_mg_level_get(const mfg_amesh *am, int level, uint32_t *loc2glob, uint32_t *boundary, uint32_t *edge, double *coefficient, uint32_t *copy_global,
                           uint32_t *copy_level, uint32_t *coarse_idx, uint32_t *fine_idx, double *weights)
{
  return guarded([&] {
    MFG_REQUIRE(am && am->mg_ready, "call mfg_amesh_build_mg first");
    MFG_REQUIRE(level >= am->mg_min_level && level < (int)am->levels.size(), "bad level");
    const mfg_amesh::MgLevel &L = am->mg[level - am->mg_min_level];
    if (loc2glob) std::copy(L.l2g.begin(), L.l2g.end(), loc2glob);
    if (boundary) std::copy(L.boundary.begin(), L.boundary.end(), boundary);
    if (edge) std::copy(L.edge.begin(), L.edge.end(), edge);
    if (coefficient) level_coefficient(am, level, coefficient);
    if (copy_global) std::copy(L.copy_global.begin(), L.copy_global.end(), copy_global);
    if (copy_level) std::copy(L.copy_level.begin(), L.copy_level.end(), copy_level);
    if (coarse_idx) std::copy(L.coarse_idx.begin(), L.coarse_idx.end(), coarse_idx);
    if (fine_idx) std::copy(L.fine_idx.begin(), L.fine_idx.end(), fine_idx);
    if (weights) std::copy(L.weights.begin(), L.weights.end(), weights);
  });
}

uint32_t mfg_amesh_n_boundary(const mfg_amesh *am) { return am && am->dofs_ready ? (uint32_t)am->boundary.size() : 0; }
// DoFs on the boundary of the domain, ascending (VectorTools::interpolate_boundary_values visits these, poisson.cu:155-158)
int mfg_amesh_get_boundary(const mfg_amesh *am, uint32_t *out)
{
  return guarded([&] {
    MFG_REQUIRE(am && am->dofs_ready && (out || am->boundary.empty()), "call mfg_amesh_distribute_dofs first");
    std::copy(am->boundary.begin(), am->boundary.end(), out);
  });
}
// DoFTools::map_dofs_to_support_points: [n_dofs][dim]
int mfg_amesh_get_support_points(const mfg_amesh *am, double *out)
{
  return guarded([&] {
    MFG_REQUIRE(am && am->dofs_ready && out, "call mfg_amesh_distribute_dofs first");
    const int dim = am->dim, n = am->n;
    for (uint32_t a = 0; a < am->n_active(); ++a)
      {
        const ACell &c = am->cell(a);
        const double h = am->cell_h(am->act_level[a]);
        for (uint32_t i = 0; i < am->npc; ++i)
          {
            uint32_t t = i;
            const uint32_t g = am->l2g_own[(size_t)a * am->npc + i];
            for (int d = 0; d < dim; ++d) { out[(size_t)g * dim + d] = am->left + h * ((double)c.x[d] + am->fe.nodes[t % n]); t /= n; }
          }
      }
  });
}
// MatrixFreeGpu::reinit on the adaptive mesh for user-written cell loops (generic FEEvaluationGpu path): masks, rewritten
// loc2glob, J^-1 per cell and the quadrature points
int mfg_mf_reinit_from_amesh(mfg_ctx *ctx, const mfg_amesh *am, mfg_dtype dt, mfg_mf **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && am && out, "null argument");
    MFG_REQUIRE(am->dofs_ready, "call mfg_amesh_distribute_dofs first");
    const uint32_t nact = am->n_active();
    mfg_mf_desc    d;
    std::memset(&d, 0, sizeof(d));
    d.dim = am->dim; d.degree = am->p; d.dtype = dt; d.n_cells = nact; d.n_dofs = am->n_dofs;
    d.loc2glob = am->l2g.data(); d.geometry = MFG_GEOM_UNIFORM; d.inv_jac = am->inv_jac.data();
    d.scatter = MFG_SCATTER_ATOMIC;
    bool any = false;
    for (uint32_t m : am->mask) any = any || m != 0;
    d.constraint_mask = any ? am->mask.data() : nullptr;
    std::vector<double> qp((size_t)nact * am->npc * am->dim);
    coefficient_at_qpoints(am, nullptr, qp.data());
    d.quadrature_points = qp.data();
    *out = mf_from_desc(ctx, d);
  });
}

int mfg_laplace_create_from_amesh(mfg_ctx *ctx, const mfg_amesh *am, mfg_dtype dt, mfg_laplace **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && am && out, "null argument");
    MFG_REQUIRE(am->dofs_ready, "call mfg_amesh_distribute_dofs first");
    const uint32_t nact = am->n_active();
    mfg_mf_desc    d;
    std::memset(&d, 0, sizeof(d));
    d.dim = am->dim; d.degree = am->p; d.dtype = dt; d.n_cells = nact; d.n_dofs = am->n_dofs;
    d.loc2glob = am->l2g.data(); d.geometry = MFG_GEOM_UNIFORM; d.inv_jac = am->inv_jac.data();
    d.scatter = MFG_SCATTER_ATOMIC;
    bool any = false;
    for (uint32_t m : am->mask) any = any || m != 0;
    d.constraint_mask = any ? am->mask.data() : nullptr;
    // quadrature points only where the generic FEEvaluationGpu path can run (it refuses hanging-node cells)
    std::vector<double> coef((size_t)nact * am->npc), qp(any ? 0 : (size_t)nact * am->npc * am->dim);
    coefficient_at_qpoints(am, coef.data(), any ? nullptr : qp.data());
    d.quadrature_points = any ? nullptr : qp.data();
    std::unique_ptr<mfg_mf> mf(mf_from_desc(ctx, d));
    std::unique_ptr<mfg_ch> ch(ch_create(ctx, dt, am->constrained.data(), am->constrained.size(), nullptr, 0));
    mfg_laplace *op = laplace_from_arrays(ctx, mf.get(), ch.get(), coef.data());
    op->owns_mf = true; op->owns_ch = true;
    mf.release(); ch.release();
    *out = op;
  });
}

}  // extern "C"
