// coloring.cu -- graph coloring of the cells for the atomics-free scatter (host code).
//
// The reference calls deal.II's GraphColoring::make_graph_coloring on the cells with the cell DoFs (after
// ConstraintMatrix::resolve_indices) as conflict indices (matrix_free_gpu/coloring.cc:8-33).  deal.II is not available here;
// this file restates its three steps (deal.II base/graph_coloring.h, from memory -- SURVEY Appendix A.5; the exact
// tie-breaking of deal.II is not verifiable here, so colorings are checked for VALIDITY, not for equality with deal.II's):
//   1. create_partitioning: zone 0 = {first cell}; zone k+1 = the unused cells that share a conflict index with zone k, in
//      the order they are discovered; when a zone comes out empty while cells remain, the first unused cell seeds a new one;
//   2. make_dsatur_coloring per zone: repeatedly take the uncolored cell with the largest saturation degree (number of
//      different colors among its neighbours), ties by the larger degree, then by the smaller position; give it the
//      smallest color none of its neighbours has;
//   3. gather_colors: cells of even zones never conflict with cells of other even zones (likewise odd), so the colors of
//      all even zones are merged (each zone's colors are dealt to the merged colors with the fewest cells so far), the same
//      for the odd zones; the result is the even colors followed by the odd ones.
#include <algorithm>
#include <set>
#include <vector>
#include "common.cuh"

namespace mfg {

// color_of_cell[c] in [0, n_colors); conflict: [n_cells][npc] indices < n_index (bit 31 is ignored)
void graph_coloring_dealii(uint32_t n_cells, uint32_t npc, const uint32_t *conflict, uint32_t n_index, uint32_t *color_of_cell, uint32_t *n_colors)
{
  *n_colors = 0;
  if (n_cells == 0) return;
  // index -> cells (CSR)
  std::vector<uint32_t> start((size_t)n_index + 1, 0);
  for (size_t t = 0; t < (size_t)n_cells * npc; ++t) ++start[(conflict[t] & 0x7fffffffu) + 1];
  for (uint32_t i = 0; i < n_index; ++i) start[i + 1] += start[i];
  std::vector<uint32_t> cells_of(start[n_index]), fill(start.begin(), start.end() - 1);
  for (uint32_t c = 0; c < n_cells; ++c)
    for (uint32_t k = 0; k < npc; ++k) cells_of[fill[conflict[(size_t)c * npc + k] & 0x7fffffffu]++] = c;
  // ---- 1. zones ----
  std::vector<int>      zone_of(n_cells, -1);
  std::vector<std::vector<uint32_t>> zones;
  uint32_t next_unused = 0, n_used = 0;
  std::vector<uint32_t> current(1, 0u);
  zone_of[0] = 0; n_used = 1;
  while (true)
    {
      zones.push_back(current);
      if (n_used == n_cells) break;
      std::vector<uint32_t> next;
      const int zid = (int)zones.size();
      for (uint32_t c : current)
        for (uint32_t k = 0; k < npc; ++k)
          {
            const uint32_t idx = conflict[(size_t)c * npc + k] & 0x7fffffffu;
            for (uint32_t p = start[idx]; p < start[idx + 1]; ++p)
              {
                const uint32_t o = cells_of[p];
                if (zone_of[o] < 0) { zone_of[o] = zid; next.push_back(o); ++n_used; }
              }
          }
      if (next.empty())
        {  // disconnected remainder: the first unused cell seeds a new zone
          while (zone_of[next_unused] >= 0) ++next_unused;
          zone_of[next_unused] = zid; next.push_back(next_unused); ++n_used;
        }
      current.swap(next);
    }
  // ---- 2. DSATUR per zone ----
  std::vector<std::vector<std::vector<uint32_t>>> zone_colors(zones.size());
  std::vector<int> local(n_cells, -1);
  for (size_t z = 0; z < zones.size(); ++z)
    {
      const std::vector<uint32_t> &Z = zones[z];
      const int nz = (int)Z.size();
      for (int i = 0; i < nz; ++i) local[Z[i]] = i;
      std::vector<std::vector<int>> nb(nz);
      for (int i = 0; i < nz; ++i)
        {
          for (uint32_t k = 0; k < npc; ++k)
            {
              const uint32_t idx = conflict[(size_t)Z[i] * npc + k] & 0x7fffffffu;
              for (uint32_t p = start[idx]; p < start[idx + 1]; ++p)
                {
                  const uint32_t o = cells_of[p];
                  if (o != Z[i] && zone_of[o] == (int)z) nb[i].push_back(local[o]);
                }
            }
          std::sort(nb[i].begin(), nb[i].end());
          nb[i].erase(std::unique(nb[i].begin(), nb[i].end()), nb[i].end());
        }
      std::vector<int> col(nz, -1), sat(nz, 0);
      std::vector<std::vector<char>> seen(nz);  // colors seen among the neighbours
      // order: saturation desc, degree desc, position asc
      auto key = [&](int i) { return std::make_tuple(-sat[i], -(int)nb[i].size(), i); };
      std::set<std::tuple<int, int, int>> queue;
      for (int i = 0; i < nz; ++i) queue.insert(key(i));
      int ncol = 0;
      while (!queue.empty())
        {
          const int i = std::get<2>(*queue.begin());
          queue.erase(queue.begin());
          std::vector<char> used(ncol + 1, 0);
          for (int j : nb[i]) if (col[j] >= 0) used[col[j]] = 1;
          int c = 0;
          while (used[c]) ++c;
          col[i] = c;
          ncol = std::max(ncol, c + 1);
          for (int j : nb[i])
            if (col[j] < 0)
              {
                if ((int)seen[j].size() <= c) seen[j].resize(c + 1, 0);
                if (!seen[j][c]) { queue.erase(key(j)); seen[j][c] = 1; ++sat[j]; queue.insert(key(j)); }
              }
        }
      zone_colors[z].assign(ncol, {});
      for (int i = 0; i < nz; ++i) zone_colors[z][col[i]].push_back(Z[i]);
      for (int i = 0; i < nz; ++i) local[Z[i]] = -1;
    }
  // ---- 3. merge the colors of the even zones, then of the odd zones ----
  std::vector<std::vector<uint32_t>> all;
  for (int parity = 0; parity < 2; ++parity)
    {
      size_t max_colors = 0;
      for (size_t z = parity; z < zones.size(); z += 2) max_colors = std::max(max_colors, zone_colors[z].size());
      if (max_colors == 0) continue;
      std::vector<std::vector<uint32_t>> merged(max_colors);
      for (size_t z = parity; z < zones.size(); z += 2)
        {
          // largest color of the zone to the smallest merged color not yet used for this zone
          std::vector<size_t> order(zone_colors[z].size());
          for (size_t k = 0; k < order.size(); ++k) order[k] = k;
          std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return zone_colors[z][a].size() > zone_colors[z][b].size(); });
          std::vector<char> taken(max_colors, 0);
          for (size_t k : order)
            {
              size_t best = max_colors;
              for (size_t m = 0; m < max_colors; ++m)
                if (!taken[m] && (best == max_colors || merged[m].size() < merged[best].size())) best = m;
              taken[best] = 1;
              merged[best].insert(merged[best].end(), zone_colors[z][k].begin(), zone_colors[z][k].end());
            }
        }
      for (auto &m : merged) if (!m.empty()) all.push_back(std::move(m));
    }
  for (size_t k = 0; k < all.size(); ++k)
    for (uint32_t c : all[k]) color_of_cell[c] = (uint32_t)k;
  *n_colors = (uint32_t)all.size();
}

}  // namespace mfg

extern "C" int mfg_graph_coloring(uint32_t n_cells, uint32_t dofs_per_cell, const uint32_t *conflict_indices_host, uint32_t n_indices,
                                  uint32_t *color_of_cell, uint32_t *n_colors)
{
  return mfg::guarded([&] {
    MFG_REQUIRE(color_of_cell && n_colors && (n_cells == 0 || conflict_indices_host), "null argument");
    for (size_t t = 0; t < (size_t)n_cells * dofs_per_cell; ++t)
      MFG_REQUIRE((conflict_indices_host[t] & 0x7fffffffu) < n_indices, "conflict index out of range");
    mfg::graph_coloring_dealii(n_cells, dofs_per_cell, conflict_indices_host, n_indices, color_of_cell, n_colors);
  });
}
