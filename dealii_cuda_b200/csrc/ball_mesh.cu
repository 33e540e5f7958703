// ball_mesh.cu -- host substrate for the reference's BALL_GRID runs (poisson_common.h:59-72: GridGenerator::hyper_ball with a
// SphericalManifold on the boundary, refine_global) and their non-affine geometry (MappingQ1, matrix_free_gpu.h:257:
// FEValues::get_inverse_jacobians / get_JxW_values per quadrature point, matrix_free_gpu.cu:326-338).
//
// Restated without deal.II:
//   * the coarse mesh: a centre square / cube and 2 dim cells between it and the circle / sphere (5 quads, 7 hexes), all cells
//     right-handed with their local axes along the global ones;
//   * refine_global: every cell into 2^dim children; new vertices are the means of the parent entity's vertices (edge
//     midpoints, face centres, cell centre), those on boundary entities are moved onto the sphere (SphericalManifold);
//   * an UNSTRUCTURED mesh: entities are identified by their vertex numbers (edge = two, face = four), so FE_Q DoFs on shared
//     edges / faces are matched whatever the orientation of the two cells: the position along an edge is counted from its lower
//     vertex number, the position in a face from its lowest vertex number towards the lower of the two neighbouring ones (the
//     Gauss-Lobatto points are symmetric, so a reversed edge maps nodes to nodes); DoFs are numbered by first touch over the
//     cells, hierarchic order inside a cell (DoFHandler::distribute_dofs);
//   * Dirichlet DoFs: those on faces that belong to one cell only;
//   * geometry: tri-linear map of the cell vertices (MappingQ1); K = (dx/dxi)^-1 and JxW = det(dx/dxi) w_q at the Gauss points.
// deal.II's vertex numbers, cell order and its placement of interior vertices next to a curved boundary are not reproduced
// (parity unpinned here as everywhere at the deal.II boundary); the operator on this mesh goes through the general-geometry path
// (MFG_GEOM_GENERAL) of mfg_mf_reinit.  Host code; checker: tests/test_ball_mesh.py (numpy).
#include <algorithm>
#include <array>
#include <cmath>
#include <map>
#include <memory>
#include <unordered_map>
#include "mesh.cuh"
#include "operators.cuh"

using namespace mfg;

struct mfg_umesh
{
  int    dim = 0, p = 0, n = 0;
  double radius = 1;
  std::vector<std::array<double, 3>>   verts;
  std::vector<std::array<uint32_t, 8>> cells;  // vertex numbers in local lexicographic order (x fastest)
  FEData1D                             fe;
  // after distribute_dofs
  bool                  dofs_ready = false;
  uint32_t              n_dofs = 0, npc = 0;
  std::vector<uint32_t> l2g, boundary;
};

namespace {

struct Key4
{
  uint32_t v[4];
  bool     operator<(const Key4 &o) const { return std::lexicographical_compare(v, v + 4, o.v, o.v + 4); }
};

Key4 sorted_key(const uint32_t *ids, int k)
{
  MFG_REQUIRE(k >= 1 && k <= 4, "an entity key holds at most four vertices");
  Key4 key{{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}};
  for (int i = 0; i < k; ++i) key.v[i] = ids[i];
  std::sort(key.v, key.v + k);
  return key;
}

// local vertices of the faces of a cell: face 2 d + side is the one with local coordinate d equal to side
void face_vertices(int dim, int d, int side, int out[4])
{
  int k = 0;
  for (int v = 0; v < (1 << dim); ++v)
    if (((v >> d) & 1) == side) out[k++] = v;
}

// faces (edges in 2D) that belong to exactly one cell
std::map<Key4, int> count_faces(const mfg_umesh *um)
{
  std::map<Key4, int> cnt;
  const int           nfv = 1 << (um->dim - 1);
  for (const auto &c : um->cells)
    for (int d = 0; d < um->dim; ++d)
      for (int side = 0; side < 2; ++side)
        {
          int      lv[4];
          uint32_t ids[4];
          face_vertices(um->dim, d, side, lv);
          for (int k = 0; k < nfv; ++k) ids[k] = c[lv[k]];
          ++cnt[sorted_key(ids, nfv)];
        }
  return cnt;
}

void project(std::array<double, 3> &x, double radius)
{
  const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  for (int d = 0; d < 3; ++d) x[d] *= radius / r;
}

void refine_global_once(mfg_umesh *um)
{
  const int dim = um->dim, nv = 1 << dim;
  // boundary entities: boundary faces, and (3D) their edges
  const std::map<Key4, int> faces = count_faces(um);
  std::map<Key4, bool>      bedge;
  if (dim == 3)
    for (const auto &c : um->cells)
      for (int d = 0; d < 3; ++d)
        for (int side = 0; side < 2; ++side)
          {
            int      lv[4];
            uint32_t ids[4];
            face_vertices(3, d, side, lv);
            for (int k = 0; k < 4; ++k) ids[k] = c[lv[k]];
            if (faces.at(sorted_key(ids, 4)) != 1) continue;
            // the four edges of the face: pairs of its vertices that differ in one local direction
            for (int a = 0; a < 4; ++a)
              for (int b = a + 1; b < 4; ++b)
                {
                  const int diff = lv[a] ^ lv[b];
                  if (diff & (diff - 1)) continue;
                  const uint32_t e[2] = {c[lv[a]], c[lv[b]]};
                  bedge[sorted_key(e, 2)] = true;
                }
          }
  std::map<Key4, uint32_t> mid;  // entity (2 or 4 vertices) -> its new vertex
  auto new_vertex = [&](const uint32_t *ids, int k, bool on_boundary) {
    const Key4 key = sorted_key(ids, k);
    auto       it = mid.find(key);
    if (it != mid.end()) return it->second;
    std::array<double, 3> x{0, 0, 0};
    for (int i = 0; i < k; ++i)
      for (int d = 0; d < 3; ++d) x[d] += um->verts[ids[i]][d] / k;
    if (on_boundary) project(x, um->radius);
    um->verts.push_back(x);
    mid[key] = (uint32_t)um->verts.size() - 1;
    return mid[key];
  };
  std::vector<std::array<uint32_t, 8>> children;
  children.reserve(um->cells.size() << dim);
  for (const auto &c : um->cells)
    {
      // the 3^dim lattice of the cell: index t_d in {0, 1, 2} per direction, 1 = a new vertex
      uint32_t  lat[27];
      const int n3 = dim == 3 ? 27 : 9;
      for (int L = 0; L < n3; ++L)
        {
          const int t[3] = {L % 3, (L / 3) % 3, dim == 3 ? L / 9 : 0};
          uint32_t  ids[8];
          int       k = 0;
          for (int v = 0; v < nv; ++v)
            {
              bool in = true;
              for (int d = 0; d < dim; ++d)
                if (t[d] != 1 && ((v >> d) & 1) != t[d] / 2) in = false;
              if (in) ids[k++] = c[v];
            }
          if (k == 1) { lat[L] = ids[0]; continue; }
          if (k == nv)  // the cell centre: interior, belongs to this cell alone
            {
              std::array<double, 3> x{0, 0, 0};
              for (int i = 0; i < k; ++i)
                for (int d = 0; d < 3; ++d) x[d] += um->verts[ids[i]][d] / k;
              um->verts.push_back(x);
              lat[L] = (uint32_t)um->verts.size() - 1;
              continue;
            }
          bool on_boundary = false;
          if (k == (1 << (dim - 1))) on_boundary = faces.at(sorted_key(ids, k)) == 1;   // a face (2D: an edge)
          else if (dim == 3 && k == 2) on_boundary = bedge.count(sorted_key(ids, 2)) != 0;
          lat[L] = new_vertex(ids, k, on_boundary);
        }
      for (int ch = 0; ch < nv; ++ch)  // children in lexicographic order, like deal.II's child numbering on aligned cells
        {
          std::array<uint32_t, 8> cc{};
          for (int v = 0; v < nv; ++v)
            {
              int L = 0, s = 1;
              for (int d = 0; d < dim; ++d) { L += (((ch >> d) & 1) + ((v >> d) & 1)) * s; s *= 3; }
              cc[v] = lat[L];
            }
          children.push_back(cc);
        }
    }
  um->cells.swap(children);
  um->dofs_ready = false;
}

struct DofKey
{
  uint32_t kind, a, b, c, d, s, t;  // kind 0 vertex, 1 edge, 2 face, 3 cell interior
  bool     operator==(const DofKey &o) const { return kind == o.kind && a == o.a && b == o.b && c == o.c && d == o.d && s == o.s && t == o.t; }
};
struct DofKeyHash
{
  size_t operator()(const DofKey &k) const
  {
    uint64_t h = 1469598103934665603ull;
    for (uint32_t w : {k.kind, k.a, k.b, k.c, k.d, k.s, k.t}) { h ^= w; h *= 1099511628211ull; }
    return (size_t)h;
  }
};

void distribute_dofs(mfg_umesh *um)
{
  const int      dim = um->dim, p = um->p, n = um->n;
  const uint32_t npc = ipow(n, dim), nc = (uint32_t)um->cells.size();
  MFG_REQUIRE((uint64_t)nc * npc < (1ull << 32), "too many cells for 32-bit local-to-global offsets");
  um->npc = npc;
  const std::vector<uint32_t> h2l = hierarchic_to_lexicographic(dim, p);
  std::unordered_map<DofKey, uint32_t, DofKeyHash> number;
  number.reserve((size_t)nc * npc / 2);
  um->l2g.assign((size_t)nc * npc, 0);
  uint32_t nxt = 0;
  for (uint32_t ci = 0; ci < nc; ++ci)
    {
      const auto &c = um->cells[ci];
      for (uint32_t hI = 0; hI < npc; ++hI)
        {
          const uint32_t li = h2l[hI];
          int            idx[3] = {0, 0, 0}, free_dirs[3], nfree = 0;
          {
            uint32_t t = li;
            for (int d = 0; d < dim; ++d) { idx[d] = t % n; t /= n; }
          }
          for (int d = 0; d < dim; ++d)
            if (idx[d] > 0 && idx[d] < p) free_dirs[nfree++] = d;
          // the local vertex at the "0" end of every free direction
          int base = 0;
          for (int d = 0; d < dim; ++d)
            if (idx[d] == p) base |= 1 << d;
          DofKey key{0, 0, 0, 0, 0, 0, 0};
          if (nfree == 0) { key.kind = 0; key.a = c[base]; }
          else if (nfree == 1)
            {
              const int      d = free_dirs[0];
              const uint32_t v0 = c[base], v1 = c[base | (1 << d)];
              key.kind = 1; key.a = std::min(v0, v1); key.b = std::max(v0, v1);
              key.s = v0 < v1 ? idx[d] : p - idx[d];                     // position counted from the lower vertex number
            }
          else if (nfree == 2 && dim == 3)
            {
              const int      d0 = free_dirs[0], d1 = free_dirs[1];
              const uint32_t corner[4] = {c[base], c[base | (1 << d0)], c[base | (1 << d1)], c[base | (1 << d0) | (1 << d1)]};
              const Key4     fk = sorted_key(corner, 4);
              key.kind = 2; key.a = fk.v[0]; key.b = fk.v[1]; key.c = fk.v[2]; key.d = fk.v[3];
              // origin = the corner with the lowest vertex number; first axis towards the lower of its two neighbours in the face
              int o = 0;
              for (int k = 1; k < 4; ++k)
                if (corner[k] < corner[o]) o = k;
              const int a0 = (o & 1) ? p - idx[d0] : idx[d0], a1 = (o & 2) ? p - idx[d1] : idx[d1];
              const uint32_t nb0 = corner[o ^ 1], nb1 = corner[o ^ 2];   // neighbour along d0, along d1
              if (nb0 < nb1) { key.s = a0; key.t = a1; }
              else { key.s = a1; key.t = a0; }
            }
          else { key.kind = 3; key.a = ci; key.s = li; }                  // interior of the cell (2D: the quad, 3D: the hex)
          auto it = number.find(key);
          uint32_t g;
          if (it == number.end())
            {
              MFG_REQUIRE(nxt < 0x7fffffffu, "more than 2^31 DoFs");
              g = nxt++;
              number.emplace(key, g);
            }
          else g = it->second;
          um->l2g[(size_t)ci * npc + li] = g;
        }
    }
  um->n_dofs = nxt;
  // Dirichlet DoFs: on faces that belong to one cell only
  const std::map<Key4, int> faces = count_faces(um);
  std::vector<uint8_t>      onb(um->n_dofs, 0);
  const int                 nfv = 1 << (dim - 1);
  for (uint32_t ci = 0; ci < nc; ++ci)
    for (int d = 0; d < dim; ++d)
      for (int side = 0; side < 2; ++side)
        {
          int      lv[4];
          uint32_t ids[4];
          face_vertices(dim, d, side, lv);
          for (int k = 0; k < nfv; ++k) ids[k] = um->cells[ci][lv[k]];
          if (faces.at(sorted_key(ids, nfv)) != 1) continue;
          for (uint32_t li = 0; li < npc; ++li)
            {
              uint32_t t = li;
              int      id = 0;
              for (int e = 0; e < dim; ++e) { if (e == d) id = t % n; t /= n; }
              if (id == (side ? p : 0)) onb[um->l2g[(size_t)ci * npc + li]] = 1;
            }
        }
  um->boundary.clear();
  for (uint32_t g = 0; g < um->n_dofs; ++g)
    if (onb[g]) um->boundary.push_back(g);
  um->dofs_ready = true;
}

// x(xi) of the tri-linear map and J = dx/dxi
void map_point(const mfg_umesh *um, const std::array<uint32_t, 8> &c, const double *xi, double *x, double *J)
{
  const int dim = um->dim, nv = 1 << dim;
  for (int d = 0; d < dim; ++d) x[d] = 0;
  for (int i = 0; i < dim * dim; ++i) J[i] = 0;
  for (int v = 0; v < nv; ++v)
    {
      double N = 1, dN[3] = {1, 1, 1};
      for (int e = 0; e < dim; ++e)
        {
          const double f = ((v >> e) & 1) ? xi[e] : 1 - xi[e], df = ((v >> e) & 1) ? 1.0 : -1.0;
          N *= f;
          for (int g = 0; g < dim; ++g) dN[g] *= g == e ? df : f;
        }
      for (int d = 0; d < dim; ++d)
        {
          x[d] += N * um->verts[c[v]][d];
          for (int e = 0; e < dim; ++e) J[d * dim + e] += dN[e] * um->verts[c[v]][d];
        }
    }
}

void invert(int dim, const double *J, double *K, double &det)
{
  if (dim == 2)
    {
      det = J[0] * J[3] - J[1] * J[2];
      K[0] = J[3] / det; K[1] = -J[1] / det; K[2] = -J[2] / det; K[3] = J[0] / det;
      return;
    }
  const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
  det = J[0] * c00 + J[1] * c01 + J[2] * c02;
  K[0] = c00 / det; K[1] = (J[2] * J[7] - J[1] * J[8]) / det; K[2] = (J[1] * J[5] - J[2] * J[4]) / det;
  K[3] = c01 / det; K[4] = (J[0] * J[8] - J[2] * J[6]) / det; K[5] = (J[2] * J[3] - J[0] * J[5]) / det;
  K[6] = c02 / det; K[7] = (J[1] * J[6] - J[0] * J[7]) / det; K[8] = (J[0] * J[4] - J[1] * J[3]) / det;
}

// per cell and Gauss point: K[d1][d2] = d xi_d1 / d x_d2, JxW, x_q, a(x_q) = 1 / (0.05 + 2 |x_q|^2)
void geometry(const mfg_umesh *um, double *inv_jac, double *jxw, double *qpoints, double *coef)
{
  const int dim = um->dim, n = um->n;
  for (size_t ci = 0; ci < um->cells.size(); ++ci)
    for (uint32_t q = 0; q < um->npc; ++q)
      {
        double   xi[3] = {0, 0, 0}, w = 1, x[3] = {0, 0, 0}, J[9], K[9], det;
        uint32_t t = q;
        for (int d = 0; d < dim; ++d) { xi[d] = um->fe.qpts[t % n]; w *= um->fe.qwts[t % n]; t /= n; }
        map_point(um, um->cells[ci], xi, x, J);
        invert(dim, J, K, det);
        MFG_REQUIRE(det > 0, "inverted cell");
        const size_t o = ci * um->npc + q;
        if (inv_jac) std::copy(K, K + dim * dim, inv_jac + o * dim * dim);
        if (jxw) jxw[o] = det * w;
        if (qpoints) std::copy(x, x + dim, qpoints + o * dim);
        if (coef)
          {
            double r2 = 0;
            for (int d = 0; d < dim; ++d) r2 += x[d] * x[d];
            coef[o] = 1.0 / (0.05 + 2.0 * r2);
          }
      }
}

}  // namespace

extern "C" {

// GridGenerator::hyper_ball(center 0, radius) + SphericalManifold on the boundary (poisson_common.h:65-70)
int mfg_umesh_hyper_ball(int dim, int degree, double radius, mfg_umesh **out)
{
  return guarded([&] {
    MFG_REQUIRE(out, "null argument");
    MFG_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    MFG_REQUIRE(degree >= 1 && degree <= 8, "degree must be in 1..8");
    MFG_REQUIRE(radius > 0, "radius must be positive");
    std::unique_ptr<mfg_umesh> um(new mfg_umesh);
    um->dim = dim; um->p = degree; um->n = degree + 1; um->radius = radius;
    um->fe = make_fe_data(degree);
    const int    nv = 1 << dim;
    const double outer = radius / std::sqrt((double)dim), inner = outer / (1.0 + std::sqrt((double)dim));
    // vertices 0 .. nv-1: inner square / cube, nv .. 2 nv - 1: the outer corners on the sphere (lexicographic, x fastest)
    for (int ring = 0; ring < 2; ++ring)
      for (int v = 0; v < nv; ++v)
        {
          std::array<double, 3> x{0, 0, 0};
          for (int d = 0; d < dim; ++d) x[d] = (((v >> d) & 1) ? 1.0 : -1.0) * (ring ? outer : inner);
          um->verts.push_back(x);
        }
    std::array<uint32_t, 8> c{};
    for (int v = 0; v < nv; ++v) c[v] = (uint32_t)v;
    um->cells.push_back(c);  // the centre cell
    for (int d = 0; d < dim; ++d)
      for (int side = 0; side < 2; ++side)
        {
          // local direction d runs outwards on the upper side and inwards on the lower one, so that it always increases with
          // the global coordinate: the cell is right-handed
          for (int v = 0; v < nv; ++v)
            {
              const int  ld = (v >> d) & 1;
              const bool is_outer = side ? ld == 1 : ld == 0;
              const int  corner = side ? (v | (1 << d)) : (v & ~(1 << d));
              c[v] = (uint32_t)(corner + (is_outer ? nv : 0));
            }
          um->cells.push_back(c);
        }
    *out = um.release();
  });
}
int mfg_umesh_destroy(mfg_umesh *um) { return guarded([&] { delete um; }); }
int mfg_umesh_refine_global(mfg_umesh *um, int times)
{
  return guarded([&] {
    MFG_REQUIRE(um && times >= 0, "bad argument");
    for (int t = 0; t < times; ++t) refine_global_once(um);
  });
}
int mfg_umesh_distribute_dofs(mfg_umesh *um) { return guarded([&] { MFG_REQUIRE(um, "null argument"); distribute_dofs(um); }); }
uint32_t mfg_umesh_n_cells(const mfg_umesh *um) { return um ? (uint32_t)um->cells.size() : 0; }
uint32_t mfg_umesh_n_vertices(const mfg_umesh *um) { return um ? (uint32_t)um->verts.size() : 0; }
uint32_t mfg_umesh_n_dofs(const mfg_umesh *um) { return um && um->dofs_ready ? um->n_dofs : 0; }
uint32_t mfg_umesh_n_boundary(const mfg_umesh *um) { return um && um->dofs_ready ? (uint32_t)um->boundary.size() : 0; }
int mfg_umesh_get_mesh(const mfg_umesh *um, double *vertices, uint32_t *cell_vertices)
{
  return guarded([&] {
    MFG_REQUIRE(um, "null argument");
    if (vertices)
      for (size_t v = 0; v < um->verts.size(); ++v)
        for (int d = 0; d < um->dim; ++d) vertices[v * um->dim + d] = um->verts[v][d];
    if (cell_vertices)
      for (size_t c = 0; c < um->cells.size(); ++c)
        for (int v = 0; v < (1 << um->dim); ++v) cell_vertices[c * (1 << um->dim) + v] = um->cells[c][v];
  });
}
int mfg_umesh_get_arrays(const mfg_umesh *um, uint32_t *loc2glob, uint32_t *boundary, double *inv_jac, double *JxW, double *quadrature_points,
                         double *coefficient)
{
  return guarded([&] {
    MFG_REQUIRE(um && um->dofs_ready, "call mfg_umesh_distribute_dofs first");
    if (loc2glob) std::copy(um->l2g.begin(), um->l2g.end(), loc2glob);
    if (boundary) std::copy(um->boundary.begin(), um->boundary.end(), boundary);
    if (inv_jac || JxW || quadrature_points || coefficient) geometry(um, inv_jac, JxW, quadrature_points, coefficient);
  });
}
// DoFTools::map_dofs_to_support_points with MappingQ1: [n_dofs][dim]
int mfg_umesh_get_support_points(const mfg_umesh *um, double *out)
{
  return guarded([&] {
    MFG_REQUIRE(um && um->dofs_ready && out, "call mfg_umesh_distribute_dofs first");
    const int dim = um->dim, n = um->n;
    for (size_t ci = 0; ci < um->cells.size(); ++ci)
      for (uint32_t i = 0; i < um->npc; ++i)
        {
          double   xi[3] = {0, 0, 0}, x[3], J[9];
          uint32_t t = i;
          for (int d = 0; d < dim; ++d) { xi[d] = um->fe.nodes[t % n]; t /= n; }
          map_point(um, um->cells[ci], xi, x, J);
          std::copy(x, x + dim, out + (size_t)um->l2g[ci * um->npc + i] * dim);
        }
  });
}
// MatrixFreeGpu::reinit on the ball for user-written cell loops (generic FEEvaluationGpu path, general geometry)
int mfg_mf_reinit_from_umesh(mfg_ctx *ctx, const mfg_umesh *um, mfg_dtype dt, mfg_mf **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && um && out, "null argument");
    MFG_REQUIRE(um->dofs_ready, "call mfg_umesh_distribute_dofs first");
    const size_t        total = um->cells.size() * um->npc;
    std::vector<double> K(total * um->dim * um->dim), jxw(total), qp(total * um->dim);
    geometry(um, K.data(), jxw.data(), qp.data(), nullptr);
    mfg_mf_desc d;
    std::memset(&d, 0, sizeof(d));
    d.dim = um->dim; d.degree = um->p; d.dtype = dt; d.n_cells = (uint32_t)um->cells.size(); d.n_dofs = um->n_dofs;
    d.loc2glob = um->l2g.data(); d.geometry = MFG_GEOM_GENERAL; d.inv_jac = K.data(); d.JxW = jxw.data(); d.quadrature_points = qp.data();
    d.scatter = MFG_SCATTER_ATOMIC;
    *out = mf_from_desc(ctx, d);
  });
}
// LaplaceOperatorGpu::reinit on the ball: general geometry (full J^-1 per quadrature point), Dirichlet boundary, reference coefficient
int mfg_laplace_create_from_umesh(mfg_ctx *ctx, const mfg_umesh *um, mfg_dtype dt, mfg_laplace **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && um && out, "null argument");
    MFG_REQUIRE(um->dofs_ready, "call mfg_umesh_distribute_dofs first");
    const size_t        total = um->cells.size() * um->npc;
    std::vector<double> K(total * um->dim * um->dim), jxw(total), qp(total * um->dim), coef(total);
    geometry(um, K.data(), jxw.data(), qp.data(), coef.data());
    mfg_mf_desc d;
    std::memset(&d, 0, sizeof(d));
    d.dim = um->dim; d.degree = um->p; d.dtype = dt; d.n_cells = (uint32_t)um->cells.size(); d.n_dofs = um->n_dofs;
    d.loc2glob = um->l2g.data(); d.geometry = MFG_GEOM_GENERAL; d.inv_jac = K.data(); d.JxW = jxw.data(); d.quadrature_points = qp.data();
    d.scatter = MFG_SCATTER_ATOMIC;
    std::unique_ptr<mfg_mf> mf(mf_from_desc(ctx, d));
    std::unique_ptr<mfg_ch> ch(ch_create(ctx, dt, um->boundary.data(), um->boundary.size(), nullptr, 0));
    mfg_laplace *op = laplace_from_arrays(ctx, mf.get(), ch.get(), coef.data());
    op->owns_mf = true; op->owns_ch = true;
    mf.release(); ch.release();
    *out = op;
  });
}

}  // extern "C"
