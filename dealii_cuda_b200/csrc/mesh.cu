// mesh.cu -- device construction of the uniform box mesh, its FE_Q(p) DoF
// numbering, lexicographic local-to-global map and Dirichlet constraint list.
//
// The reference builds these serially on the host through deal.II, one
// FEValues::reinit per cell (matrix_free_gpu.cu:262-339) -- minutes
// at 1e8 DoFs.  Here the numbering is closed-form per cell once the exclusive
// scan of "DoFs first touched by cell c" is known:
//   * cells are visited along the Morton curve (deal.II order after
//     refine_global); the Morton key is monotone in every coordinate, so every
//     vertex/line/quad is first touched by its lexicographically lowest cell;
//   * a cell therefore numbers exactly the DoFs with local index i_d > 0 in all
//     directions d where it has a lower neighbour: prod_d (p + [c_d == 0]) DoFs;
//   * inside the cell they are numbered in hierarchic order (vertices, lines,
//     quads, hex) -- a table of 2^dim variants indexed by the boundary flags.
#include <memory>
#include <cub/cub.cuh>
#include "mesh.cuh"

namespace mfg {

std::vector<uint32_t> hierarchic_to_lexicographic(int dim, int p)
{
  const int n = p + 1;
  std::vector<uint32_t> h2l;
  auto lex = [&](int x, int y, int z) { return (uint32_t)(x + n * (y + n * z)); };
  if (dim == 2)
    {
      for (int v = 0; v < 4; ++v) h2l.push_back(lex((v & 1) * p, (v >> 1) * p, 0));
      // lines: 0: x=0 (along y), 1: x=1, 2: y=0 (along x), 3: y=1
      for (int l = 0; l < 4; ++l)
        for (int t = 1; t < p; ++t)
          h2l.push_back(l < 2 ? lex((l & 1) * p, t, 0) : lex(t, (l & 1) * p, 0));
      for (int y = 1; y < p; ++y) for (int x = 1; x < p; ++x) h2l.push_back(lex(x, y, 0));
    }
  else
    {
      for (int v = 0; v < 8; ++v) h2l.push_back(lex((v & 1) * p, ((v >> 1) & 1) * p, (v >> 2) * p));
      // lines 0-3 on z=0, 4-7 on z=1: (x=0 | x=1) along y, (y=0 | y=1) along x
      for (int zz = 0; zz < 2; ++zz)
        for (int l = 0; l < 4; ++l)
          for (int t = 1; t < p; ++t)
            h2l.push_back(l < 2 ? lex((l & 1) * p, t, zz * p) : lex(t, (l & 1) * p, zz * p));
      // lines 8-11 along z at (x,y) = (0,0),(1,0),(0,1),(1,1)
      for (int l = 0; l < 4; ++l)
        for (int t = 1; t < p; ++t) h2l.push_back(lex((l & 1) * p, (l >> 1) * p, t));
      // quads: x-faces (y fastest), y-faces (z fastest), z-faces (x fastest)
      for (int f = 0; f < 2; ++f) for (int z = 1; z < p; ++z) for (int y = 1; y < p; ++y) h2l.push_back(lex(f * p, y, z));
      for (int f = 0; f < 2; ++f) for (int x = 1; x < p; ++x) for (int z = 1; z < p; ++z) h2l.push_back(lex(x, f * p, z));
      for (int f = 0; f < 2; ++f) for (int y = 1; y < p; ++y) for (int x = 1; x < p; ++x) h2l.push_back(lex(x, y, f * p));
      for (int z = 1; z < p; ++z) for (int y = 1; y < p; ++y) for (int x = 1; x < p; ++x) h2l.push_back(lex(x, y, z));
    }
  return h2l;
}

namespace {

struct MeshParams
{
  MortonMap mm;
  int       dim, p, n;
  uint32_t  nc[3];
  uint32_t  npc, n_cells;
  uint32_t  dirichlet_faces;
};

__global__ void count_new_dofs(MeshParams P, uint32_t *cnt)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_cells) return;
  uint32_t x[3]; P.mm.decode(c, x);
  uint32_t k = 1;
  for (int d = 0; d < P.dim; ++d) k *= (uint32_t)P.p + (x[d] == 0 ? 1u : 0u);
  cnt[c] = k;
}

// global index of local lexicographic dof (i0,i1,i2) of the cell with coordinates x
__device__ inline uint32_t dof_of(const MeshParams &P, const uint32_t *cell_first, const uint16_t *rank_table,
                                  const uint32_t x[3], const int i[3])
{
  uint32_t o[3] = {0, 0, 0}; int il[3] = {0, 0, 0}; uint32_t flags = 0;
  for (int d = 0; d < P.dim; ++d)
    {
      const bool lower = (i[d] == 0 && x[d] > 0);
      o[d]  = x[d] - (lower ? 1u : 0u);
      il[d] = lower ? P.p : i[d];
      flags |= (o[d] == 0 ? 1u : 0u) << d;
    }
  const uint32_t oc  = P.mm.encode(o);
  const uint32_t lix = il[0] + P.n * (il[1] + P.n * il[2]);
  return cell_first[oc] + rank_table[flags * P.npc + lix];
}

__global__ void build_l2g(MeshParams P, const uint32_t *cell_first, const uint16_t *rank_table, uint32_t *l2g, uint8_t *cflag)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)P.n_cells * P.npc) return;
  const uint32_t c = (uint32_t)(t / P.npc), li = (uint32_t)(t % P.npc);
  uint32_t x[3]; P.mm.decode(c, x);
  const int i[3] = {(int)(li % P.n), (int)((li / P.n) % P.n), (int)(li / (P.n * P.n))};
  const uint32_t g = dof_of(P, cell_first, rank_table, x, i);
  l2g[t] = g;
  // interpolate_boundary_values(dof_handler, 0, ZeroFunction): all DoFs on Dirichlet faces
  bool onb = false;
  for (int d = 0; d < P.dim; ++d)
    {
      if (x[d] == 0 && i[d] == 0 && (P.dirichlet_faces >> (2 * d)) & 1u) onb = true;
      if (x[d] == P.nc[d] - 1 && i[d] == P.p && (P.dirichlet_faces >> (2 * d + 1)) & 1u) onb = true;
    }
  if (onb) cflag[g] = 1;
}

__global__ void lattice_lookup(MeshParams P, const uint32_t *cell_first, const uint16_t *rank_table, size_t npts,
                               const uint32_t *xyz, uint32_t *out)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npts) return;
  uint32_t x[3] = {0, 0, 0}; int i[3] = {0, 0, 0};
  for (int d = 0; d < P.dim; ++d)
    {
      const uint32_t X = xyz[3 * t + d];
      uint32_t c = X / P.p; if (c >= P.nc[d]) c = P.nc[d] - 1;
      x[d] = c; i[d] = (int)(X - c * P.p);
    }
  out[t] = dof_of(P, cell_first, rank_table, x, i);
}

__global__ void cell_coords_kernel(MeshParams P, uint32_t *out)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.n_cells) return;
  uint32_t x[3]; P.mm.decode(c, x);
  out[3 * (size_t)c + 0] = x[0]; out[3 * (size_t)c + 1] = x[1]; out[3 * (size_t)c + 2] = x[2];
}

// support point of every DoF: x = origin + h (cell + node_i); shared DoFs get the same value from each of their cells
struct NodeTable { double x[9]; };
__global__ void support_points_kernel(MeshParams P, const uint32_t *__restrict__ l2g, NodeTable nodes, double ox, double oy, double oz, double h,
                                      double *__restrict__ out)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)P.n_cells * P.npc) return;
  const uint32_t c = (uint32_t)(t / P.npc), e = (uint32_t)(t % P.npc);
  uint32_t x[3]; P.mm.decode(c, x);
  const int li[3] = {(int)(e % P.n), (int)((e / P.n) % P.n), (int)(e / (P.n * P.n))};
  const double o[3] = {ox, oy, oz};
  const uint32_t g = l2g[t];
  for (int d = 0; d < P.dim; ++d) out[(size_t)g * P.dim + d] = o[d] + h * ((double)x[d] + nodes.x[li[d]]);
}

struct IotaOp { __host__ __device__ uint32_t operator()(uint32_t i) const { return i; } };

MeshParams params_of(const mfg_mesh *m)
{
  MeshParams P;
  P.mm.dim = m->dim; for (int d = 0; d < 3; ++d) { P.mm.lg[d] = m->lg[d]; P.nc[d] = m->nc[d]; }
  P.dim = m->dim; P.p = m->p; P.n = m->n; P.npc = m->npc; P.n_cells = m->n_cells; P.dirichlet_faces = m->dirichlet_faces;
  return P;
}

}  // namespace

mfg_mesh *build_box_mesh(mfg_ctx *ctx, const mfg_box_desc &d)
{
  MFG_REQUIRE(d.dim == 2 || d.dim == 3, "dim must be 2 or 3");
  MFG_REQUIRE(d.degree >= 1 && d.degree <= 8, "degree must be in 1..8");
  MFG_REQUIRE(d.h > 0, "h must be positive");
  int totbits = 0;
  for (int k = 0; k < d.dim; ++k) { MFG_REQUIRE(d.log2_cells[k] >= 0 && d.log2_cells[k] <= 10, "log2_cells out of range"); totbits += d.log2_cells[k]; }
  MFG_REQUIRE(totbits <= 26, "too many cells for one device partition");
  std::unique_ptr<mfg_mesh> m(new mfg_mesh);
  m->ctx = ctx; m->dim = d.dim; m->p = d.degree; m->n = d.degree + 1; m->h = d.h;
  m->dirichlet_faces = d.dirichlet_faces;
  m->n_cells = 1;
  for (int k = 0; k < 3; ++k)
    {
      m->lg[k] = k < d.dim ? d.log2_cells[k] : 0; m->nc[k] = 1u << m->lg[k]; m->origin[k] = k < d.dim ? d.origin[k] : 0.0;
      m->n_cells *= m->nc[k];
    }
  m->npc = ipow(m->n, m->dim);
  m->fe  = make_fe_data(m->p);
  {
    unsigned long long nd = 1;
    for (int k = 0; k < m->dim; ++k) nd *= (unsigned long long)m->p * m->nc[k] + 1;
    MFG_REQUIRE(nd < (1ull << 31), "n_dofs must stay below 2^31 per device partition (32-bit indices, bit 31 reserved)");
    m->n_dofs = (uint32_t)nd;
  }
  cudaStream_t s = ctx->stream;

  // rank table: position of an owned local DoF among the DoFs its cell numbers, hierarchic order
  const std::vector<uint32_t> h2l = hierarchic_to_lexicographic(m->dim, m->p);
  std::vector<uint16_t> rank(8 * (size_t)m->npc, 0xffff);
  for (uint32_t flags = 0; flags < (1u << m->dim); ++flags)
    {
      uint16_t r = 0;
      for (uint32_t hI = 0; hI < m->npc; ++hI)
        {
          const uint32_t li = h2l[hI];
          const int i[3] = {(int)(li % m->n), (int)((li / m->n) % m->n), (int)(li / (m->n * m->n))};
          bool owned = true;
          for (int k = 0; k < m->dim; ++k) if (i[k] == 0 && !((flags >> k) & 1u)) owned = false;
          if (owned) rank[flags * m->npc + li] = r++;
        }
    }
  m->rank_table.upload(rank.data(), rank.size(), s);

  const MeshParams P = params_of(m.get());
  DevBuf<uint32_t> cnt(m->n_cells);
  m->cell_first.alloc(m->n_cells);
  count_new_dofs<<<(m->n_cells + 255) / 256, 256, 0, s>>>(P, cnt.p);
  MFG_CUDA_LAST();
  {
    size_t tmp_bytes = 0;
    MFG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.p, m->cell_first.p, (int)m->n_cells, s));
    DevBuf<uint8_t> tmp(tmp_bytes);
    MFG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.p, m->cell_first.p, (int)m->n_cells, s));
    MFG_CUDA(cudaStreamSynchronize(s));
  }
  m->l2g.alloc((size_t)m->n_cells * m->npc);
  m->cflag.alloc(m->n_dofs);
  MFG_CUDA(cudaMemsetAsync(m->cflag.p, 0, m->n_dofs, s));
  {
    const size_t tot = (size_t)m->n_cells * m->npc;
    build_l2g<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(P, m->cell_first.p, m->rank_table.p, m->l2g.p, m->cflag.p);
    MFG_CUDA_LAST();
  }
  // ascending list of constrained DoFs (ConstraintHandlerGpu::reinit, constraint_handler_gpu.cu:77-83)
  {
    DevBuf<uint32_t> sel(m->n_dofs), nsel(1);
    cub::TransformInputIterator<uint32_t, IotaOp, cub::CountingInputIterator<uint32_t>> iota(cub::CountingInputIterator<uint32_t>(0), IotaOp());
    size_t tmp_bytes = 0;
    MFG_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, m->cflag.p, sel.p, nsel.p, (int)m->n_dofs, s));
    DevBuf<uint8_t> tmp(tmp_bytes);
    MFG_CUDA(cub::DeviceSelect::Flagged(tmp.p, tmp_bytes, iota, m->cflag.p, sel.p, nsel.p, (int)m->n_dofs, s));
    uint32_t nc = 0;
    MFG_CUDA(cudaMemcpyAsync(&nc, nsel.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    MFG_CUDA(cudaStreamSynchronize(s));
    m->n_constrained = nc;
    m->constrained.alloc(nc);
    if (nc) MFG_CUDA(cudaMemcpyAsync(m->constrained.p, sel.p, nc * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    MFG_CUDA(cudaStreamSynchronize(s));
  }
  return m.release();
}

void mesh_lattice_to_dof(const mfg_mesh *m, size_t npts, const uint32_t *xyz_host, uint32_t *out_host)
{
  if (!npts) return;
  cudaStream_t s = m->ctx->stream;
  DevBuf<uint32_t> xyz, out(npts);
  xyz.upload(xyz_host, 3 * npts, s);
  lattice_lookup<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(params_of(m), m->cell_first.p, m->rank_table.p, npts, xyz.p, out.p);
  MFG_CUDA_LAST();
  out.download(out_host, s);
}

void mesh_lattice_to_dof_device(const mfg_mesh *m, size_t npts, const uint32_t *xyz_dev, uint32_t *out_dev)
{
  if (!npts) return;
  lattice_lookup<<<(unsigned)((npts + 255) / 256), 256, 0, m->ctx->stream>>>(params_of(m), m->cell_first.p, m->rank_table.p, npts, xyz_dev, out_dev);
  MFG_CUDA_LAST();
}

void mesh_support_points(const mfg_mesh *m, double *out_host)
{
  cudaStream_t s = m->ctx->stream;
  DevBuf<double> out((size_t)m->n_dofs * m->dim);
  NodeTable nt;
  for (int i = 0; i < m->n; ++i) nt.x[i] = m->fe.nodes[i];
  const size_t total = (size_t)m->n_cells * m->npc;
  if (total) support_points_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(params_of(m), m->l2g.p, nt, m->origin[0], m->origin[1], m->origin[2], m->h, out.p);
  MFG_CUDA_LAST();
  out.download(out_host, s);
}

void mesh_cell_coords(const mfg_mesh *m, uint32_t *out_host)
{
  cudaStream_t s = m->ctx->stream;
  DevBuf<uint32_t> out((size_t)m->n_cells * 3);
  cell_coords_kernel<<<(m->n_cells + 255) / 256, 256, 0, s>>>(params_of(m), out.p);
  MFG_CUDA_LAST();
  out.download(out_host, s);
}

}  // namespace mfg
