// operators.cu -- MatrixFreeGpu / ConstraintHandlerGpu / LaplaceOperatorGpu host logic.
#include <algorithm>
#include <numeric>
#include <cstdlib>
#include "kernels_general.cuh"
#include "kernels_stage.cuh"
#include "kernels_slab3.cuh"
#include "operators.cuh"

namespace mfg {

namespace {

// idx_out[s][i] = l2g[perm ? perm[s] : s][i] | (constrained ? bit31 : 0)
__global__ void build_kernel_indices(const uint32_t *l2g, const uint32_t *__restrict__ perm,
                                     const uint8_t *__restrict__ cflag, uint32_t npc, size_t total, uint32_t *out)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const uint32_t s = (uint32_t)(t / npc), i = (uint32_t)(t % npc);
  const uint32_t c = perm ? perm[s] : s;
  uint32_t g = l2g[(size_t)c * npc + i] & ~CONSTRAINED_BIT;
  if (cflag && cflag[g]) g |= CONSTRAINED_BIT;
  out[t] = g;
}

__global__ void flags_from_list(const uint32_t *list, size_t n, uint8_t *flag)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[list[i]] = 1;
}

__global__ void pack_bits(const uint8_t *flag, size_t n, uint32_t *bits)
{
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w * 32 >= n) return;
  uint32_t v = 0;
  for (int b = 0; b < 32; ++b) { const size_t i = w * 32 + b; if (i < n && flag[i]) v |= 1u << b; }
  bits[w] = v;
}

// slab2 kernel: idxP[g][s = j + n k][lane <-> (c, i)] = idx[g*CW + c][i + n j + n^2 k]  (bit 31 set for idle lanes and
// for cells beyond the mesh)
__global__ void build_slab2_indices(const uint32_t *__restrict__ idx, uint32_t n_cells, uint32_t n_groups, int n, Slab2Geom gm,
                                    uint32_t *__restrict__ out)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int    ns = n * n;
  if (t >= (size_t)n_groups * ns * 32) return;
  const int      lane = (int)(t % 32), s = (int)((t / 32) % ns);
  const uint32_t g = (uint32_t)(t / (32 * (size_t)ns));
  // lane map of the kernel (slab2_lane): lane = 16 ch + n cl + i with c = cl + hc ch (no split when cw is odd)
  const bool split = gm.cw % 2 == 0;
  const int  ch = split ? lane / 16 : 0, l16 = split ? lane % 16 : lane;
  uint32_t v = CONSTRAINED_BIT;
  if (l16 < gm.hc * n)
    {
      const int      c = gm.hc * ch + l16 / n, i = l16 % n;
      const uint32_t cell = g * gm.cw + c;
      if (cell < n_cells) v = idx[(size_t)cell * n * ns + i + n * s];
    }
  out[t] = v;
}

// slab2 kernel: face-merge mask of every group (kernels_slab2.cuh): bit (10 d + c) is set when the upper face of cell c in
// direction d and the lower face of cell c + 2^d of the same group carry identical index entries.  A merge whose receiving
// entries would themselves be handed over by an earlier direction without the sender doing the same is dropped, so that
// no contribution can be lost whatever the cell order of the mesh is.
// xzy: the merges run in the order x, z, y (slab3 kernel: x and z in layout C, y in layout A) instead of x, y, z.
__global__ void build_slab2_merge(const uint32_t *__restrict__ idx, uint32_t n_cells, uint32_t n_groups, int n, int cw, int dirs,
                                  uint32_t *__restrict__ out, bool xzy)
{
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const int npc = n * n * n;
  uint32_t  m = 0;
  if (cw <= SLAB2_MERGE_MAX_CW)
    for (int d = 0; d < 3; ++d)
      {
        if (!((dirs >> d) & 1)) continue;
        const int step = 1 << d, sd = d == 0 ? 1 : d == 1 ? n : n * n;  // stride of the direction inside a cell tensor
        for (int c = 0; c + step < cw; ++c)
          {
            const uint32_t a = g * cw + c, b = a + step;
            if (b >= n_cells) continue;
            bool same = true;
            for (int q = 0; q < npc && same; ++q)
              if ((q / sd) % n == n - 1) same = idx[(size_t)a * npc + q] == idx[(size_t)b * npc + q - (n - 1) * sd];
            if (same) m |= 1u << (10 * d + c);
          }
      }
  auto bit = [&](int d, int c) { return c < SLAB2_MERGE_MAX_CW && ((m >> (10 * d + c)) & 1u); };
  if (xzy)
    {
      for (int c = 0; c + 4 < cw && cw <= SLAB2_MERGE_MAX_CW; ++c)
        if (bit(2, c) && bit(0, c + 4) && !bit(0, c)) m &= ~(1u << (20 + c));
      for (int c = 0; c + 2 < cw && cw <= SLAB2_MERGE_MAX_CW; ++c)
        if (bit(1, c) && ((bit(0, c + 2) && !bit(0, c)) || (bit(2, c + 2) && !bit(2, c)))) m &= ~(1u << (10 + c));
      out[g] = m;
      return;
    }
  for (int c = 0; c + 2 < cw && cw <= SLAB2_MERGE_MAX_CW; ++c)
    if (bit(1, c) && bit(0, c + 2) && !bit(0, c)) m &= ~(1u << (10 + c));
  for (int c = 0; c + 4 < cw && cw <= SLAB2_MERGE_MAX_CW; ++c)
    if (bit(2, c) && ((bit(0, c + 4) && !bit(0, c)) || (bit(1, c + 4) && !bit(1, c)))) m &= ~(1u << (20 + c));
  out[g] = m;
}

// slab2 kernel: coefficient image of a group, element (c,i,j,k) at SC c + SI i + SJ j + SK k (padding stays zero)
template <typename Number>
__global__ void build_slab2_weights(const Number *__restrict__ cw, uint32_t n_cells, int n, Slab2Geom gm, Number *__restrict__ out)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int    npc = n * n * n;
  if (t >= (size_t)n_cells * npc) return;
  const uint32_t cell = (uint32_t)(t / npc);
  const int      q = (int)(t % npc), i = q % n, j = (q / n) % n, k = q / (n * n);
  const uint32_t g = cell / gm.cw, c = cell % gm.cw;
  out[(size_t)g * gm.cwf + gm.bc.SL * (c % gm.hc) + gm.bc.SH * (c / gm.hc) + gm.bc.SI * i + gm.bc.SJ * j + gm.bc.SK * k] = cw[t];
}

struct QuadData { double xq[9], wq[9]; };

// cw[s][q] = a(x_q) * (1/h)^2 * h^dim * w_q, a = 1/(0.05 + 2|x|^2)
// (LocalCoeffOp + Coefficient::value, laplace_operator_gpu.h:191-211, poisson_common.h:146-158)
template <typename Number>
__global__ void eval_cw_uniform(MortonMap mm, int dim, int n, uint32_t npc, uint32_t n_cells, const uint32_t *__restrict__ perm,
                                double ox, double oy, double oz, double h, QuadData qd, Number *__restrict__ cw)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n_cells * npc) return;
  const uint32_t s = (uint32_t)(t / npc), q = (uint32_t)(t % npc);
  const uint32_t c = perm ? perm[s] : s;
  uint32_t x[3]; mm.decode(c, x);
  const int qi[3] = {(int)(q % n), (int)((q / n) % n), (int)(q / (n * n))};
  const double o[3] = {ox, oy, oz};
  double r2 = 0, w = 1;
  for (int d = 0; d < dim; ++d)
    {
      const double X = o[d] + h * ((double)x[d] + qd.xq[qi[d]]);
      r2 += X * X; w *= h * qd.wq[qi[d]];
    }
  const double a = 1.0 / (0.05 + 2.0 * r2);
  cw[t] = (Number)(a * w / (h * h));
}

struct DiagTables { double val[81], grad[81], hang[81]; };

// compute_diagonal (laplace_operator_gpu.h:355-421): diag_i = sum_q cw_q sum_d (dphi_i/dxi_d (q))^2.
// The reference applies the cell operator to every local unit vector (O(n^(2dim+1)) per cell); the same number is
// obtained here in closed form.  Like the reference's DiagonalLocalOperator, cells with hanging nodes pass their
// local diagonal through the TRANSPOSED hanging-node interpolation before it is scattered (:397-399).
template <typename Number>
__global__ void diagonal_kernel(const uint32_t *__restrict__ idx, const Number *__restrict__ cw, int dim, int n, uint32_t npc,
                                uint32_t n_cells, DiagTables tb, Number *__restrict__ diag, const uint32_t *__restrict__ hn_mask)
{
  extern __shared__ double dsm[];  // 2 * npc
  double *a = dsm, *b = dsm + npc;
  const uint32_t cell = blockIdx.x;
  if (cell >= n_cells) return;
  const unsigned mask = hn_mask ? hn_mask[cell] : 0u;
  const int p = n - 1;
  for (uint32_t i = threadIdx.x; i < npc; i += blockDim.x)
    {
      const int ii[3] = {(int)(i % n), (int)((i / n) % n), (int)(i / (n * n))};
      double acc = 0;
      for (uint32_t q = 0; q < npc; ++q)
        {
          const int qi[3] = {(int)(q % n), (int)((q / n) % n), (int)(q / (n * n))};
          double v[3], g2[3];
          for (int d = 0; d < dim; ++d)
            {
              const double x = tb.val[ii[d] * n + qi[d]], y = tb.grad[ii[d] * n + qi[d]];
              v[d] = x * x; g2[d] = y * y;
            }
          double s;
          if (dim == 2) s = g2[0] * v[1] + v[0] * g2[1];
          else s = g2[0] * v[1] * v[2] + v[0] * g2[1] * v[2] + v[0] * v[1] * g2[2];
          acc += s * (double)cw[(size_t)cell * npc + q];
        }
      a[i] = acc;
    }
  __syncthreads();
  if (mask)
    for (int d = 0; d < dim; ++d)
      {
        for (uint32_t i = threadIdx.x; i < npc; i += blockDim.x)
          {
            const int c[3] = {(int)(i % n), (int)((i / n) % n), (int)(i / (n * n))};
            bool flag;
            if (dim == 2)
              {
                const int o = 1 - d;
                const bool on = (mask & (1u << o)) ? (c[o] == 0) : (c[o] == p);
                flag = (mask & (8u << o)) && on;
              }
            else
              {
                const int f1 = (d + 1) % 3, f2 = (d + 2) % 3;
                const bool on1 = (mask & (1u << f1)) ? (c[f1] == 0) : (c[f1] == p);
                const bool on2 = (mask & (1u << f2)) ? (c[f2] == 0) : (c[f2] == p);
                const unsigned ebit = d == 0 ? (1u << 7) : d == 1 ? (1u << 8) : (1u << 6);
                flag = ((mask & (8u << f1)) && on1) || ((mask & (8u << f2)) && on2) || ((mask & ebit) && on1 && on2);
              }
            double val = a[i];
            if (flag)
              {
                const int stride = d == 0 ? 1 : d == 1 ? n : n * n, k = c[d], base = (int)i - k * stride;
                const bool first = mask & (1u << d);
                double acc = 0;
                for (int j = 0; j < n; ++j) acc += (first ? tb.hang[j * n + k] : tb.hang[(p - j) * n + p - k]) * a[base + j * stride];
                val = acc;
              }
            b[i] = val;
          }
        __syncthreads();
        double *t = a; a = b; b = t;
      }
  for (uint32_t i = threadIdx.x; i < npc; i += blockDim.x)
    {
      const uint32_t g = idx[(size_t)cell * npc + i];
      if (!(g & CONSTRAINED_BIT)) atomicAdd(diag + g, (Number)a[i]);
    }
}

template <typename Number> __global__ void k_set(Number *v, const uint32_t *list, size_t n, Number val)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[list[i]] = val;
}
// save_constrained_dofs_kernel (constraint_handler_gpu.cu:233-244)
template <typename Number> __global__ void k_save(Number *in, Number *tmp_in, const uint32_t *list, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t c = list[i]; tmp_in[i] = in[c]; in[c] = 0; }
}
// (constraint_handler_gpu.cu:248-262)
template <typename Number> __global__ void k_save2(const Number *out, Number *in, Number *tmp_out, Number *tmp_in, const uint32_t *list, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t c = list[i]; tmp_out[i] = out[c]; tmp_in[i] = in[c]; in[c] = 0; }
}
// (constraint_handler_gpu.cu:264-274)
template <typename Number> __global__ void k_load(Number *in, const Number *tmp_in, const uint32_t *list, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) in[list[i]] = tmp_in[i];
}
// (constraint_handler_gpu.cu:277-289)
template <typename Number> __global__ void k_load_add(Number *out, Number *in, const Number *tmp_out, const Number *tmp_in, const uint32_t *list, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t c = list[i]; out[c] = tmp_out[i] + tmp_in[i]; in[c] = tmp_in[i]; }
}

// (at least one block: every kernel launched with it checks its index, and an empty partition must not fail the launch)
inline unsigned nblk(size_t n, int t = 256) { return (unsigned)std::max<size_t>(1, (n + t - 1) / t); }

}  // namespace

// ---------------------------------------------------------------------------
// coloring
// ---------------------------------------------------------------------------
// On the structured box every cell conflicts only with its 3^dim-1 neighbours,
// so the 2^dim parity classes are a valid (and minimal) coloring.  The
// reference calls deal.II's GraphColoring::make_graph_coloring (coloring.cc:20-33);
// any valid coloring yields the same operator up to summation order.
void mesh_parity_colors(const mfg_mesh *m, std::vector<uint32_t> &color_of_cell, uint32_t &n_colors)
{
  std::vector<uint32_t> xyz((size_t)m->n_cells * 3);
  mesh_cell_coords(m, xyz.data());
  color_of_cell.resize(m->n_cells);
  uint32_t used[8] = {0};
  for (uint32_t c = 0; c < m->n_cells; ++c)
    {
      uint32_t col = 0;
      for (int d = 0; d < m->dim; ++d) col |= (xyz[3 * (size_t)c + d] & 1u) << d;
      color_of_cell[c] = col; used[col] = 1;
    }
  // compress to the colors actually used (e.g. a single cell -> one color)
  uint32_t remap[8], k = 0;
  for (int i = 0; i < 8; ++i) remap[i] = used[i] ? k++ : 0;
  for (auto &c : color_of_cell) c = remap[c];
  n_colors = k;
}

// ---------------------------------------------------------------------------
// MatrixFreeGpu
// ---------------------------------------------------------------------------
mfg_mf *mf_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter)
{
  std::unique_ptr<mfg_mf> mf(new mfg_mf);
  mf->ctx = ctx; mf->dim = mesh->dim; mf->p = mesh->p; mf->n = mesh->n; mf->npc = mesh->npc;
  mf->n_cells = mesh->n_cells; mf->n_dofs = mesh->n_dofs; mf->dt = dt; mf->scatter = scatter; mf->fe = mesh->fe; mf->mesh = mesh;
  cudaStream_t s = ctx->stream;
  if (scatter == MFG_SCATTER_COLOR)
    {
      std::vector<uint32_t> col; uint32_t ncol = 0;
      mesh_parity_colors(mesh, col, ncol);
      std::vector<uint32_t> perm(mesh->n_cells);
      std::iota(perm.begin(), perm.end(), 0u);
      std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return col[a] < col[b]; });
      mf->color_offsets.assign(ncol + 1, 0);
      for (uint32_t c = 0; c < mesh->n_cells; ++c) mf->color_offsets[col[c] + 1]++;
      for (uint32_t k = 0; k < ncol; ++k) mf->color_offsets[k + 1] += mf->color_offsets[k];
      mf->cell_perm.upload(perm.data(), perm.size(), s);
    }
  else mf->color_offsets = {0u, mesh->n_cells};
  const size_t total = (size_t)mesh->n_cells * mesh->npc;
  mf->idx.alloc(total);
  build_kernel_indices<<<nblk(total), 256, 0, s>>>(mesh->l2g.p, mf->cell_perm.p, mesh->cflag.p, mesh->npc, total, mf->idx.p);
  MFG_CUDA_LAST();
  return mf.release();
}

mfg_mf *mf_from_desc(mfg_ctx *ctx, const mfg_mf_desc &d)
{
  MFG_REQUIRE(d.dim == 2 || d.dim == 3, "dim must be 2 or 3");
  MFG_REQUIRE(d.degree >= 1 && d.degree <= 8, "degree must be in 1..8");
  MFG_REQUIRE(d.loc2glob != nullptr && d.inv_jac != nullptr, "loc2glob and inv_jac are required");
  MFG_REQUIRE(d.geometry == MFG_GEOM_UNIFORM || d.geometry == MFG_GEOM_GENERAL, "unknown geometry kind");
  if (d.geometry == MFG_GEOM_GENERAL)
    {
      MFG_REQUIRE(d.JxW != nullptr, "general geometry needs JxW per quadrature point");
      if (d.constraint_mask) throw Error(MFG_ERR_UNSUPPORTED, "hanging nodes with general geometry are not implemented");
    }
  MFG_REQUIRE(d.n_dofs < CONSTRAINED_BIT, "n_dofs must be < 2^31");
  std::unique_ptr<mfg_mf> mf(new mfg_mf);
  mf->ctx = ctx; mf->dim = d.dim; mf->p = d.degree; mf->n = d.degree + 1; mf->npc = ipow(mf->n, mf->dim);
  mf->n_cells = d.n_cells; mf->n_dofs = d.n_dofs; mf->dt = d.dtype; mf->scatter = d.scatter; mf->fe = make_fe_data(d.degree);
  if (d.scatter == MFG_SCATTER_COLOR)
    {
      MFG_REQUIRE(d.n_colors >= 1 && d.color_offsets != nullptr, "coloring needs n_colors and color_offsets");
      mf->color_offsets.assign(d.color_offsets, d.color_offsets + d.n_colors + 1);
      MFG_REQUIRE(mf->color_offsets.front() == 0 && mf->color_offsets.back() == d.n_cells, "color_offsets must span all cells");
    }
  else mf->color_offsets = {0u, d.n_cells};
  const size_t total = (size_t)d.n_cells * mf->npc;
  for (size_t i = 0; i < total; ++i) MFG_REQUIRE(d.loc2glob[i] < d.n_dofs, "loc2glob entry out of range");
  if (d.constraint_mask)
    {
      // hanging nodes: cells without constraints first (they take the fast kernels), constrained cells last
      MFG_REQUIRE(d.scatter == MFG_SCATTER_ATOMIC, "hanging nodes need the atomic scatter");
      std::vector<uint32_t> perm(d.n_cells);
      std::iota(perm.begin(), perm.end(), 0u);
      std::stable_partition(perm.begin(), perm.end(), [&](uint32_t c) { return d.constraint_mask[c] == 0; });
      uint32_t n_plain = 0;
      for (uint32_t c = 0; c < d.n_cells; ++c) n_plain += d.constraint_mask[c] == 0;
      std::vector<uint32_t> l2g_perm(total), mask_perm(d.n_cells);
      for (uint32_t s = 0; s < d.n_cells; ++s)
        {
          std::copy(d.loc2glob + (size_t)perm[s] * mf->npc, d.loc2glob + (size_t)(perm[s] + 1) * mf->npc, l2g_perm.begin() + (size_t)s * mf->npc);
          mask_perm[s] = d.constraint_mask[perm[s]];
          MFG_REQUIRE(mask_perm[s] < 512u, "constraint mask has more than 9 bits");
        }
      mf->n_plain = n_plain;
      mf->color_offsets = {0u, n_plain};
      mf->idx.upload(l2g_perm.data(), total, ctx->stream);
      mf->hn_mask.upload(mask_perm.data(), d.n_cells, ctx->stream);
      mf->cell_perm.upload(perm.data(), d.n_cells, ctx->stream);
    }
  else
    mf->idx.upload(d.loc2glob, total, ctx->stream);
  if (d.JxW) mf->jxw_host.assign(d.JxW, d.JxW + total);
  mf->invjac_host.assign(d.inv_jac, d.inv_jac + (d.geometry == MFG_GEOM_GENERAL ? total * d.dim * d.dim : (size_t)d.n_cells));
  if (d.geometry == MFG_GEOM_GENERAL)
    {
      // G = JxW K K^T with K = J^-1 [cell][q][d1][d2] as FEValues::get_inverse_jacobians delivers it (matrix_free_gpu.cu:326-338):
      // get_gradient applies K^T (fee_gpu.cuh:236-240), submit_gradient K and JxW (:276-280)
      const int dim = mf->dim, nc = dim * (dim + 1) / 2;
      static const int pairs3[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {0, 2}, {1, 2}}, pairs2[3][2] = {{0, 0}, {1, 1}, {0, 1}};
      mf->general = true;
      mf->gsym_host.resize(total * nc);
      for (uint32_t c = 0; c < d.n_cells; ++c)
        for (uint32_t q = 0; q < mf->npc; ++q)
          {
            const double *K = d.inv_jac + ((size_t)c * mf->npc + q) * dim * dim;
            const double  jxw = d.JxW[(size_t)c * mf->npc + q];
            for (int k = 0; k < nc; ++k)
              {
                const int a = dim == 3 ? pairs3[k][0] : pairs2[k][0], b = dim == 3 ? pairs3[k][1] : pairs2[k][1];
                double    s = 0;
                for (int e = 0; e < dim; ++e) s += K[a * dim + e] * K[b * dim + e];
                mf->gsym_host[((size_t)c * nc + k) * mf->npc + q] = jxw * s;
              }
          }
      if (d.quadrature_points) mf->qpoints_host.assign(d.quadrature_points, d.quadrature_points + total * mf->dim);
      return mf.release();
    }
  // merged geometry factor inv_jac^2 * JxW_q  (fee_gpu.cuh:228,270: grad = J0*g ; submit = grad*J0*jxw)
  mf->geom_host.resize(total);
  for (uint32_t c = 0; c < d.n_cells; ++c)
    for (uint32_t q = 0; q < mf->npc; ++q)
      {
        double jxw;
        if (d.JxW) jxw = d.JxW[(size_t)c * mf->npc + q];
        else
          {
            jxw = 1.0;
            uint32_t qq = q;
            for (int k = 0; k < mf->dim; ++k) { jxw *= mf->fe.qwts[qq % mf->n] / d.inv_jac[c]; qq /= mf->n; }
          }
        mf->geom_host[(size_t)c * mf->npc + q] = d.inv_jac[c] * d.inv_jac[c] * jxw;
      }
  if (d.quadrature_points) mf->qpoints_host.assign(d.quadrature_points, d.quadrature_points + total * mf->dim);
  return mf.release();
}

// ---------------------------------------------------------------------------
// generic FEEvaluationGpu path: the per-cell arrays of MatrixFreeGpu::GpuData (matrix_free_gpu.h:261-278) on the device
// ---------------------------------------------------------------------------
template <typename Number>
__global__ void fill_uniform_geometry(MortonMap mm, int dim, int n, uint32_t npc, uint32_t n_cells, const uint32_t *__restrict__ perm, double ox,
                                      double oy, double oz, double h, QuadData qd, Number *__restrict__ jxw, Number *__restrict__ invjac,
                                      Number *__restrict__ qpts)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n_cells * npc) return;
  const uint32_t s = (uint32_t)(t / npc), q = (uint32_t)(t % npc);
  const uint32_t c = perm ? perm[s] : s;
  uint32_t x[3]; mm.decode(c, x);
  const int qi[3] = {(int)(q % n), (int)((q / n) % n), (int)(q / (n * n))};
  const double o[3] = {ox, oy, oz};
  double w = 1;
  for (int d = 0; d < dim; ++d)
    {
      qpts[t * dim + d] = (Number)(o[d] + h * ((double)x[d] + qd.xq[qi[d]]));
      w *= h * qd.wq[qi[d]];
    }
  jxw[t] = (Number)w;
  if (q == 0) invjac[s] = (Number)(1.0 / h);
}

void mf_get_gpu_data(mfg_mf *mf, mfg_gpu_data *out)
{
  cudaStream_t s = mf->ctx->stream;
  const size_t total = (size_t)mf->n_cells * mf->npc, es = mf->dt == MFG_F64 ? 8 : 4;
  if (mf->gd_jxw.n == 0 && total)
    {
      if (mf->mesh)
        {
          const mfg_mesh *mesh = mf->mesh;
          mf->gd_jxw.alloc(total * es); mf->gd_invjac.alloc((size_t)mf->n_cells * es); mf->gd_qpts.alloc(total * mf->dim * es);
          QuadData qd;
          for (int i = 0; i < mesh->n; ++i) { qd.xq[i] = mesh->fe.qpts[i]; qd.wq[i] = mesh->fe.qwts[i]; }
          MortonMap mm; mm.dim = mesh->dim; for (int d = 0; d < 3; ++d) mm.lg[d] = mesh->lg[d];
          if (mf->dt == MFG_F64)
            fill_uniform_geometry<double><<<nblk(total), 256, 0, s>>>(mm, mf->dim, mf->n, mf->npc, mf->n_cells, mf->cell_perm.p, mesh->origin[0], mesh->origin[1],
                                                                      mesh->origin[2], mesh->h, qd, (double *)mf->gd_jxw.p, (double *)mf->gd_invjac.p, (double *)mf->gd_qpts.p);
          else
            fill_uniform_geometry<float><<<nblk(total), 256, 0, s>>>(mm, mf->dim, mf->n, mf->npc, mf->n_cells, mf->cell_perm.p, mesh->origin[0], mesh->origin[1],
                                                                     mesh->origin[2], mesh->h, qd, (float *)mf->gd_jxw.p, (float *)mf->gd_invjac.p, (float *)mf->gd_qpts.p);
          MFG_CUDA_LAST();
        }
      else
        {
          // arrays of mfg_mf_desc, permuted into the kernel cell order (hanging-node cells last) and converted to Number
          MFG_REQUIRE(!mf->invjac_host.empty(), "mf has no geometry");
          std::vector<uint32_t> perm;
          if (mf->cell_perm.n) { perm.resize(mf->n_cells); mf->cell_perm.download(perm.data(), s); }
          const size_t per_inv = mf->general ? (size_t)mf->npc * mf->dim * mf->dim : 1;
          std::vector<double> jxw(total), inv((size_t)mf->n_cells * per_inv), qp;
          if (!mf->qpoints_host.empty()) qp.resize(total * mf->dim);
          for (uint32_t k = 0; k < mf->n_cells; ++k)
            {
              const uint32_t c = perm.empty() ? k : perm[k];
              for (uint32_t q = 0; q < mf->npc; ++q)
                {
                  double w;
                  if (!mf->jxw_host.empty()) w = mf->jxw_host[(size_t)c * mf->npc + q];
                  else
                    {
                      w = 1.0; uint32_t qq = q;
                      for (int d = 0; d < mf->dim; ++d) { w *= mf->fe.qwts[qq % mf->n] / mf->invjac_host[c]; qq /= mf->n; }
                    }
                  jxw[(size_t)k * mf->npc + q] = w;
                }
              std::copy(mf->invjac_host.begin() + (size_t)c * per_inv, mf->invjac_host.begin() + (size_t)(c + 1) * per_inv, inv.begin() + (size_t)k * per_inv);
              if (!qp.empty())
                std::copy(mf->qpoints_host.begin() + (size_t)c * mf->npc * mf->dim, mf->qpoints_host.begin() + (size_t)(c + 1) * mf->npc * mf->dim,
                          qp.begin() + (size_t)k * mf->npc * mf->dim);
            }
          auto up = [&](DevBuf<uint8_t> &b, const std::vector<double> &v) {
            if (v.empty()) return;
            b.alloc(v.size() * es);
            if (mf->dt == MFG_F64) MFG_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * 8, cudaMemcpyHostToDevice, s));
            else
              {
                std::vector<float> f(v.begin(), v.end());
                MFG_CUDA(cudaMemcpyAsync(b.p, f.data(), f.size() * 4, cudaMemcpyHostToDevice, s));
                MFG_CUDA(cudaStreamSynchronize(s));
              }
          };
          up(mf->gd_jxw, jxw); up(mf->gd_invjac, inv); up(mf->gd_qpts, qp);
        }
      MFG_CUDA(cudaStreamSynchronize(s));
    }
  std::memset(out, 0, sizeof(*out));
  out->loc2glob = mf->idx.p;
  out->JxW = mf->gd_jxw.p; out->inv_jac = mf->gd_invjac.p; out->quadrature_points = mf->gd_qpts.n ? mf->gd_qpts.p : nullptr;
  out->n_cells = mf->n_cells; out->n_dofs = mf->n_dofs; out->dim = mf->dim; out->degree = mf->p; out->dtype = mf->dt;
  out->general = mf->general ? 1 : 0;
  out->use_coloring = mf->scatter == MFG_SCATTER_COLOR ? 1 : 0;
  out->n_colors = mf->n_colors();
  out->color_offsets = mf->color_offsets.data();
  out->n_plain_cells = mf->hn_mask.n ? mf->n_plain : mf->n_cells;
  out->constraint_mask = mf->hn_mask.n ? mf->hn_mask.p : nullptr;
  out->cuda_stream = (void *)s;
  for (int i = 0; i < mf->n * mf->n; ++i) { out->shape_values[i] = mf->fe.val[i]; out->shape_gradients[i] = mf->fe.grad[i]; out->colloc_gradients[i] = mf->fe.colloc[i]; }
}

// ---------------------------------------------------------------------------
// ConstraintHandlerGpu
// ---------------------------------------------------------------------------
mfg_ch *ch_create(mfg_ctx *ctx, mfg_dtype dt, const uint32_t *constrained_host, size_t nc, const uint32_t *edge_host, size_t ne)
{
  std::unique_ptr<mfg_ch> ch(new mfg_ch);
  ch->ctx = ctx; ch->dt = dt;
  ch->constrained.upload(constrained_host, nc, ctx->stream);
  ch->edge.upload(edge_host, ne, ctx->stream);
  const size_t es = dt == MFG_F64 ? 8 : 4;
  ch->tmp_src.alloc(nc * es); ch->tmp_dst.alloc(nc * es);
  return ch.release();
}

mfg_ch *ch_from_mesh(mfg_ctx *ctx, mfg_dtype dt, const mfg_mesh *mesh)
{
  std::unique_ptr<mfg_ch> ch(new mfg_ch);
  ch->ctx = ctx; ch->dt = dt;
  ch->constrained.alloc(mesh->n_constrained);
  if (mesh->n_constrained)
    MFG_CUDA(cudaMemcpyAsync(ch->constrained.p, mesh->constrained.p, mesh->n_constrained * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  const size_t es = dt == MFG_F64 ? 8 : 4;
  ch->tmp_src.alloc(mesh->n_constrained * es); ch->tmp_dst.alloc(mesh->n_constrained * es);
  return ch.release();
}

#define CH_DISPATCH(ch, CALL)                               \
  do {                                                      \
    if ((ch)->dt == MFG_F64) { typedef double T; CALL; }    \
    else { typedef float T; CALL; }                         \
    MFG_CUDA_LAST();                                        \
  } while (0)

static void ch_check(const mfg_ch *ch, const mfg_vec *v) { MFG_REQUIRE(v->dt == ch->dt, "vector dtype differs from constraint handler dtype"); }

void ch_set(mfg_ch *ch, mfg_vec *v, double val)
{
  ch_check(ch, v); const size_t n = ch->n(); if (!n) return;
  CH_DISPATCH(ch, (k_set<T><<<nblk(n, 128), 128, 0, ch->ctx->stream>>>((T *)v->p, ch->constrained.p, n, (T)val)));
}
void ch_save(mfg_ch *ch, mfg_vec *v)
{
  ch_check(ch, v); const size_t n = ch->n(); if (!n) return;
  CH_DISPATCH(ch, (k_save<T><<<nblk(n, 128), 128, 0, ch->ctx->stream>>>((T *)v->p, (T *)ch->tmp_src.p, ch->constrained.p, n)));
}
void ch_save2(mfg_ch *ch, const mfg_vec *v1, mfg_vec *v2)
{
  ch_check(ch, v1); ch_check(ch, v2); const size_t n = ch->n(); if (!n) return;
  CH_DISPATCH(ch, (k_save2<T><<<nblk(n, 128), 128, 0, ch->ctx->stream>>>((const T *)v1->p, (T *)v2->p, (T *)ch->tmp_dst.p, (T *)ch->tmp_src.p, ch->constrained.p, n)));
}
void ch_load(mfg_ch *ch, mfg_vec *v)
{
  ch_check(ch, v); const size_t n = ch->n(); if (!n) return;
  CH_DISPATCH(ch, (k_load<T><<<nblk(n, 128), 128, 0, ch->ctx->stream>>>((T *)v->p, (const T *)ch->tmp_src.p, ch->constrained.p, n)));
}
void ch_load_and_add(mfg_ch *ch, mfg_vec *v1, mfg_vec *v2)
{
  ch_check(ch, v1); ch_check(ch, v2); const size_t n = ch->n(); if (!n) return;
  CH_DISPATCH(ch, (k_load_add<T><<<nblk(n, 128), 128, 0, ch->ctx->stream>>>((T *)v1->p, (T *)v2->p, (const T *)ch->tmp_dst.p, (const T *)ch->tmp_src.p, ch->constrained.p, n)));
}

// ---------------------------------------------------------------------------
// LaplaceOperatorGpu
// ---------------------------------------------------------------------------
static void laplace_finish_setup(mfg_laplace *op, const uint8_t *cflag_dev)
{
  // bitmask of constrained DoFs for the fused `dst = 0 / dst[c] = src[c]` pass
  const size_t nd = op->mf->n_dofs;
  op->cbits.alloc((nd + 31) / 32);
  pack_bits<<<nblk((nd + 31) / 32), 256, 0, op->ctx->stream>>>(cflag_dev, nd, op->cbits.p);
  MFG_CUDA_LAST();
}

mfg_laplace *laplace_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_scatter scatter)
{
  std::unique_ptr<mfg_laplace> op(new mfg_laplace);
  op->ctx = ctx;
  op->mf = mf_from_mesh(ctx, mesh, dt, scatter); op->owns_mf = true;
  op->ch = ch_from_mesh(ctx, dt, mesh); op->owns_ch = true;
  laplace_finish_setup(op.get(), mesh->cflag.p);
  // evaluate_coefficient (laplace_operator_gpu.h:204-211) fused with the geometry factors
  const size_t total = (size_t)mesh->n_cells * mesh->npc;
  const size_t es = dt == MFG_F64 ? 8 : 4;
  op->cw.alloc((total + 32 * (size_t)mesh->npc) * es);  // padded with zero cells: the slab kernel copies whole groups
  MFG_CUDA(cudaMemsetAsync(op->cw.p, 0, op->cw.bytes(), ctx->stream));
  QuadData qd;
  for (int i = 0; i < mesh->n; ++i) { qd.xq[i] = mesh->fe.qpts[i]; qd.wq[i] = mesh->fe.qwts[i]; }
  MortonMap mm; mm.dim = mesh->dim; for (int d = 0; d < 3; ++d) mm.lg[d] = mesh->lg[d];
  if (dt == MFG_F64)
    eval_cw_uniform<double><<<nblk(total), 256, 0, ctx->stream>>>(mm, mesh->dim, mesh->n, mesh->npc, mesh->n_cells, op->mf->cell_perm.p,
                                                                  mesh->origin[0], mesh->origin[1], mesh->origin[2], mesh->h, qd, (double *)op->cw.p);
  else
    eval_cw_uniform<float><<<nblk(total), 256, 0, ctx->stream>>>(mm, mesh->dim, mesh->n, mesh->npc, mesh->n_cells, op->mf->cell_perm.p,
                                                                 mesh->origin[0], mesh->origin[1], mesh->origin[2], mesh->h, qd, (float *)op->cw.p);
  MFG_CUDA_LAST();
  MFG_CUDA(cudaStreamSynchronize(ctx->stream));
  return op.release();
}

mfg_laplace *laplace_from_arrays(mfg_ctx *ctx, mfg_mf *mf, mfg_ch *ch, const double *coef_host)
{
  MFG_REQUIRE(mf && ch && coef_host, "mf, ch and coefficient are required");
  MFG_REQUIRE(mf->dt == ch->dt, "mf and ch dtypes differ");
  MFG_REQUIRE(mf->n_cells == 0 || !mf->geom_host.empty() || !mf->gsym_host.empty() || mf->mesh, "mf has no geometry");
  std::unique_ptr<mfg_laplace> op(new mfg_laplace);
  op->ctx = ctx; op->mf = mf; op->ch = ch;
  // mark constrained DoFs in the kernel index array (in place)
  DevBuf<uint8_t> flag(mf->n_dofs);
  MFG_CUDA(cudaMemsetAsync(flag.p, 0, mf->n_dofs, ctx->stream));
  if (ch->n()) { flags_from_list<<<nblk(ch->n()), 256, 0, ctx->stream>>>(ch->constrained.p, ch->n(), flag.p); MFG_CUDA_LAST(); }
  const size_t total = (size_t)mf->n_cells * mf->npc;
  build_kernel_indices<<<nblk(total), 256, 0, ctx->stream>>>(mf->idx.p, nullptr, flag.p, mf->npc, total, mf->idx.p);
  MFG_CUDA_LAST();
  laplace_finish_setup(op.get(), flag.p);
  op->cw.alloc((total * (mf->general ? mf->dim * (mf->dim + 1) / 2 : 1) + 32 * (size_t)mf->npc) * (mf->dt == MFG_F64 ? 8 : 4));
  MFG_CUDA(cudaMemsetAsync(op->cw.p, 0, op->cw.bytes(), ctx->stream));
  laplace_set_coefficient_host(op.get(), coef_host);
  MFG_CUDA(cudaStreamSynchronize(ctx->stream));
  return op.release();
}

// coefficient values a(x_q), [n_cells][npc], ORIGINAL cell order -> merged weights in kernel cell order
void laplace_set_coefficient_host(mfg_laplace *op, const double *coef_host)
{
  const mfg_mf *mf = op->mf;
  const size_t total = (size_t)mf->n_cells * mf->npc;
  if (total == 0) { op->diagonal_is_available = false; op->cwP_valid = false; return; }  // empty partition
  if (mf->general)
    {
      // merged tensor a(x_q) JxW K K^T, [cell][component][q]
      const int nc = mf->dim * (mf->dim + 1) / 2;
      std::vector<double> merged(total * nc);
      for (uint32_t c = 0; c < mf->n_cells; ++c)
        for (int k = 0; k < nc; ++k)
          for (uint32_t q = 0; q < mf->npc; ++q)
            merged[((size_t)c * nc + k) * mf->npc + q] = coef_host[(size_t)c * mf->npc + q] * mf->gsym_host[((size_t)c * nc + k) * mf->npc + q];
      if (mf->dt == MFG_F64) MFG_CUDA(cudaMemcpyAsync(op->cw.p, merged.data(), merged.size() * 8, cudaMemcpyHostToDevice, op->ctx->stream));
      else
        {
          std::vector<float> m32(merged.begin(), merged.end());
          MFG_CUDA(cudaMemcpyAsync(op->cw.p, m32.data(), m32.size() * 4, cudaMemcpyHostToDevice, op->ctx->stream));
          MFG_CUDA(cudaStreamSynchronize(op->ctx->stream));
        }
      MFG_CUDA(cudaStreamSynchronize(op->ctx->stream));
      op->diagonal_is_available = false;
      return;
    }
  std::vector<uint32_t> perm;
  if (mf->cell_perm.n) { perm.resize(mf->n_cells); mf->cell_perm.download(perm.data(), op->ctx->stream); }
  std::vector<double> geom_q;  // uniform mesh: same factors for every cell
  if (mf->geom_host.empty())
    {
      geom_q.resize(mf->npc);
      const double h = mf->mesh->h;
      for (uint32_t q = 0; q < mf->npc; ++q)
        {
          double w = 1.0; uint32_t qq = q;
          for (int k = 0; k < mf->dim; ++k) { w *= h * mf->fe.qwts[qq % mf->n]; qq /= mf->n; }
          geom_q[q] = w / (h * h);
        }
    }
  std::vector<double> merged(total);
  for (uint32_t s = 0; s < mf->n_cells; ++s)
    {
      const uint32_t c = perm.empty() ? s : perm[s];
      for (uint32_t q = 0; q < mf->npc; ++q)
        merged[(size_t)s * mf->npc + q] = coef_host[(size_t)c * mf->npc + q] * (geom_q.empty() ? mf->geom_host[(size_t)c * mf->npc + q] : geom_q[q]);
    }
  if (mf->dt == MFG_F64)
    MFG_CUDA(cudaMemcpyAsync(op->cw.p, merged.data(), total * 8, cudaMemcpyHostToDevice, op->ctx->stream));
  else
    {
      std::vector<float> mf32(merged.begin(), merged.end());
      MFG_CUDA(cudaMemcpyAsync(op->cw.p, mf32.data(), total * 4, cudaMemcpyHostToDevice, op->ctx->stream));
      MFG_CUDA(cudaStreamSynchronize(op->ctx->stream));
    }
  MFG_CUDA(cudaStreamSynchronize(op->ctx->stream));
  op->diagonal_is_available = false;
  op->cwP_valid = false;
}

// (re)build the slab2 kernel's private arrays from idx / cw
// configuration of the slab2 kernel for the operator's variant; + 256 = the build with the face merge (configurations 3, 7)
static inline bool slab3_variant(int v) { return v >= 50 && v <= 54; }  // 50 = flavour chosen by measurement; 51..54: flavours 0..3

// flavour of the slab3 kernel: bit 0 = asynchronous gather, bit 1 = early face merges.  auto (variant 0 / 50): early merges
// except for degree 4 in FP64, no asynchronous gather (measured on B200, profiles/r02_slab3_flavours.txt: the merges gain
// 3-11 % for FP32 and for degree 2, 3; the cp.async gather costs 10-15 % everywhere); 51..54 select flavour 0..3
static int slab3_flavour(const mfg_laplace *op)
{
  if (op->variant >= 51 && op->variant <= 54) return op->variant - 51;
  return (op->mf->p == 4 && op->mf->dt == MFG_F64) ? 0 : 2;
}

// face merges of the slab kernels: bit d = direction d, bit 3 = order x, z, y (slab3, staged); the staged kernel keeps its
// own masks in its plan and runs its left-over groups on the slab3 kernel without merges
static int slab2_merge_dirs(const mfg_laplace *op) { return laplace_active_variant(op) == 50 && (slab3_flavour(op) & 2) ? 8 + 7 : 0; }

static void laplace_prepare_slab2(mfg_laplace *op, uint32_t n_plain)
{
  const mfg_mf   *mf = op->mf;
  const Slab2Geom gm = slab2_geom(mf->p, mf->dt);
  const uint32_t  n_groups = (n_plain + gm.cw - 1) / gm.cw;
  cudaStream_t    s = op->ctx->stream;
  const int       ns = mf->n * mf->n;
  if (op->idxP.n != (size_t)n_groups * ns * 32 || op->slab2_groups != n_groups)
    {
      op->idxP.alloc((size_t)n_groups * ns * 32);
      if (n_groups) build_slab2_indices<<<nblk(op->idxP.n), 256, 0, s>>>(mf->idx.p, n_plain, n_groups, mf->n, gm, op->idxP.p);
      MFG_CUDA_LAST();
      op->slab2_groups = n_groups;
      op->cwP_valid = false;
      op->merge_dirs_built = -1;
    }
  const int dirs = slab2_merge_dirs(op);
  if (op->merge_dirs_built != dirs)
    {
      if (op->mergeP.n != n_groups) op->mergeP.alloc(n_groups);
      if (n_groups) build_slab2_merge<<<nblk(n_groups), 256, 0, s>>>(mf->idx.p, n_plain, n_groups, mf->n, gm.cw, dirs & 7, op->mergeP.p, (dirs & 8) != 0);
      MFG_CUDA_LAST();
      op->merge_dirs_built = dirs;
    }
  if (!op->cwP_valid)
    {
      const size_t es = mf->dt == MFG_F64 ? 8 : 4;
      if (op->cwP.n != (size_t)n_groups * gm.cwf * es) op->cwP.alloc((size_t)n_groups * gm.cwf * es);
      if (n_groups)
        {
          MFG_CUDA(cudaMemsetAsync(op->cwP.p, 0, op->cwP.bytes(), s));
          const size_t total = (size_t)n_plain * mf->npc;
          if (mf->dt == MFG_F64) build_slab2_weights<double><<<nblk(total), 256, 0, s>>>((const double *)op->cw.p, n_plain, mf->n, gm, (double *)op->cwP.p);
          else build_slab2_weights<float><<<nblk(total), 256, 0, s>>>((const float *)op->cw.p, n_plain, mf->n, gm, (float *)op->cwP.p);
          MFG_CUDA_LAST();
        }
      op->cwP_valid = true;
    }
}

// grouped kernels (one warp per group of 32 / n cells, work list, programmatic dependent launch): slab2 and staged
static inline bool grouped_variant(int v) { return v == 40 || v == 50; }

// plan of the staged kernel (stage_plan.cu), built once per operator from the index array
static void laplace_prepare_stage(mfg_laplace *op, uint32_t n_plain)
{
  laplace_prepare_slab2(op, n_plain);  // coefficient images (shared layout), index rows for the groups the plan leaves over
  if (op->st_built) return;
  const mfg_mf   *mf = op->mf;
  const StageGeom sg = stage_geom(mf->p, mf->dt);
  cudaStream_t    s = op->ctx->stream;
  std::vector<uint32_t> idx((size_t)mf->n_cells * mf->npc);
  mf->idx.download(idx.data(), s);
  StagePlanIn in;
  in.n = sg.n; in.cw = sg.cw; in.hc = sg.hc; in.wb = mf->dt == MFG_F64 ? 8 : 4; in.xcap = sg.xcap; in.hmax = sg.hmax; in.ocap = sg.ocap; in.lcap = sg.lcap; in.nclass = sg.nclass;
  in.n_plain = n_plain; in.n_cells = mf->n_cells; in.n_dofs = mf->n_dofs; in.idx = idx.data();
  in.merge_dirs = op->stage_merge_dirs & 7;
  StagePlan plan;
  build_stage_plan(in, plan);
  op->st_gdesc.upload(plan.gdesc.data(), plan.gdesc.size(), s);
  op->st_halo.upload(plan.halo.data(), plan.halo.size(), s);
  op->st_ptab.upload(plan.ptab.data(), plan.ptab.size(), s);
  op->st_fallback.upload(plan.fallback.data(), plan.fallback.size(), s);
  op->st_fb_iface = 0;
  op->st_pstride = plan.pstride;
  MFG_REQUIRE(plan.pstride == sg.pstride, "stage plan: table layout differs from the kernel's");
  std::copy(plan.class_pat, plan.class_pat + 8, op->st_class_pat);
  const uint64_t ns = std::max<uint32_t>(1, plan.n_staged);
  const uint32_t st[8] = {plan.n_groups, plan.n_staged, plan.n_patterns, (uint32_t)(16 * plan.n_own / ns), (uint32_t)(16 * plan.n_halo / ns),
                          (uint32_t)(16 * plan.n_plain_dofs / ns), (uint32_t)(16 * plan.n_red_dofs / ns),
                          (uint32_t)(16 * (plan.rd_wavefronts + plan.wr_wavefronts + 2 * plan.cp_wavefronts) / ns)};
  std::copy(st, st + 8, op->st_stats);
  op->st_built = true;
}

// kernels of this library one vmult enqueues (the cudaMemsetAsync of dst is not counted)
int laplace_launches_per_vmult(const mfg_laplace *op)
{
  // with the slab2 kernel the zero pass is a kernel of this library too (zero_fill_pdl)
  const int  v = laplace_active_variant(op);
  const bool zero_kernel = grouped_variant(v) && op->mf->hn_mask.n == 0;
  // (the slab3 kernel copies the constrained rows itself when it runs behind the zero kernel)
  return (op->ch->n() && !(v == 50 && zero_kernel) ? 1 : 0) + (int)op->mf->n_colors() + (op->mf->hn_mask.n ? 1 : 0) + (zero_kernel ? 1 : 0) +
         (v == 40 && op->st_built && op->st_fallback.n ? 1 : 0);
}

// kernel variants: 1 = column kernel (kernels_v0.cuh: every dim / degree / dtype / scatter, hanging-node cells),
//                  40 = staged kernel (kernels_stage.cuh: 3D, degree 2..5, atomic scatter),
//                  50 = slab3 kernel (kernels_slab3.cuh: 3D, degree <= 5, atomic scatter), 51..54 = its flavours 0..3.
//                  0 = auto: slab3 where it exists, else the column kernel (general geometry: its own kernel).
int laplace_active_variant(const mfg_laplace *op)
{
  const mfg_mf *mf = op->mf;
  const bool slab_ok = slab2_supported(mf->dim, mf->p, mf->dt) && mf->scatter == MFG_SCATTER_ATOMIC && !mf->general;
  const bool stage_ok = stage_supported(mf->dim, mf->p, mf->dt) && mf->scatter == MFG_SCATTER_ATOMIC && !mf->general;
  if (op->variant == 40)
    {
      if (!stage_ok) throw Error(MFG_ERR_UNSUPPORTED, "variant 40 (staged kernel) needs dim 3, degree 2..5, atomic scatter, uniform geometry");
      return 40;
    }
  if (slab3_variant(op->variant))
    {
      if (!slab_ok) throw Error(MFG_ERR_UNSUPPORTED, "variants 50..54 (slab3 kernel) need dim 3, degree <= 5, atomic scatter, uniform geometry");
      return 50;
    }
  if (op->variant != 0 && op->variant != 1) throw Error(MFG_ERR_UNSUPPORTED, "unknown kernel variant (0 auto, 1 column, 40 staged, 50..54 slab3)");
  if (mf->general || op->variant == 1) return 1;
  return slab_ok ? 50 : 1;
}

// dst = 0 as a kernel (16-byte stores) that releases its programmatic dependents at once: the cell kernel launched
// behind it loads, gathers and contracts its first groups while this one drains, and waits (griddepcontrol.wait)
// only before its first red
template <typename Number> __global__ void zero_fill_pdl(Number *__restrict__ dst, size_t n)
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int V = 16 / (int)sizeof(Number);
  const size_t  nv = n / V, stride = (size_t)gridDim.x * blockDim.x;
  float4 *d4 = reinterpret_cast<float4 *>(dst);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x == 0 && threadIdx.x < n - nv * V) dst[nv * V + threadIdx.x] = Number(0);
}

// part: -1 = the whole apply; 0 = zero/constraint pass; 1 = the cell groups that touch interface DoFs (multi-GPU: what
// the exchange waits for); 2 = the remaining groups.  Without an interface partition part 2 runs all cells, part 1 none.
// Parts 1 and 2 may run concurrently on different streams (they write disjoint... they only add into dst).
template <typename Number>
static void vmult_impl(mfg_laplace *op, Number *dst, const Number *src, bool add, int part = -1, cudaStream_t s_other = nullptr)
{
  const mfg_mf *mf = op->mf;
  cudaStream_t  s  = s_other ? s_other : op->ctx->stream;
  const int  av = laplace_active_variant(op);
  const bool split = part >= 1 && op->glist.n != 0 && grouped_variant(av);
  if (part == 1 && !split) return;
  // whole apply with the slab2 kernel: zero kernel -> cell kernel as its programmatic dependent -> constrained rows
  const bool pdl_fill = part == -1 && !add && grouped_variant(av) && mf->hn_mask.n == 0 &&
                        (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
  if (pdl_fill)
    {
      const unsigned nb = (unsigned)std::min<size_t>(((size_t)mf->n_dofs * sizeof(Number) / 16 + 255) / 256, (size_t)op->ctx->sm_count * 16);
      zero_fill_pdl<Number><<<std::max(1u, nb), 256, 0, s>>>(dst, mf->n_dofs);
      MFG_CUDA_LAST();
    }
  else if (part <= 0)
    {
  // vmult: dst = 0 (laplace_operator_gpu.h:221) fused with dst[c] = src[c];
  // vmult_add: dst[c] += src[c]  (load_and_add_constrained_values, :302)
  if (!add)
    {
      // (cudaMemsetAsync + a kernel over the constraint list measured 14 us faster per apply at 17 M DoFs than one fused
      // pass with a bit mask, profiles/r01_*)
      MFG_CUDA(cudaMemsetAsync(dst, 0, (size_t)mf->n_dofs * sizeof(Number), s));
      if (op->ch->n()) { constrained_copy<Number><<<nblk(op->ch->n()), 256, 0, s>>>(dst, src, op->ch->constrained.p, op->ch->n()); MFG_CUDA_LAST(); }
    }
  else if (op->ch->n())
    {
      constrained_add<Number><<<nblk(op->ch->n()), 256, 0, s>>>(dst, src, op->ch->constrained.p, op->ch->n());
      MFG_CUDA_LAST();
    }
    }
  if (part == 0) return;
  const bool     hanging = mf->hn_mask.n != 0;
  const uint32_t n_plain = hanging ? mf->n_plain : mf->n_cells;
  bool timed = false, fused_ccopy = false;
  auto time_begin = [&]() {
    timed = op->timing && (op->timing_counter++ % (size_t)op->timing_stride) == 0;
    if (!timed) return;
    if (op->ev_used + 2 > op->ev.size())
      for (int k = 0; k < 2; ++k) { cudaEvent_t e; MFG_CUDA(cudaEventCreate(&e)); op->ev.push_back(e); }
    MFG_CUDA(cudaEventRecord(op->ev[op->ev_used], s));
  };
  auto time_end = [&]() {
    if (timed) { MFG_CUDA(cudaEventRecord(op->ev[op->ev_used + 1], s)); op->ev_used += 2; }
  };
  const bool atomic = mf->scatter == MFG_SCATTER_ATOMIC;
  auto launch_v0 = [&](uint32_t c0, uint32_t c1, const uint32_t *mask) {
    if (mf->dim == 2)
      launch_laplace_v0_dim<2, Number>(mf->p, atomic, mf->idx.p, (const Number *)op->cw.p, src, dst, c0, c1, mf->fe.val.data(),
                                       mf->fe.colloc.data(), s, mask, mf->fe.hanging.data());
    else
      launch_laplace_v0_dim<3, Number>(mf->p, atomic, mf->idx.p, (const Number *)op->cw.p, src, dst, c0, c1, mf->fe.val.data(),
                                       mf->fe.colloc.data(), s, mask, mf->fe.hanging.data());
  };
  if (av == 40)
    {
      // staged kernel; the groups its plan leaves over run on the slab2 kernel (work list) right behind it
      laplace_prepare_stage(op, n_plain);
      const uint32_t *fl = op->st_fallback.p + (split && part == 2 ? op->st_fb_iface : 0);
      const uint32_t  nf = (uint32_t)(!split ? op->st_fallback.n : part == 2 ? op->st_fallback.n - op->st_fb_iface : op->st_fb_iface);
      // whole apply: every group; multi-GPU: the interface groups from the work list, then all groups without the interface flag
      const int mode = !split ? 0 : part == 1 ? 1 : 2;
      time_begin();
      launch_laplace_stage<Number>(mf->p, op->st_gdesc.p, op->st_halo.p, op->st_ptab.p, op->st_pstride, op->st_class_pat, (const Number *)op->cwP.p, src, dst,
                                   op->slab2_groups, mf->fe.val.data(), mf->fe.colloc.data(), op->ctx->sm_count, s, op->glist.p, op->n_iface_groups, mode,
                                   (split && part == 2) || pdl_fill, pdl_fill, add, op->ctx->device, op->stage_sync);
      if (nf)
        launch_laplace_slab3<Number>(mf->p, op->idxP.p, (const Number *)op->cwP.p, src, dst, nf, mf->fe.val.data(), mf->fe.colloc.data(), op->ctx->sm_count, s,
                                     op->mergeP.p, fl, false, 0, op->ctx->device, 0, nullptr, 0u);
      time_end();
    }
  else if (av == 50)
    {
      laplace_prepare_slab2(op, n_plain);
      const uint32_t *gl = split ? op->glist.p + (part == 2 ? op->n_iface_groups : 0) : nullptr;
      const uint32_t  ng = !split ? op->slab2_groups : part == 2 ? op->slab2_groups - op->n_iface_groups : op->n_iface_groups;
      time_begin();
      launch_laplace_slab3<Number>(mf->p, op->idxP.p, (const Number *)op->cwP.p, src, dst, ng, mf->fe.val.data(), mf->fe.colloc.data(), op->ctx->sm_count, s,
                                   op->mergeP.p, gl, (split && part == 2) || pdl_fill, pdl_fill ? 1 : 0, op->ctx->device, slab3_flavour(op),
                                   pdl_fill && ng ? op->ch->constrained.p : nullptr, pdl_fill && ng ? (uint32_t)op->ch->n() : 0u);
      fused_ccopy = pdl_fill && ng;  // (no groups: no kernel, the copy below runs)
      time_end();
    }
  else if (mf->general)
    {
      // full J^-1 per quadrature point (kernels_general.cuh), one launch per color
      for (uint32_t c = 0; c + 1 < mf->color_offsets.size(); ++c)
        {
          time_begin();
          if (mf->dim == 2)
            launch_laplace_general_dim<2, Number>(mf->p, atomic, mf->idx.p, (const Number *)op->cw.p, src, dst, mf->color_offsets[c], mf->color_offsets[c + 1],
                                                  mf->fe.val.data(), mf->fe.colloc.data(), s);
          else
            launch_laplace_general_dim<3, Number>(mf->p, atomic, mf->idx.p, (const Number *)op->cw.p, src, dst, mf->color_offsets[c], mf->color_offsets[c + 1],
                                                  mf->fe.val.data(), mf->fe.colloc.data(), s);
          time_end();
        }
    }
  else
    {
      // cell_loop (matrix_free_gpu.h:369-380): one launch per color
      for (uint32_t c = 0; c + 1 < mf->color_offsets.size(); ++c)
        {
          time_begin();
          launch_v0(mf->color_offsets[c], mf->color_offsets[c + 1], nullptr);
          time_end();
        }
    }
  // cells with hanging-node constraints (sorted to the end): column kernel with the interpolation fused into
  // gather and scatter (resolve_hanging_nodes_shmem, fee_gpu.cuh:333-335, 349-351)
  if (hanging && n_plain < mf->n_cells && part != 1)
    {
      time_begin();
      launch_v0(n_plain, mf->n_cells, mf->hn_mask.p);
      time_end();
    }
  // identity on the constrained rows (the cell kernel never writes them; the slab3 kernel does this copy itself)
  if (pdl_fill && op->ch->n() && !fused_ccopy) { constrained_copy<Number><<<nblk(op->ch->n()), 256, 0, s>>>(dst, src, op->ch->constrained.p, op->ch->n()); MFG_CUDA_LAST(); }
}

// Cell loop of the slab3 kernel alone, for the fused CG loop (solver.cu): dst must be ZERO on entry and written by the
// kernel in front of this one in the stream (the cell kernel is launched as its programmatic dependent and waits for it
// before its first write); the kernel also copies the constrained rows (dst[c] = src[c]) and stores per-warp partial
// sums of src . (A src) to dot_out (returns how many).  False if the operator's active kernel cannot do this.
bool laplace_cell_dot(mfg_laplace *op, void *dst, const void *src, double *dot_out, uint32_t *n_dot)
{
  const mfg_mf *mf = op->mf;
  if (laplace_active_variant(op) != 50 || mf->hn_mask.n != 0 || (slab3_flavour(op) & 1)) return false;
  laplace_prepare_slab2(op, mf->n_cells);
  if (op->slab2_groups == 0) return false;
  cudaStream_t s = op->ctx->stream;
  if (mf->dt == MFG_F64)
    launch_laplace_slab3<double>(mf->p, op->idxP.p, (const double *)op->cwP.p, (const double *)src, (double *)dst, op->slab2_groups, mf->fe.val.data(),
                                 mf->fe.colloc.data(), op->ctx->sm_count, s, op->mergeP.p, nullptr, true, 2, op->ctx->device, slab3_flavour(op),
                                 op->ch->constrained.p, (uint32_t)op->ch->n(), dot_out, n_dot);
  else
    launch_laplace_slab3<float>(mf->p, op->idxP.p, (const float *)op->cwP.p, (const float *)src, (float *)dst, op->slab2_groups, mf->fe.val.data(),
                                mf->fe.colloc.data(), op->ctx->sm_count, s, op->mergeP.p, nullptr, true, 2, op->ctx->device, slab3_flavour(op),
                                op->ch->constrained.p, (uint32_t)op->ch->n(), dot_out, n_dot);
  return true;
}

void laplace_kernel_time(mfg_laplace *op, double *total_ms, int *n_launches)
{
  MFG_CUDA(cudaStreamSynchronize(op->ctx->stream));
  double tot = 0;
  for (size_t i = 0; i + 1 < op->ev_used; i += 2)
    {
      float ms = 0;
      MFG_CUDA(cudaEventElapsedTime(&ms, op->ev[i], op->ev[i + 1]));
      tot += ms;
    }
  if (total_ms) *total_ms = tot;
  if (n_launches) *n_launches = (int)(op->ev_used / 2);
  op->ev_used = 0;
}

void laplace_vmult(mfg_laplace *op, void *dst, const void *src, bool add, int part, void *cuda_stream)
{
  MFG_REQUIRE(dst != src, "vmult: dst and src must not alias");
  MFG_REQUIRE(part >= -1 && part <= 2, "vmult: part must be -1, 0, 1 or 2");
  if (op->mf->dt == MFG_F64) vmult_impl<double>(op, (double *)dst, (const double *)src, add, part, (cudaStream_t)cuda_stream);
  else vmult_impl<float>(op, (float *)dst, (const float *)src, add, part, (cudaStream_t)cuda_stream);
}

// slab2 work list: 1 where a group holds an unconstrained entry of a flagged DoF
__global__ void mark_interface_groups(const uint32_t *__restrict__ idxP, size_t n_entries, int per_group, const uint8_t *__restrict__ flag,
                                      uint8_t *__restrict__ gflag)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_entries) return;
  const uint32_t id = idxP[t];
  if (!(id & CONSTRAINED_BIT) && flag[id]) gflag[t / per_group] = 1;
}

// staged kernel: bit 31 of the fourth descriptor word = the group contributes to an exchanged DoF
__global__ void set_interface_bits(uint32_t *__restrict__ gdesc, const uint8_t *__restrict__ gflag, uint32_t ng)
{
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < ng) gdesc[4 * (size_t)g + 3] = (gdesc[4 * (size_t)g + 3] & 0x7fffffffu) | (gflag && gflag[g] ? 0x80000000u : 0u);
}

// Multi-GPU: the DoFs whose partial sums are exchanged after the cell loop.  Splits the cell groups of the
// slab2 kernel into those that contribute to such a DoF (launched first) and the rest (launched while the exchange
// runs); returns the number of groups in the first set, 0 if the active kernel has no work list.
uint32_t laplace_set_interface_dofs(mfg_laplace *op, const uint32_t *dofs_host, size_t n)
{
  const mfg_mf *mf = op->mf;
  op->glist.release();
  op->n_iface_groups = 0;
  op->st_fb_iface = 0;
  if (op->st_built && op->slab2_groups)
    {
      set_interface_bits<<<nblk(op->slab2_groups), 256, 0, op->ctx->stream>>>(op->st_gdesc.p, nullptr, op->slab2_groups);
      MFG_CUDA_LAST();
    }
  const int av = laplace_active_variant(op);
  if (!grouped_variant(av) || n == 0) return 0;
  // cells with hanging nodes always run in part 2 and could add into an exchanged DoF while it is packed
  if (mf->hn_mask.n != 0) throw Error(MFG_ERR_UNSUPPORTED, "interface split is not available for meshes with hanging nodes");
  cudaStream_t   s = op->ctx->stream;
  const uint32_t n_plain = mf->hn_mask.n ? mf->n_plain : mf->n_cells;
  if (av == 40) laplace_prepare_stage(op, n_plain);
  else laplace_prepare_slab2(op, n_plain);
  const uint32_t ng = op->slab2_groups;
  if (ng == 0) return 0;
  for (size_t i = 0; i < n; ++i) MFG_REQUIRE(dofs_host[i] < mf->n_dofs, "interface DoF index out of range");
  DevBuf<uint32_t> list; list.upload(dofs_host, n, s);
  DevBuf<uint8_t> flag(mf->n_dofs), gflag(ng);
  MFG_CUDA(cudaMemsetAsync(flag.p, 0, mf->n_dofs, s));
  MFG_CUDA(cudaMemsetAsync(gflag.p, 0, ng, s));
  flags_from_list<<<nblk(n), 256, 0, s>>>(list.p, n, flag.p);
  MFG_CUDA_LAST();
  const int per_group = mf->n * mf->n * 32;
  mark_interface_groups<<<nblk(op->idxP.n), 256, 0, s>>>(op->idxP.p, op->idxP.n, per_group, flag.p, gflag.p);
  MFG_CUDA_LAST();
  std::vector<uint8_t> gf(ng);
  gflag.download(gf.data(), s);
  std::vector<uint32_t> order;
  order.reserve(ng);
  for (uint32_t g = 0; g < ng; ++g) if (gf[g]) order.push_back(g);
  op->n_iface_groups = (uint32_t)order.size();
  for (uint32_t g = 0; g < ng; ++g) if (!gf[g]) order.push_back(g);
  op->glist.upload(order.data(), ng, s);
  if (av == 40)
    {
      set_interface_bits<<<nblk(ng), 256, 0, s>>>(op->st_gdesc.p, gflag.p, ng);
      MFG_CUDA_LAST();
      MFG_CUDA(cudaStreamSynchronize(s));
    }
  if (av == 40 && op->st_fallback.n)
    {
      // left-over groups of the staged plan in the same two parts
      std::vector<uint32_t> fb(op->st_fallback.n), fo;
      op->st_fallback.download(fb.data(), s);
      for (uint32_t g : fb) if (gf[g]) fo.push_back(g);
      op->st_fb_iface = (uint32_t)fo.size();
      for (uint32_t g : fb) if (!gf[g]) fo.push_back(g);
      op->st_fallback.upload(fo.data(), fo.size(), s);
    }
  return op->n_iface_groups;
}

void laplace_compute_diagonal(mfg_laplace *op)
{
  const mfg_mf *mf = op->mf;
  cudaStream_t  s  = op->ctx->stream;
  if (!op->inv_diag)
    {
      op->inv_diag.reset(new mfg_vec);
      op->inv_diag->ctx = op->ctx; op->inv_diag->dt = mf->dt; op->inv_diag->n = op->inv_diag->cap = mf->n_dofs;
      MFG_CUDA(cudaMalloc(&op->inv_diag->p, (size_t)mf->n_dofs * op->inv_diag->esize()));
    }
  vec_fill(op->inv_diag.get(), 0.0);
  if (mf->general)
    {
      // diag_i = sum_q sum_de G_de(q) d_d phi_i(q) d_e phi_i(q) with the full metric tensor (kernels_general.cuh)
      DevBuf<double> tv, tg;
      tv.upload(mf->fe.val.data(), (size_t)mf->n * mf->n, s);
      tg.upload(mf->fe.grad.data(), (size_t)mf->n * mf->n, s);
      const int th = (int)std::min<uint32_t>(256, ((mf->npc + 31) / 32) * 32);
      if (mf->n_cells)
        {
          if (mf->dt == MFG_F64 && mf->dim == 2) diagonal_general<2, double><<<mf->n_cells, th, 0, s>>>(mf->idx.p, (const double *)op->cw.p, mf->n, mf->n_cells, tv.p, tg.p, (double *)op->inv_diag->p);
          else if (mf->dt == MFG_F64) diagonal_general<3, double><<<mf->n_cells, th, 0, s>>>(mf->idx.p, (const double *)op->cw.p, mf->n, mf->n_cells, tv.p, tg.p, (double *)op->inv_diag->p);
          else if (mf->dim == 2) diagonal_general<2, float><<<mf->n_cells, th, 0, s>>>(mf->idx.p, (const float *)op->cw.p, mf->n, mf->n_cells, tv.p, tg.p, (float *)op->inv_diag->p);
          else diagonal_general<3, float><<<mf->n_cells, th, 0, s>>>(mf->idx.p, (const float *)op->cw.p, mf->n, mf->n_cells, tv.p, tg.p, (float *)op->inv_diag->p);
          MFG_CUDA_LAST();
        }
      MFG_CUDA(cudaStreamSynchronize(s));  // tv / tg go out of scope
      ch_set(op->ch, op->inv_diag.get(), 1.0);
      vec_invert(op->inv_diag.get());
      op->diagonal_is_available = true;
      return;
    }
  DiagTables tb;
  for (int i = 0; i < mf->n * mf->n; ++i) { tb.val[i] = mf->fe.val[i]; tb.grad[i] = mf->fe.grad[i]; tb.hang[i] = mf->fe.hanging[i]; }
  const size_t dsm = 2 * (size_t)mf->npc * sizeof(double);
  const uint32_t *hmask = mf->hn_mask.n ? mf->hn_mask.p : nullptr;
  const int threads = (int)std::min<uint32_t>(256, ((mf->npc + 31) / 32) * 32);
  if (mf->n_cells == 0) { /* empty partition: only the constrained rows below */ }
  else if (mf->dt == MFG_F64)
    diagonal_kernel<double><<<mf->n_cells, threads, dsm, s>>>(mf->idx.p, (const double *)op->cw.p, mf->dim, mf->n, mf->npc, mf->n_cells, tb, (double *)op->inv_diag->p, hmask);
  else
    diagonal_kernel<float><<<mf->n_cells, threads, dsm, s>>>(mf->idx.p, (const float *)op->cw.p, mf->dim, mf->n, mf->npc, mf->n_cells, tb, (float *)op->inv_diag->p, hmask);
  MFG_CUDA_LAST();
  // constraint_handler.set_constrained_values(inv_diag, 1.0); inv_diag.invert()  (:416-418)
  ch_set(op->ch, op->inv_diag.get(), 1.0);
  vec_invert(op->inv_diag.get());
  op->diagonal_is_available = true;
}

}  // namespace mfg
