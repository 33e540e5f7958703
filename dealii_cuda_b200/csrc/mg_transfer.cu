// mg_transfer.cu -- MGTransferMatrixFreeGpu<dim,Number> for globally refined meshes
// (matrix_free_gpu/mg_transfer_matrix_free_gpu.{h,cu}): prolongate (.cu:592-622) and restrict_and_add (.cu:626-654)
// between a level mesh and its global refinement.  Like the reference's mg_kernel (.cu:540-556), one CTA handles one
// coarse cell: (p+1)^dim coarse values <-> (2p+1)^dim fine values through dim 1-D passes with the (2p+1)x(p+1)
// prolongation matrix P[f][i] = phi_i^coarse(x_f), inverse-valence weights (.cu:357-387) and atomic adds.  Differences:
// the fine index block of a coarse cell is generated on the device from the fine mesh's closed-form numbering
// (the reference builds it on the host through deal.II's internal::MGTransfer::setup_transfer, .cu:173-257); coarse
// Dirichlet DoFs are skipped on read/write instead of copying src and zeroing a temporary (.cu:602-605, 650).
#include "operators.cuh"

struct mfg_mgt
{
  mfg_ctx *ctx = nullptr;
  mfg_dtype dt = MFG_F64;
  int dim = 0, p = 0;
  uint32_t n_coarse_cells = 0, n_coarse_dofs = 0, n_fine_dofs = 0, nc[3] = {1, 1, 1};
  mfg::DevBuf<uint32_t> coarse_idx;  // [n_coarse_cells][(p+1)^dim], bit 31 = coarse Dirichlet DoF
  mfg::DevBuf<uint32_t> fine_idx;    // [n_coarse_cells][(2p+1)^dim] lexicographic
  mfg::DevBuf<uint32_t> cell_xyz;    // [n_coarse_cells][3] (valence weights)
  mfg::DevBuf<double>   wtab;        // adaptive hierarchies: [n_coarse_cells][3^dim] weights per block region (empty: closed-form valence)
  double P[17 * 9];                  // P[f*(p+1)+i]
};

namespace mfg {
void mesh_lattice_to_dof_device(const mfg_mesh *m, size_t npts, const uint32_t *xyz_dev, uint32_t *out_dev);

namespace {

struct PMat { double P[17 * 9]; };

__global__ void fine_lattice_points(int dim, int p, uint32_t n_cells, const uint32_t *__restrict__ cxyz, uint32_t *__restrict__ pts)
{
  const int nf = 2 * p + 1;
  const uint32_t per = dim == 3 ? nf * nf * nf : nf * nf;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)n_cells * per) return;
  const uint32_t c = (uint32_t)(t / per), l = (uint32_t)(t % per);
  const uint32_t a[3] = {l % nf, (l / nf) % nf, dim == 3 ? l / (nf * nf) : 0};
  for (int d = 0; d < 3; ++d) pts[3 * t + d] = d < dim ? 2 * p * cxyz[3 * (size_t)c + d] + a[d] : 0;
}

// 1-D pass along direction d on a tensor with extents ext[]: PROLONG expands (p+1 -> 2p+1), else contracts with P^T
template <bool PROLONG>
__device__ void pass(const double *in, double *out, int dim, int d, const int *ext_in, const int *ext_out, int p, const double *P)
{
  const int nfine = 2 * p + 1, ncoarse = p + 1;
  const int tot = ext_out[0] * ext_out[1] * ext_out[2];
  int sin[3] = {1, ext_in[0], ext_in[0] * ext_in[1]};
  for (int o = threadIdx.x; o < tot; o += blockDim.x)
    {
      int c[3] = {o % ext_out[0], (o / ext_out[0]) % ext_out[1], o / (ext_out[0] * ext_out[1])};
      const int k = c[d];
      c[d] = 0;
      const int base = c[0] * sin[0] + c[1] * sin[1] + c[2] * sin[2];
      double acc = 0;
      if (PROLONG) for (int i = 0; i < ncoarse; ++i) acc += P[k * ncoarse + i] * in[base + i * sin[d]];
      else for (int f = 0; f < nfine; ++f) acc += P[f * ncoarse + k] * in[base + f * sin[d]];
      out[o] = acc;
    }
  (void)dim;
}

__device__ inline double valence_weight(int dim, int p, const uint32_t *cx, const uint32_t *nc, const int *a)
{
  double w = 1.0;
  for (int d = 0; d < dim; ++d)
    if ((a[d] == 0 && cx[d] > 0) || (a[d] == 2 * p && cx[d] + 1 < nc[d])) w *= 0.5;
  return w;
}

struct MgArgs { int dim, p; uint32_t n_cells; uint32_t nc[3]; const uint32_t *coarse_idx, *fine_idx, *cxyz; const double *wtab; };

// weight of lattice point a of a block: closed-form valence on globally refined meshes, else the block's table of 3^dim
// regions (low face / interior / high face per direction; weights_on_refined, mg_transfer_matrix_free_gpu.cu:357-387)
__device__ inline double block_weight(const MgArgs &A, uint32_t cell, const uint32_t *cx, const int *a)
{
  if (A.wtab == nullptr) return valence_weight(A.dim, A.p, cx, A.nc, a);
  int r = 0, s = 1;
  for (int d = 0; d < A.dim; ++d) { r += (a[d] == 0 ? 0 : a[d] == 2 * A.p ? 2 : 1) * s; s *= 3; }
  const int n3 = A.dim == 3 ? 27 : 9;
  return A.wtab[(size_t)cell * n3 + r];
}

template <typename Number, bool PROLONG>
__global__ void mg_kernel(MgArgs A, PMat pm, Number *__restrict__ dst, const Number *__restrict__ src)
{
  extern __shared__ double sm[];
  const int p = A.p, dim = A.dim, nfine = 2 * p + 1, ncoarse = p + 1;
  const int nF = dim == 3 ? nfine * nfine * nfine : nfine * nfine, nC = dim == 3 ? ncoarse * ncoarse * ncoarse : ncoarse * ncoarse;
  double *b0 = sm, *b1 = sm + nF;
  const uint32_t cell = blockIdx.x;
  const uint32_t *cidx = A.coarse_idx + (size_t)cell * nC, *fidx = A.fine_idx + (size_t)cell * nF;
  const uint32_t *cx = A.cxyz + 3 * (size_t)cell;
  int ext[3] = {PROLONG ? ncoarse : nfine, PROLONG ? ncoarse : nfine, dim == 3 ? (PROLONG ? ncoarse : nfine) : 1};
  if (PROLONG)
    for (int i = threadIdx.x; i < nC; i += blockDim.x)
      {
        const uint32_t g = cidx[i];
        b0[i] = (g & 0x80000000u) ? 0.0 : (double)src[g];   // set_mg_constrained_dofs(src_with_bc, to_level-1, 0)
      }
  else
    for (int f = threadIdx.x; f < nF; f += blockDim.x)
      {
        const int a[3] = {f % nfine, (f / nfine) % nfine, dim == 3 ? f / (nfine * nfine) : 0};
        b0[f] = (double)src[fidx[f]] * block_weight(A, cell, cx, a);   // weigh_values
      }
  __syncthreads();
  double *in = b0, *out = b1;
  for (int d = 0; d < dim; ++d)
    {
      int ext_out[3] = {ext[0], ext[1], ext[2]};
      ext_out[d] = PROLONG ? nfine : ncoarse;
      pass<PROLONG>(in, out, dim, d, ext, ext_out, p, pm.P);
      ext[d] = ext_out[d];
      __syncthreads();
      double *t = in; in = out; out = t;
    }
  if (PROLONG)
    for (int f = threadIdx.x; f < nF; f += blockDim.x)
      {
        const int a[3] = {f % nfine, (f / nfine) % nfine, dim == 3 ? f / (nfine * nfine) : 0};
        atomicAdd(dst + fidx[f], (Number)(in[f] * block_weight(A, cell, cx, a)));
      }
  else
    for (int i = threadIdx.x; i < nC; i += blockDim.x)
      {
        const uint32_t g = cidx[i];
        if (!(g & 0x80000000u)) atomicAdd(dst + g, (Number)in[i]);   // set_mg_constrained_dofs(increment, from_level-1, 0)
      }
}

__global__ void mark_coarse(const uint32_t *l2g, const uint8_t *cflag, size_t n, uint32_t *out)
{
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { const uint32_t g = l2g[t]; out[t] = g | (cflag[g] ? 0x80000000u : 0u); }
}

}  // namespace
}  // namespace mfg

using namespace mfg;

extern "C" {

int mfg_mgt_build(mfg_ctx *ctx, const mfg_mesh *coarse, const mfg_mesh *fine, mfg_dtype dt, mfg_mgt **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && coarse && fine && out, "null argument");
    MFG_REQUIRE(coarse->dim == fine->dim && coarse->p == fine->p, "meshes differ in dimension or degree");
    for (int d = 0; d < coarse->dim; ++d) MFG_REQUIRE(fine->lg[d] == coarse->lg[d] + 1, "fine mesh must be the global refinement of the coarse mesh");
    std::unique_ptr<mfg_mgt> t(new mfg_mgt);
    t->ctx = ctx; t->dt = dt; t->dim = coarse->dim; t->p = coarse->p;
    t->n_coarse_cells = coarse->n_cells; t->n_coarse_dofs = coarse->n_dofs; t->n_fine_dofs = fine->n_dofs;
    for (int d = 0; d < 3; ++d) t->nc[d] = coarse->nc[d];
    const int p = t->p, n = p + 1, nf = 2 * p + 1;
    // P[f][i] = phi_i(x_f): fine support points of the two children in coarse reference coordinates
    const FEData1D &fe = coarse->fe;
    for (int f = 0; f < nf; ++f)
      {
        const long double x = f <= p ? (long double)fe.nodes[f] / 2 : 0.5L + (long double)fe.nodes[f - p] / 2;
        for (int i = 0; i < n; ++i)
          {
            long double v = 1;
            for (int m = 0; m < n; ++m) if (m != i) v *= (x - fe.nodes[m]) / ((long double)fe.nodes[i] - fe.nodes[m]);
            t->P[f * n + i] = (double)v;
          }
      }
    cudaStream_t s = ctx->stream;
    // 3D degree 7, 8: the two staged tensors of (2p+1)^dim doubles exceed the 48 KB a kernel gets without opting in
    const size_t smem_needed = 2 * (size_t)ipow(nf, t->dim) * sizeof(double);
    if (smem_needed > 48 * 1024)
      {
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
      }
    const size_t nC = (size_t)coarse->n_cells * coarse->npc;
    t->coarse_idx.alloc(nC);
    mark_coarse<<<(unsigned)((nC + 255) / 256), 256, 0, s>>>(coarse->l2g.p, coarse->cflag.p, nC, t->coarse_idx.p);
    MFG_CUDA_LAST();
    std::vector<uint32_t> cxyz((size_t)coarse->n_cells * 3);
    mesh_cell_coords(coarse, cxyz.data());
    t->cell_xyz.upload(cxyz.data(), cxyz.size(), s);
    const size_t per = ipow(nf, t->dim), npts = (size_t)coarse->n_cells * per;
    DevBuf<uint32_t> pts(3 * npts);
    fine_lattice_points<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(t->dim, p, coarse->n_cells, t->cell_xyz.p, pts.p);
    MFG_CUDA_LAST();
    t->fine_idx.alloc(npts);
    mesh_lattice_to_dof_device(fine, npts, pts.p, t->fine_idx.p);
    MFG_CUDA(cudaStreamSynchronize(s));
    *out = t.release();
  });
}
// transfer between two levels of an ADAPTIVE hierarchy from explicit blocks (mfg_amesh_mg_level_get): one block per refined
// cell of the coarse level
int mfg_mgt_build_from_blocks(mfg_ctx *ctx, mfg_dtype dt, int dim, int degree, uint32_t n_blocks, const uint32_t *coarse_idx_host, const uint32_t *fine_idx_host,
                              const double *weights_host, uint32_t n_coarse_dofs, uint32_t n_fine_dofs, mfg_mgt **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out, "null argument");
    MFG_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
    MFG_REQUIRE(degree >= 1 && degree <= 8, "degree must be in 1..8");
    MFG_REQUIRE(n_blocks == 0 || (coarse_idx_host && fine_idx_host && weights_host), "null block arrays");
    std::unique_ptr<mfg_mgt> t(new mfg_mgt);
    t->ctx = ctx; t->dt = dt; t->dim = dim; t->p = degree;
    t->n_coarse_cells = n_blocks; t->n_coarse_dofs = n_coarse_dofs; t->n_fine_dofs = n_fine_dofs;
    const int p = degree, n = p + 1, nf = 2 * p + 1;
    const size_t nC = ipow(n, dim), nF = ipow(nf, dim), n3 = ipow(3, dim);
    for (size_t i = 0; i < (size_t)n_blocks * nC; ++i) MFG_REQUIRE((coarse_idx_host[i] & 0x7fffffffu) < n_coarse_dofs, "coarse index out of range");
    for (size_t i = 0; i < (size_t)n_blocks * nF; ++i) MFG_REQUIRE(fine_idx_host[i] < n_fine_dofs, "fine index out of range");
    const FEData1D fe = make_fe_data(degree);
    for (int f = 0; f < nf; ++f)
      {
        const long double x = f <= p ? (long double)fe.nodes[f] / 2 : 0.5L + (long double)fe.nodes[f - p] / 2;
        for (int i = 0; i < n; ++i)
          {
            long double v = 1;
            for (int m = 0; m < n; ++m) if (m != i) v *= (x - fe.nodes[m]) / ((long double)fe.nodes[i] - fe.nodes[m]);
            t->P[f * n + i] = (double)v;
          }
      }
    cudaStream_t s = ctx->stream;
    const size_t smem_needed = 2 * nF * sizeof(double);
    if (smem_needed > 48 * 1024)
      {
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
        MFG_CUDA(cudaFuncSetAttribute(mg_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_needed));
      }
    if (n_blocks)
      {
        t->coarse_idx.upload(coarse_idx_host, (size_t)n_blocks * nC, s);
        t->fine_idx.upload(fine_idx_host, (size_t)n_blocks * nF, s);
        t->wtab.upload(weights_host, (size_t)n_blocks * n3, s);
        std::vector<uint32_t> zero((size_t)n_blocks * 3, 0u);
        t->cell_xyz.upload(zero.data(), zero.size(), s);
      }
    *out = t.release();
  });
}
int mfg_mgt_destroy(mfg_mgt *t) { return guarded([&] { delete t; }); }

static void mgt_run(mfg_mgt *t, mfg_vec *dst, const mfg_vec *src, bool prolong)
{
  MFG_REQUIRE(t && dst && src, "null argument");
  MFG_REQUIRE(dst->dt == t->dt && src->dt == t->dt, "vector dtype differs from transfer dtype");
  MFG_REQUIRE(dst->n == (prolong ? t->n_fine_dofs : t->n_coarse_dofs) && src->n == (prolong ? t->n_coarse_dofs : t->n_fine_dofs), "vector sizes do not match the levels");
  MgArgs A; A.dim = t->dim; A.p = t->p; A.n_cells = t->n_coarse_cells;
  for (int d = 0; d < 3; ++d) A.nc[d] = t->nc[d];
  A.coarse_idx = t->coarse_idx.p; A.fine_idx = t->fine_idx.p; A.cxyz = t->cell_xyz.p; A.wtab = t->wtab.n ? t->wtab.p : nullptr;
  if (t->n_coarse_cells == 0) { if (prolong) vec_fill(dst, 0.0); return; }
  PMat pm; std::memcpy(pm.P, t->P, sizeof(pm.P));
  const size_t nF = ipow(2 * t->p + 1, t->dim), smem = 2 * nF * sizeof(double);
  cudaStream_t s = t->ctx->stream;
  if (prolong) vec_fill(dst, 0.0);  // dst = 0  (.cu:600)
  if (t->dt == MFG_F64)
    {
      if (prolong) mg_kernel<double, true><<<t->n_coarse_cells, 128, smem, s>>>(A, pm, (double *)dst->p, (const double *)src->p);
      else mg_kernel<double, false><<<t->n_coarse_cells, 128, smem, s>>>(A, pm, (double *)dst->p, (const double *)src->p);
    }
  else
    {
      if (prolong) mg_kernel<float, true><<<t->n_coarse_cells, 128, smem, s>>>(A, pm, (float *)dst->p, (const float *)src->p);
      else mg_kernel<float, false><<<t->n_coarse_cells, 128, smem, s>>>(A, pm, (float *)dst->p, (const float *)src->p);
    }
  MFG_CUDA_LAST();
}
int mfg_mgt_prolongate(mfg_mgt *t, mfg_vec *dst_fine, const mfg_vec *src_coarse) { return guarded([&] { mgt_run(t, dst_fine, src_coarse, true); }); }
int mfg_mgt_restrict_and_add(mfg_mgt *t, mfg_vec *dst_coarse, const mfg_vec *src_fine) { return guarded([&] { mgt_run(t, dst_coarse, src_fine, false); }); }

}  // extern "C"
