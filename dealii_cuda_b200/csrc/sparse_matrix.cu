// sparse_matrix.cu -- the assembled-matrix competitor of the matrix-free operator: CUDAWrappers::SparseMatrix<Number>
// (matrix_free_gpu/cuda_sparse_matrix.{h,cu}) as the reference's bmop_spm.cu / poisson_spm.cu / test_spm.cu use it.
//
// The reference assembles a dealii::SparseMatrix on the host with FEValues (bmop_spm.cu:150-201: cell matrices
// sum_q a(x_q) grad phi_i . grad phi_j JxW, ConstraintMatrix::distribute_local_to_global), copies it to the device as CSR and
// multiplies with cusparse<t>csrmv (cuda_sparse_matrix.cu:414-429, an API removed from CUDA since).  Here:
//   * assembly is host code of the library from the same arrays the matrix-free operator takes (loc2glob, J^-1, coefficient at
//     the quadrature points, constraint list): cell matrix A_ij = sum_q cw_q sum_d G_d[i][q] G_d[j][q] with the tensor-product
//     gradient tables; constrained rows and columns are eliminated, their diagonal is 1 (the operator's constrained rows are
//     the identity, laplace_operator_gpu.h:286-303);
//   * vmult is a hand-written CSR kernel, one warp per row (rows of a Q4 3D matrix hold 125..729 entries), so that the
//     comparison does not depend on a library being present.
// Uniform geometry, no hanging nodes (what bmop_spm.cu measures).
#include <algorithm>
#include <memory>
#include "operators.cuh"

using namespace mfg;

struct mfg_csr  // host CSR
{
  uint32_t              n = 0;
  std::vector<uint32_t> row_ptr, col;
  std::vector<double>   val;
};

struct mfg_spm
{
  mfg_ctx *ctx = nullptr;
  mfg_dtype dt = MFG_F64;
  uint32_t  n = 0;
  size_t    nnz = 0;
  DevBuf<uint32_t> row_ptr, col;
  DevBuf<uint8_t>  val;
};

namespace {

template <typename T>
__global__ void csr_vmult_warp_per_row(uint32_t n, const uint32_t *__restrict__ row_ptr, const uint32_t *__restrict__ col, const T *__restrict__ val,
                                       const T *__restrict__ x, T *__restrict__ y)
{
  const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const uint32_t b = row_ptr[row], e = row_ptr[row + 1];
  T acc = 0;
  for (uint32_t k = b + lane; k < e; k += 32) acc += val[k] * x[col[k]];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if (lane == 0) y[row] = acc;
}

void assemble(int dim, int degree, uint32_t n_cells, uint32_t n_dofs, const uint32_t *l2g, const double *inv_jac, const double *coef,
              const uint32_t *constrained, size_t n_constrained, mfg_csr &A)
{
  MFG_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  MFG_REQUIRE(degree >= 1 && degree <= 8, "degree must be in 1..8");
  const FEData1D fe = make_fe_data(degree);
  const int      n = degree + 1;
  const uint32_t npc = ipow(n, dim);
  std::vector<uint8_t> is_c(n_dofs, 0);
  for (size_t i = 0; i < n_constrained; ++i) { MFG_REQUIRE(constrained[i] < n_dofs, "constrained index out of range"); is_c[constrained[i]] = 1; }
  // DoF -> cells
  std::vector<uint32_t> dstart((size_t)n_dofs + 1, 0);
  for (size_t t = 0; t < (size_t)n_cells * npc; ++t) { MFG_REQUIRE(l2g[t] < n_dofs, "loc2glob entry out of range"); ++dstart[l2g[t] + 1]; }
  for (uint32_t i = 0; i < n_dofs; ++i) dstart[i + 1] += dstart[i];
  std::vector<uint32_t> dcell(dstart.back()), fill(dstart.begin(), dstart.end() - 1);
  for (uint32_t c = 0; c < n_cells; ++c)
    for (uint32_t i = 0; i < npc; ++i) dcell[fill[l2g[(size_t)c * npc + i]]++] = c;
  // sparsity: row i couples with every unconstrained DoF of the cells that hold i; constrained rows: the diagonal only
  A.n = n_dofs;
  A.row_ptr.assign((size_t)n_dofs + 1, 0);
  std::vector<uint32_t> tmp;
  std::vector<std::vector<uint32_t>> rows(n_dofs);
  for (uint32_t r = 0; r < n_dofs; ++r)
    {
      if (is_c[r]) { rows[r] = {r}; continue; }
      tmp.clear();
      for (uint32_t k = dstart[r]; k < dstart[r + 1]; ++k)
        for (uint32_t j = 0; j < npc; ++j)
          {
            const uint32_t g = l2g[(size_t)dcell[k] * npc + j];
            if (!is_c[g]) tmp.push_back(g);
          }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      rows[r] = tmp;
    }
  uint64_t total = 0;
  for (uint32_t r = 0; r < n_dofs; ++r) total += rows[r].size();
  MFG_REQUIRE(total < (1ull << 32), "more than 2^32 matrix entries");
  for (uint32_t r = 0; r < n_dofs; ++r) A.row_ptr[r + 1] = A.row_ptr[r] + (uint32_t)rows[r].size();
  A.col.resize(A.row_ptr[n_dofs]);
  A.val.assign(A.row_ptr[n_dofs], 0.0);
  for (uint32_t r = 0; r < n_dofs; ++r)
    {
      std::copy(rows[r].begin(), rows[r].end(), A.col.begin() + A.row_ptr[r]);
      std::vector<uint32_t>().swap(rows[r]);
      if (is_c[r]) A.val[A.row_ptr[r]] = 1.0;
    }
  // gradient tables G_d[i][q] = prod_e (e == d ? phi'_{i_e}(x_{q_e}) : phi_{i_e}(x_{q_e}))   (reference-cell derivatives)
  std::vector<double> G((size_t)dim * npc * npc);
  for (int d = 0; d < dim; ++d)
    for (uint32_t i = 0; i < npc; ++i)
      for (uint32_t q = 0; q < npc; ++q)
        {
          double   v = 1.0;
          uint32_t ii = i, qq = q;
          for (int e = 0; e < dim; ++e)
            {
              const int ie = ii % n, qe = qq % n;
              ii /= n; qq /= n;
              v *= (e == d ? fe.grad : fe.val)[ie * n + qe];
            }
          G[((size_t)d * npc + i) * npc + q] = v;
        }
  std::vector<double> cw(npc), Gw((size_t)dim * npc * npc), Ac((size_t)npc * npc);
  for (uint32_t c = 0; c < n_cells; ++c)
    {
      // merged weight a(x_q) J^-2 JxW_q, JxW_q = J^dim w_q  (fee_gpu.cuh:219-284 on a uniform cell)
      for (uint32_t q = 0; q < npc; ++q)
        {
          double   w = 1.0;
          uint32_t qq = q;
          for (int e = 0; e < dim; ++e) { w *= fe.qwts[qq % n] / inv_jac[c]; qq /= n; }
          cw[q] = coef[(size_t)c * npc + q] * inv_jac[c] * inv_jac[c] * w;
        }
      for (int d = 0; d < dim; ++d)
        for (uint32_t i = 0; i < npc; ++i)
          for (uint32_t q = 0; q < npc; ++q) Gw[((size_t)d * npc + i) * npc + q] = G[((size_t)d * npc + i) * npc + q] * cw[q];
      std::fill(Ac.begin(), Ac.end(), 0.0);
      for (int d = 0; d < dim; ++d)
        for (uint32_t i = 0; i < npc; ++i)
          {
            const double *gi = &Gw[((size_t)d * npc + i) * npc];
            for (uint32_t j = 0; j < npc; ++j)
              {
                const double *gj = &G[((size_t)d * npc + j) * npc];
                double        s = 0;
                for (uint32_t q = 0; q < npc; ++q) s += gi[q] * gj[q];
                Ac[(size_t)i * npc + j] += s;
              }
          }
      const uint32_t *row = l2g + (size_t)c * npc;
      for (uint32_t i = 0; i < npc; ++i)
        {
          if (is_c[row[i]]) continue;
          const uint32_t *cb = A.col.data() + A.row_ptr[row[i]], *ce = A.col.data() + A.row_ptr[row[i] + 1];
          for (uint32_t j = 0; j < npc; ++j)
            {
              if (is_c[row[j]]) continue;
              const uint32_t *pos = std::lower_bound(cb, ce, row[j]);
              A.val[pos - A.col.data()] += Ac[(size_t)i * npc + j];
            }
        }
    }
}

}  // namespace

extern "C" {

int mfg_csr_assemble_laplace(int dim, int degree, uint32_t n_cells, uint32_t n_dofs, const uint32_t *loc2glob_host, const double *inv_jac_host,
                             const double *coefficient_host, const uint32_t *constrained_host, size_t n_constrained, mfg_csr **out)
{
  return guarded([&] {
    MFG_REQUIRE(out && loc2glob_host && inv_jac_host && coefficient_host && (constrained_host || !n_constrained), "null argument");
    std::unique_ptr<mfg_csr> A(new mfg_csr);
    assemble(dim, degree, n_cells, n_dofs, loc2glob_host, inv_jac_host, coefficient_host, constrained_host, n_constrained, *A);
    *out = A.release();
  });
}
int mfg_csr_destroy(mfg_csr *A) { return guarded([&] { delete A; }); }
int mfg_csr_sizes(const mfg_csr *A, uint32_t *n_rows, size_t *nnz)
{
  return guarded([&] { MFG_REQUIRE(A, "null argument"); if (n_rows) *n_rows = A->n; if (nnz) *nnz = A->val.size(); });
}
int mfg_csr_get(const mfg_csr *A, uint32_t *row_ptr, uint32_t *col, double *val)
{
  return guarded([&] {
    MFG_REQUIRE(A, "null argument");
    if (row_ptr) std::copy(A->row_ptr.begin(), A->row_ptr.end(), row_ptr);
    if (col) std::copy(A->col.begin(), A->col.end(), col);
    if (val) std::copy(A->val.begin(), A->val.end(), val);
  });
}

// CUDAWrappers::SparseMatrix::reinit(host matrix) (cuda_sparse_matrix.cu:60-120): CSR arrays to the device, values in dtype
int mfg_spm_create(mfg_ctx *ctx, mfg_dtype dt, const mfg_csr *A, mfg_spm **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && A && out, "null argument");
    std::unique_ptr<mfg_spm> S(new mfg_spm);
    S->ctx = ctx; S->dt = dt; S->n = A->n; S->nnz = A->val.size();
    cudaStream_t s = ctx->stream;
    S->row_ptr.upload(A->row_ptr.data(), A->row_ptr.size(), s);
    S->col.upload(A->col.data(), A->col.size(), s);
    if (dt == MFG_F64) S->val.upload((const uint8_t *)A->val.data(), A->val.size() * 8, s);
    else
      {
        std::vector<float> v32(A->val.begin(), A->val.end());
        S->val.upload((const uint8_t *)v32.data(), v32.size() * 4, s);
      }
    *out = S.release();
  });
}
// the same for the operator of a uniform mesh object: arrays from the mesh, coefficient 1/(0.05 + 2|x|^2) at the Gauss points
int mfg_spm_create_from_mesh(mfg_ctx *ctx, const mfg_mesh *mesh, mfg_dtype dt, mfg_spm **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && mesh && out, "null argument");
    const uint32_t nc = mesh->n_cells, npc = mesh->npc;
    std::vector<uint32_t> l2g((size_t)nc * npc), con(mesh->n_constrained), cxyz((size_t)nc * 3);
    mesh->l2g.download(l2g.data(), ctx->stream);
    mesh->constrained.download(con.data(), ctx->stream);
    mesh_cell_coords(mesh, cxyz.data());
    std::vector<double> inv_jac(nc, 1.0 / mesh->h), coef((size_t)nc * npc);
    for (uint32_t c = 0; c < nc; ++c)
      for (uint32_t q = 0; q < npc; ++q)
        {
          uint32_t t = q;
          double   r2 = 0;
          for (int d = 0; d < mesh->dim; ++d)
            {
              const double x = mesh->origin[d] + mesh->h * ((double)cxyz[3 * (size_t)c + d] + mesh->fe.qpts[t % mesh->n]);
              t /= mesh->n;
              r2 += x * x;
            }
          coef[(size_t)c * npc + q] = 1.0 / (0.05 + 2.0 * r2);
        }
    mfg_csr A;
    assemble(mesh->dim, mesh->p, nc, mesh->n_dofs, l2g.data(), inv_jac.data(), coef.data(), con.data(), con.size(), A);
    mfg_spm *S = nullptr;
    const int rc = mfg_spm_create(ctx, dt, &A, &S);
    if (rc != MFG_OK) throw Error(rc, std::string(mfg_last_error()));
    *out = S;
  });
}
int mfg_spm_destroy(mfg_spm *S) { return guarded([&] { delete S; }); }
uint32_t mfg_spm_m(const mfg_spm *S) { return S ? S->n : 0; }
size_t mfg_spm_n_nonzero_elements(const mfg_spm *S) { return S ? S->nnz : 0; }
size_t mfg_spm_memory_consumption(const mfg_spm *S) { return S ? S->row_ptr.bytes() + S->col.bytes() + S->val.bytes() : 0; }
// SparseMatrix::vmult (cuda_sparse_matrix.cu:414-429): dst = A src
int mfg_spm_vmult(mfg_spm *S, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    MFG_REQUIRE(S && dst && src && dst != src, "null or aliased argument");
    MFG_REQUIRE(dst->n == S->n && src->n == S->n && dst->dt == S->dt && src->dt == S->dt, "vector does not fit the matrix");
    if (S->n == 0) return;
    const unsigned threads = 256, rows_per_block = threads / 32, blocks = (S->n + rows_per_block - 1) / rows_per_block;
    cudaStream_t   s = S->ctx->stream;
    if (S->dt == MFG_F64)
      csr_vmult_warp_per_row<double><<<blocks, threads, 0, s>>>(S->n, S->row_ptr.p, S->col.p, (const double *)S->val.p, (const double *)src->p, (double *)dst->p);
    else
      csr_vmult_warp_per_row<float><<<blocks, threads, 0, s>>>(S->n, S->row_ptr.p, S->col.p, (const float *)S->val.p, (const float *)src->p, (float *)dst->p);
    MFG_CUDA_LAST();
  });
}

}  // extern "C"
