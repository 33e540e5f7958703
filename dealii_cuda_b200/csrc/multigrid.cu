// multigrid.cu -- PreconditionChebyshev, the multigrid V-cycle and the V-cycle-preconditioned CG of the reference's
// poisson_mg.cu / bmop_mg.cu (poisson_mg.cu:430-552, bmop_mg.cu:300-340) behind the C ABI.
//
// In the reference this is deal.II template code instantiated on GpuVector (PreconditionChebyshev, Multigrid,
// PreconditionMG, MGSmootherPrecondition, MGCoarseIterative = SolverCG to a 1e-10 reduction, SolverCG outside), every
// vector operation a separate BLAS-1 kernel.  Here it is C++ host code over the same library objects (level
// LaplaceOperatorGpu, MGTransferMatrixFreeGpu) with ONE fused vector kernel per Chebyshev matrix-vector product:
//   t = Dinv (b - A x);  d = f1 d + f2 t;  x += d              (PreconditionChebyshev::vector_updates)
// deal.II conventions restated (deal.II 8.5 precondition.h, from memory -- deal.II is not available):
//   * eigenvalue estimate: eig_cg_n_iterations steps of SolverCG on Dinv A from x = 0 with the right-hand side
//     1/sqrt(n) (entry 0 set to 0), stopping early at |g| <= 1e-2 (eig_cg_residual); lambda_max / lambda_min = largest /
//     smallest eigenvalue of the Lanczos tridiagonal matrix built from the CG coefficients;
//   * beta = 1.2 lambda_max, alpha = lambda_max / smoothing_range (range > 1), delta = (beta - alpha) / 2,
//     theta = (beta + alpha) / 2, rho_0 = delta / theta, sigma = theta / delta, rho_{k+1} = 1 / (2 sigma - rho_k);
//   * vmult (zero initial guess): x = d = Dinv b / theta, then `degree` products; step (given x): the same with the
//     first update d = Dinv (b - A x) / theta;
//   * V-cycle (Multigrid::level_v_step): pre-smooth from zero, t = b - A x, restrict_and_add into a zeroed defect,
//     recurse, prolongate, x += t, post-smooth (step); coarsest level: unpreconditioned CG to 1e-10 |b|.
// Globally refined meshes: there are no refinement edges, so the edge matrices (vmult_interface_down / up,
// laplace_operator_gpu.h:306-352) are the zero operator and are not called.
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>
#include "operators.cuh"

using namespace mfg;

extern "C" int mfg_solver_cg(mfg_laplace *op, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int use_jacobi, int *iters,
                             double *last_residual, double *residual_history);

namespace {

#define MG_CHECK(call)                                                                    \
  do {                                                                                    \
    const int rc__ = (call);                                                              \
    if (rc__ != MFG_OK) throw Error(rc__, std::string(mfg_last_error()));                 \
  } while (0)

struct Vec  // owning GpuVector
{
  mfg_vec *v = nullptr;
  Vec() = default;
  Vec(mfg_ctx *ctx, mfg_dtype dt, size_t n) { MG_CHECK(mfg_vec_create(ctx, dt, n, &v)); }
  Vec(const Vec &) = delete;
  Vec &operator=(const Vec &) = delete;
  Vec(Vec &&o) noexcept : v(o.v) { o.v = nullptr; }
  Vec &operator=(Vec &&o) noexcept { if (this != &o) { if (v) mfg_vec_destroy(v); v = o.v; o.v = nullptr; } return *this; }
  ~Vec() { if (v) mfg_vec_destroy(v); }
};

// x, d: updated; ax = A x on entry (first: ignored when zero_start); b: right-hand side
// zero_start: d = f2 Dinv b, x = d;   else: t = Dinv (b - ax), d = f1 d + f2 t, x += d
template <typename T>
__global__ void cheb_update(T *__restrict__ x, T *__restrict__ d, const T *__restrict__ ax, const T *__restrict__ b, const T *__restrict__ dinv, size_t n,
                            T f1, T f2, bool zero_start, bool first)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      const T r = zero_start ? b[i] : b[i] - ax[i];
      const T di = (first ? T(0) : f1 * d[i]) + f2 * dinv[i] * r;
      d[i] = di;
      x[i] = zero_start ? di : x[i] + di;
    }
}

// eigenvalues of the symmetric tridiagonal matrix (diag a, off-diagonal b) by bisection on the Sturm count
int sturm_count(const std::vector<double> &a, const std::vector<double> &b, double x)
{
  int    cnt = 0;
  double q = a[0] - x;
  if (q < 0) ++cnt;
  for (size_t i = 1; i < a.size(); ++i)
    {
      q = a[i] - x - b[i - 1] * b[i - 1] / (q != 0 ? q : 1e-300);
      if (q < 0) ++cnt;
    }
  return cnt;  // number of eigenvalues < x
}
double tridiag_eigenvalue(const std::vector<double> &a, const std::vector<double> &b, int k)  // k-th smallest (0-based)
{
  double lo = 1e300, hi = -1e300;
  for (size_t i = 0; i < a.size(); ++i)
    {
      const double r = (i ? std::fabs(b[i - 1]) : 0) + (i + 1 < a.size() ? std::fabs(b[i]) : 0);
      lo = std::min(lo, a[i] - r); hi = std::max(hi, a[i] + r);
    }
  for (int it = 0; it < 200; ++it)
    {
      const double mid = 0.5 * (lo + hi);
      if (sturm_count(a, b, mid) > k) hi = mid; else lo = mid;
    }
  return 0.5 * (lo + hi);
}

// SolverCG control flow (poisson.cu:233-260) with a preconditioner given as a callable h = M^-1 g
template <typename Precond>
void solve_cg_preconditioned(mfg_ctx *ctx, mfg_dtype dt, mfg_laplace *op, Precond &&precond, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter,
                             int *iters, double *last_residual, double *history)
{
  const size_t n = op->mf->n_dofs;
  Vec g(ctx, dt, n), h(ctx, dt, n), d(ctx, dt, n);
  if (vec_all_zero(x)) vec_equ(g.v, -1.0, b);
  else { laplace_vmult(op, g.v->p, x->p, false); vec_sadd(g.v, 1.0, -1.0, b); }
  double res = std::sqrt(vec_dot(g.v, g.v));
  int it = 0;
  if (history) history[0] = res;
  if (res > abs_tol)
    {
      precond(h.v, g.v);
      vec_equ(d.v, -1.0, h.v);
      double gh = vec_dot(g.v, h.v);
      for (it = 1; it <= max_iter; ++it)
        {
          laplace_vmult(op, h.v->p, d.v->p, false);
          const double alpha = gh / vec_dot(d.v, h.v);
          vec_sadd(x, 1.0, alpha, d.v);
          res = std::sqrt(vec_add_and_dot(g.v, alpha, h.v, g.v));
          if (history) history[it] = res;
          if (res <= abs_tol) break;
          precond(h.v, g.v);
          const double gh_new = vec_dot(g.v, h.v), beta = gh_new / gh;
          gh = gh_new;
          vec_sadd(d.v, beta, -1.0, h.v);
        }
      if (it > max_iter) it = max_iter;
    }
  if (iters) *iters = it;
  if (last_residual) *last_residual = res;
}

}  // namespace

struct mfg_cheb
{
  mfg_laplace *op = nullptr;
  int          degree = 0;
  double       smoothing_range = 0, lambda_max = 0, lambda_min = 0, theta = 0, delta = 0;
  int          eig_iterations_done = 0;
  Vec          d, t;  // update1, update2

  void update(mfg_vec *x, const mfg_vec *b, double f1, double f2, bool zero_start, bool first)
  {
    const mfg_mf *mf = op->mf;
    const size_t  n = mf->n_dofs;
    const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)op->ctx->sm_count * 16));
    cudaStream_t  s = op->ctx->stream;
    if (mf->dt == MFG_F64)
      cheb_update<double><<<nb, 256, 0, s>>>((double *)x->p, (double *)d.v->p, (const double *)t.v->p, (const double *)b->p, (const double *)op->inv_diag->p, n, f1, f2, zero_start, first);
    else
      cheb_update<float><<<nb, 256, 0, s>>>((float *)x->p, (float *)d.v->p, (const float *)t.v->p, (const float *)b->p, (const float *)op->inv_diag->p, n, (float)f1, (float)f2, zero_start, first);
    MFG_CUDA_LAST();
  }
  // PreconditionChebyshev::vmult (zero_start) / ::step
  void apply(mfg_vec *x, const mfg_vec *b, bool zero_start)
  {
    double rhok = delta / theta;
    const double sigma = theta / delta;
    if (!zero_start) laplace_vmult(op, t.v->p, x->p, false);
    update(x, b, 0.0, 1.0 / theta, zero_start, true);
    for (int k = 0; k < degree; ++k)
      {
        laplace_vmult(op, t.v->p, x->p, false);
        const double rhokp = 1.0 / (2.0 * sigma - rhok), f1 = rhokp * rhok, f2 = 2.0 * rhokp / delta;
        rhok = rhokp;
        update(x, b, f1, f2, false, false);
      }
  }
  // eigenvalue estimate: CG (SolverCG control flow) on Dinv A, Lanczos matrix from its coefficients
  void estimate(int n_iterations, double eig_cg_residual)
  {
    mfg_ctx *ctx = op->ctx;
    const mfg_mf *mf = op->mf;
    const size_t n = mf->n_dofs;
    Vec g(ctx, mf->dt, n), h(ctx, mf->dt, n), dd(ctx, mf->dt, n), rhs(ctx, mf->dt, n);
    vec_fill(rhs.v, 1.0 / std::sqrt((double)n));
    if (n) MFG_CUDA(cudaMemsetAsync(rhs.v->p, 0, rhs.v->esize(), ctx->stream));  // entry 0 = 0: triggers the high frequencies
    vec_equ(g.v, -1.0, rhs.v);                                                    // x = 0: g = -b
    auto precondition = [&]() { MG_CHECK(mfg_vec_copy(h.v, g.v)); vec_scale(h.v, op->inv_diag.get()); };
    precondition();
    vec_equ(dd.v, -1.0, h.v);
    double gh = vec_dot(g.v, h.v), res = std::sqrt(vec_dot(g.v, g.v));
    std::vector<double> alphas, betas;
    for (int it = 1; it <= n_iterations && res > eig_cg_residual; ++it)
      {
        laplace_vmult(op, h.v->p, dd.v->p, false);
        const double alpha = gh / vec_dot(dd.v, h.v);
        alphas.push_back(alpha);
        res = std::sqrt(vec_add_and_dot(g.v, alpha, h.v, g.v));
        precondition();
        const double gh_new = vec_dot(g.v, h.v), beta = gh_new / gh;
        gh = gh_new;
        betas.push_back(beta);
        vec_sadd(dd.v, beta, -1.0, h.v);
      }
    eig_iterations_done = (int)alphas.size();
    if (alphas.empty()) { lambda_max = lambda_min = 1.0; return; }
    const size_t k = alphas.size();
    std::vector<double> a(k), b(k > 1 ? k - 1 : 0);
    for (size_t j = 0; j < k; ++j)
      {
        a[j] = 1.0 / alphas[j] + (j ? betas[j - 1] / alphas[j - 1] : 0.0);
        if (j + 1 < k) b[j] = std::sqrt(betas[j]) / alphas[j];
      }
    lambda_max = tridiag_eigenvalue(a, b, (int)k - 1);
    lambda_min = tridiag_eigenvalue(a, b, 0);
  }
};

struct mfg_mg
{
  mfg_ctx *ctx = nullptr;
  mfg_dtype dt = MFG_F64;
  int min_level = 0, max_level = 0;
  std::vector<mfg_mesh *>    meshes;
  std::vector<mfg_laplace *> ops;
  std::vector<mfg_mgt *>     transfers;   // [l]: level l-1 -> l (l > min_level)
  std::vector<std::unique_ptr<mfg_cheb>> smoothers;
  std::vector<Vec> x, b, t;
  long coarse_iterations = 0;
  int  L(int level) const { return level - min_level; }
  ~mfg_mg()
  {
    smoothers.clear();
    for (auto *p : transfers) if (p) mfg_mgt_destroy(p);
    for (auto *p : ops) if (p) mfg_laplace_destroy(p);
    for (auto *p : meshes) if (p) mfg_mesh_destroy(p);
  }
  void cycle(int level)  // Multigrid::level_v_step
  {
    const int l = L(level);
    mfg_laplace *op = ops[l];
    if (level == min_level)
      {
        vec_fill(x[l].v, 0.0);
        const double bn = std::sqrt(vec_dot(b[l].v, b[l].v));
        int its = 0;
        MG_CHECK(mfg_solver_cg(op, x[l].v, b[l].v, 1e-10 * std::max(bn, 1e-300), 10000, 0, &its, nullptr, nullptr));
        coarse_iterations += its;
        return;
      }
    smoothers[l]->apply(x[l].v, b[l].v, true);                       // pre-smoothing from a zero guess
    laplace_vmult(op, t[l].v->p, x[l].v->p, false);
    vec_sadd(t[l].v, -1.0, 1.0, b[l].v);                             // t = b - A x
    vec_fill(b[l - 1].v, 0.0);
    MG_CHECK(mfg_mgt_restrict_and_add(transfers[l], b[l - 1].v, t[l].v));
    cycle(level - 1);
    MG_CHECK(mfg_mgt_prolongate(transfers[l], t[l].v, x[l - 1].v));
    vec_sadd(x[l].v, 1.0, 1.0, t[l].v);
    smoothers[l]->apply(x[l].v, b[l].v, false);                      // post-smoothing
  }
  void vmult(mfg_vec *dst, const mfg_vec *src)  // PreconditionMG::vmult: copy_to_mg, cycle, copy_from_mg
  {
    const int top = L(max_level);
    MG_CHECK(mfg_vec_copy(b[top].v, src));
    cycle(max_level);
    MG_CHECK(mfg_vec_copy(dst, x[top].v));
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Multigrid with local smoothing on an adaptively refined mesh.  Host hierarchy: csrc/adaptive_mesh.cu (mfg_amesh_build_mg);
// checker: oracle/adaptive_mg.py.  Every device operation below is a kernel the globally refined path already uses
// (level operators from explicit arrays, the transfer kernel with a weight table, indexed copies, BLAS-1).
// ---------------------------------------------------------------------------------------------------------------------
struct mfg_amg
{
  mfg_ctx *ctx = nullptr;
  mfg_dtype dt = MFG_F64;
  int min_level = 0, max_level = 0;
  mfg_laplace *active_op = nullptr;
  struct Level
  {
    mfg_laplace *op = nullptr, *raw = nullptr;   // level operator (boundary + edge constrained); the same cell loop without constraints
    mfg_ch      *ch = nullptr;                   // borrowed from op: constrained + edge lists
    mfg_mgt     *transfer = nullptr;             // level-1 -> level
    std::unique_ptr<mfg_cheb> smoother;
    Vec          x, b, t, w1, w2, w3;
    DevBuf<uint32_t> copy_global, copy_level;
    size_t       n = 0, n_edge = 0;
  };
  std::vector<Level> lv;
  long coarse_iterations = 0;
  Level &L(int level) { return lv[level - min_level]; }
  ~mfg_amg()
  {
    for (auto &l : lv)
      {
        l.smoother.reset();
        if (l.transfer) mfg_mgt_destroy(l.transfer);
        if (l.raw) mfg_laplace_destroy(l.raw);
        if (l.op) mfg_laplace_destroy(l.op);
      }
    if (active_op) mfg_laplace_destroy(active_op);
  }
  // dst = edge rows of A_raw (src with the constrained entries zeroed), zero elsewhere  (laplace_operator_gpu.h:306-331)
  void interface_down(Level &l, mfg_vec *dst, const mfg_vec *src)
  {
    vec_fill(dst, 0.0);
    if (!l.n_edge) return;
    MG_CHECK(mfg_vec_copy(l.w1.v, src));
    ch_set(l.ch, l.w1.v, 0.0);
    laplace_vmult(l.raw, l.w2.v->p, l.w1.v->p, false);
    vec_copy_with_indices(dst, l.w2.v, l.ch->edge.p, l.ch->edge.p, l.ch->edge.n);
  }
  // dst = A_raw (src restricted to the edge DoFs), constrained rows zeroed  (laplace_operator_gpu.h:333-352)
  void interface_up(Level &l, mfg_vec *dst, const mfg_vec *src)
  {
    if (!l.n_edge) { vec_fill(dst, 0.0); return; }
    vec_fill(l.w1.v, 0.0);
    vec_copy_with_indices(l.w1.v, src, l.ch->edge.p, l.ch->edge.p, l.ch->edge.n);
    laplace_vmult(l.raw, dst->p, l.w1.v->p, false);
    ch_set(l.ch, dst, 0.0);
  }
  void cycle(int level)  // Multigrid::level_v_step with edge_out = edge_in = the interface operators
  {
    Level &l = L(level);
    if (level == min_level)
      {
        vec_fill(l.x.v, 0.0);
        const double bn = std::sqrt(vec_dot(l.b.v, l.b.v));
        int its = 0;
        // (poisson_mg.cu:73-80: reduction 1e-10; single precision cannot resolve that, so 1e-5 there)
        MG_CHECK(mfg_solver_cg(l.op, l.x.v, l.b.v, (dt == MFG_F64 ? 1e-10 : 1e-5) * std::max(bn, 1e-300), 10000, 0, &its, nullptr, nullptr));
        coarse_iterations += its;
        return;
      }
    Level &lc = L(level - 1);
    l.smoother->apply(l.x.v, l.b.v, true);                          // pre-smoothing from a zero guess
    laplace_vmult(l.op, l.t.v->p, l.x.v->p, false);                 // t = A x
    if (l.n_edge)
      {
        interface_down(l, l.w3.v, l.x.v);                           // edge_out->vmult_add
        vec_sadd(l.t.v, 1.0, 1.0, l.w3.v);
      }
    vec_sadd(l.t.v, -1.0, 1.0, l.b.v);                              // t = b - t
    MG_CHECK(mfg_mgt_restrict_and_add(l.transfer, lc.b.v, l.t.v));  // the coarse defect already holds copy_to_mg's part
    cycle(level - 1);
    MG_CHECK(mfg_mgt_prolongate(l.transfer, l.t.v, lc.x.v));
    vec_sadd(l.x.v, 1.0, 1.0, l.t.v);
    if (l.n_edge)
      {
        interface_up(l, l.t.v, l.x.v);                              // edge_in->Tvmult
        vec_sadd(l.b.v, 1.0, -1.0, l.t.v);                          // defect -= t
      }
    l.smoother->apply(l.x.v, l.b.v, false);                         // post-smoothing
  }
  void vmult(mfg_vec *dst, const mfg_vec *src)  // PreconditionMG::vmult
  {
    for (auto &l : lv)                                              // copy_to_mg (mg_transfer_matrix_free_gpu.cu:688-727)
      {
        vec_fill(l.b.v, 0.0);
        vec_copy_with_indices(l.b.v, src, l.copy_level.p, l.copy_global.p, l.copy_global.n);
      }
    cycle(max_level);
    vec_fill(dst, 0.0);                                             // copy_from_mg (.cu:731-757)
    for (auto &l : lv) vec_copy_with_indices(dst, l.x.v, l.copy_global.p, l.copy_level.p, l.copy_global.n);
  }
};

extern "C" {

int mfg_chebyshev_create(mfg_laplace *op, int degree, double smoothing_range, int eig_cg_n_iterations, mfg_cheb **out)
{
  return guarded([&] {
    MFG_REQUIRE(op && out, "null argument");
    MFG_REQUIRE(degree >= 0 && eig_cg_n_iterations >= 0, "degree and eig_cg_n_iterations must be non-negative");
    std::unique_ptr<mfg_cheb> c(new mfg_cheb);
    c->op = op; c->degree = degree; c->smoothing_range = smoothing_range;
    if (!op->diagonal_is_available) laplace_compute_diagonal(op);
    c->d = Vec(op->ctx, op->mf->dt, op->mf->n_dofs);
    c->t = Vec(op->ctx, op->mf->dt, op->mf->n_dofs);
    if (eig_cg_n_iterations > 0) c->estimate(eig_cg_n_iterations, 1e-2);
    else c->lambda_max = c->lambda_min = 1.0;  // (AdditionalData::max_eigenvalue default)
    const double beta = 1.2 * c->lambda_max;
    const double alpha = smoothing_range > 1.0 ? c->lambda_max / smoothing_range : std::min(0.9 * c->lambda_max, c->lambda_min);
    c->delta = 0.5 * (beta - alpha); c->theta = 0.5 * (beta + alpha);
    *out = c.release();
  });
}
int mfg_chebyshev_destroy(mfg_cheb *c) { return guarded([&] { delete c; }); }
int mfg_chebyshev_vmult(mfg_cheb *c, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    MFG_REQUIRE(c && dst && src && dst != src, "null or aliased argument");
    MFG_REQUIRE(dst->n == c->op->mf->n_dofs && src->n == dst->n && dst->dt == c->op->mf->dt && src->dt == dst->dt, "vector does not fit the operator");
    c->apply(dst, src, true);
  });
}
int mfg_chebyshev_step(mfg_cheb *c, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    MFG_REQUIRE(c && dst && src && dst != src, "null or aliased argument");
    MFG_REQUIRE(dst->n == c->op->mf->n_dofs && src->n == dst->n && dst->dt == c->op->mf->dt && src->dt == dst->dt, "vector does not fit the operator");
    c->apply(dst, src, false);
  });
}
// the fused vector kernel of one Chebyshev product on its own (PreconditionChebyshev::vector_updates), for callers that own the
// operator product -- the multigrid over the box partition, whose A x includes the interface exchange
int mfg_vec_chebyshev_update(mfg_ctx *ctx, mfg_vec *x, mfg_vec *d, const mfg_vec *ax, const mfg_vec *b, const mfg_vec *dinv, double f1, double f2,
                             int zero_start, int first)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && x && d && b && dinv && (ax || zero_start), "null argument");
    const size_t n = x->n;
    MFG_REQUIRE(d->n == n && b->n == n && dinv->n == n && (!ax || ax->n == n), "vector sizes differ");
    MFG_REQUIRE(d->dt == x->dt && b->dt == x->dt && dinv->dt == x->dt && (!ax || ax->dt == x->dt), "vector types differ");
    MFG_REQUIRE(x != d && x != b && d != b, "aliased argument");
    if (!n) return;
    const unsigned nb = (unsigned)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 16));
    cudaStream_t   s = ctx->stream;
    const void    *axp = ax ? ax->p : b->p;   // (never read with zero_start)
    if (x->dt == MFG_F64)
      cheb_update<double><<<nb, 256, 0, s>>>((double *)x->p, (double *)d->p, (const double *)axp, (const double *)b->p, (const double *)dinv->p, n, f1, f2,
                                            zero_start != 0, first != 0);
    else
      cheb_update<float><<<nb, 256, 0, s>>>((float *)x->p, (float *)d->p, (const float *)axp, (const float *)b->p, (const float *)dinv->p, n, (float)f1,
                                           (float)f2, zero_start != 0, first != 0);
    MFG_CUDA_LAST();
  });
}
int mfg_chebyshev_info(const mfg_cheb *c, double *lambda_max, double *lambda_min, double *theta, double *delta, int *eig_iterations)
{
  return guarded([&] {
    MFG_REQUIRE(c, "null argument");
    if (lambda_max) *lambda_max = c->lambda_max;
    if (lambda_min) *lambda_min = c->lambda_min;
    if (theta) *theta = c->theta;
    if (delta) *delta = c->delta;
    if (eig_iterations) *eig_iterations = c->eig_iterations_done;
  });
}

int mfg_mg_create(mfg_ctx *ctx, int dim, int degree, int min_level, int max_level, mfg_dtype dt, double left, double right, int smoother_degree,
                  double smoothing_range, int eig_cg_n_iterations, mfg_mg **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out, "null argument");
    MFG_REQUIRE(min_level >= 0 && max_level >= min_level, "levels must satisfy 0 <= min_level <= max_level");
    std::unique_ptr<mfg_mg> mg(new mfg_mg);
    mg->ctx = ctx; mg->dt = dt; mg->min_level = min_level; mg->max_level = max_level;
    const int nl = max_level - min_level + 1;
    mg->meshes.assign(nl, nullptr); mg->ops.assign(nl, nullptr); mg->transfers.assign(nl, nullptr);
    mg->smoothers.resize(nl);
    for (int l = 0; l < nl; ++l)
      {
        MG_CHECK(mfg_mesh_hyper_cube(ctx, dim, degree, min_level + l, left, right, &mg->meshes[l]));
        MG_CHECK(mfg_laplace_create(ctx, mg->meshes[l], dt, MFG_SCATTER_ATOMIC, &mg->ops[l]));
        if (l > 0) MG_CHECK(mfg_mgt_build(ctx, mg->meshes[l - 1], mg->meshes[l], dt, &mg->transfers[l]));
        const size_t n = mg->meshes[l]->n_dofs;
        mg->x.emplace_back(ctx, dt, n); mg->b.emplace_back(ctx, dt, n); mg->t.emplace_back(ctx, dt, n);
        if (l > 0)
          {
            mfg_cheb *c = nullptr;
            MG_CHECK(mfg_chebyshev_create(mg->ops[l], smoother_degree, smoothing_range, eig_cg_n_iterations, &c));
            mg->smoothers[l].reset(c);
          }
      }
    *out = mg.release();
  });
}
int mfg_mg_destroy(mfg_mg *mg) { return guarded([&] { delete mg; }); }
int mfg_mg_vcycle(mfg_mg *mg, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    MFG_REQUIRE(mg && dst && src, "null argument");
    const size_t n = mg->meshes.back()->n_dofs;
    MFG_REQUIRE(dst->n == n && src->n == n && dst->dt == mg->dt && src->dt == mg->dt, "vector does not fit the finest level");
    mg->vmult(dst, src);
  });
}
int mfg_mg_level_operator(mfg_mg *mg, int level, mfg_laplace **op)
{
  return guarded([&] {
    MFG_REQUIRE(mg && op && level >= mg->min_level && level <= mg->max_level, "bad level");
    *op = mg->ops[mg->L(level)];
  });
}
int mfg_mg_info(const mfg_mg *mg, int level, double *lambda_max, long *coarse_iterations, size_t *n_dofs)
{
  return guarded([&] {
    MFG_REQUIRE(mg && level >= mg->min_level && level <= mg->max_level, "bad level");
    const int l = mg->L(level);
    if (lambda_max) *lambda_max = mg->smoothers[l] ? mg->smoothers[l]->lambda_max : 0.0;
    if (coarse_iterations) *coarse_iterations = mg->coarse_iterations;
    if (n_dofs) *n_dofs = mg->meshes[l]->n_dofs;
  });
}
// SolverCG preconditioned by one V-cycle per iteration (poisson_mg.cu:504-518), on the finest level operator
int mfg_mg_solve_cg(mfg_mg *mg, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int *iters, double *last_residual, double *history)
{
  return guarded([&] {
    MFG_REQUIRE(mg && x && b, "null argument");
    mfg_laplace *op = mg->ops.back();
    const size_t n = op->mf->n_dofs;
    MFG_REQUIRE(x->n == n && b->n == n && x->dt == mg->dt && b->dt == mg->dt, "vector does not fit the finest level");
    solve_cg_preconditioned(mg->ctx, mg->dt, op, [&](mfg_vec *h, const mfg_vec *g) { mg->vmult(h, g); }, x, b, abs_tol, max_iter, iters, last_residual, history);
  });
}

// ---- multigrid on adaptively refined meshes ------------------------------------------------------------------------------
int mfg_amg_create(mfg_ctx *ctx, mfg_amesh *am, int min_level, mfg_dtype dt, int smoother_degree, double smoothing_range, int eig_cg_n_iterations,
                   mfg_amg **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && am && out, "null argument");
    int dim = 0, degree = 0, n_levels = 0;
    double left = 0, right = 0;
    MG_CHECK(mfg_amesh_info(am, &dim, &degree, &left, &right, &n_levels, nullptr));
    MG_CHECK(mfg_amesh_build_mg(am, min_level));
    std::unique_ptr<mfg_amg> mg(new mfg_amg);
    mg->ctx = ctx; mg->dt = dt; mg->min_level = min_level; mg->max_level = n_levels - 1;
    MG_CHECK(mfg_laplace_create_from_amesh(ctx, am, dt, &mg->active_op));
    mg->lv.resize(n_levels - min_level);
    const uint32_t npc = ipow(degree + 1, dim), nF = ipow(2 * degree + 1, dim), n3 = ipow(3, dim);
    uint32_t n_dofs_below = 0;
    for (int level = min_level; level < n_levels; ++level)
      {
        mfg_amg::Level &l = mg->L(level);
        uint32_t sz[6];
        MG_CHECK(mfg_amesh_mg_level_sizes(am, level, sz));
        const uint32_t nc = sz[0], nd = sz[1], nb = sz[2], ne = sz[3], nblk = sz[4], ncp = sz[5];
        std::vector<uint32_t> l2g((size_t)nc * npc), boundary(nb), edge(ne), cg(ncp), cl(ncp), cidx((size_t)nblk * npc), fidx((size_t)nblk * nF);
        std::vector<double>   coef((size_t)nc * npc), w((size_t)nblk * n3);
        MG_CHECK(mfg_amesh_mg_level_get(am, level, l2g.data(), boundary.data(), edge.data(), coef.data(), cg.data(), cl.data(), cidx.data(), fidx.data(), w.data()));
        l.n = nd; l.n_edge = ne;
        // LaplaceOperatorGpu::reinit(dof_handler, mg_constrained_dofs, level): the level's cells, no hanging nodes; constraints =
        // boundary + refinement-edge DoFs, edge list = the refinement-edge DoFs (constraint_handler_gpu.cu:99-123)
        std::vector<double> inv_jac(nc, (double)((uint64_t)1 << level) / (right - left));
        std::vector<uint32_t> constrained(nb + ne);
        std::merge(boundary.begin(), boundary.end(), edge.begin(), edge.end(), constrained.begin());
        constrained.erase(std::unique(constrained.begin(), constrained.end()), constrained.end());
        mfg_mf_desc d;
        std::memset(&d, 0, sizeof(d));
        d.dim = dim; d.degree = degree; d.dtype = dt; d.n_cells = nc; d.n_dofs = nd; d.loc2glob = l2g.data();
        d.geometry = MFG_GEOM_UNIFORM; d.inv_jac = inv_jac.data(); d.scatter = MFG_SCATTER_ATOMIC;
        {
          std::unique_ptr<mfg_mf> mf(mf_from_desc(ctx, d));
          std::unique_ptr<mfg_ch> ch(ch_create(ctx, dt, constrained.data(), constrained.size(), edge.data(), edge.size()));
          l.op = laplace_from_arrays(ctx, mf.get(), ch.get(), coef.data());
          l.op->owns_mf = true; l.op->owns_ch = true;
          l.ch = ch.get();
          mf.release(); ch.release();
        }
        if (ne)
          {
            // the cell loop without the constraint handler, for the interface operators (they read and write edge rows)
            std::unique_ptr<mfg_mf> mf(mf_from_desc(ctx, d));
            std::unique_ptr<mfg_ch> ch(ch_create(ctx, dt, nullptr, 0, nullptr, 0));
            l.raw = laplace_from_arrays(ctx, mf.get(), ch.get(), coef.data());
            l.raw->owns_mf = true; l.raw->owns_ch = true;
            l.raw->variant = 1;  // column kernel: no second copy of the index map in consumption order
            mf.release(); ch.release();
          }
        l.x = Vec(ctx, dt, nd); l.b = Vec(ctx, dt, nd); l.t = Vec(ctx, dt, nd);
        if (ne) { l.w1 = Vec(ctx, dt, nd); l.w2 = Vec(ctx, dt, nd); l.w3 = Vec(ctx, dt, nd); }
        l.copy_global.upload(cg.data(), cg.size(), ctx->stream);
        l.copy_level.upload(cl.data(), cl.size(), ctx->stream);
        if (level > min_level)
          {
            MG_CHECK(mfg_mgt_build_from_blocks(ctx, dt, dim, degree, nblk, cidx.data(), fidx.data(), w.data(), n_dofs_below, nd, &l.transfer));
            mfg_cheb *c = nullptr;
            MG_CHECK(mfg_chebyshev_create(l.op, smoother_degree, smoothing_range, eig_cg_n_iterations, &c));
            l.smoother.reset(c);
          }
        n_dofs_below = nd;
      }
    *out = mg.release();
  });
}
int mfg_amg_destroy(mfg_amg *mg) { return guarded([&] { delete mg; }); }
static void amg_check_active(const mfg_amg *mg, const mfg_vec *a, const mfg_vec *b)
{
  const size_t n = mg->active_op->mf->n_dofs;
  MFG_REQUIRE(a && b && a != b && a->n == n && b->n == n && a->dt == mg->dt && b->dt == mg->dt, "vector does not fit the active mesh");
}
static mfg_amg::Level &amg_level(mfg_amg *mg, int level)
{
  MFG_REQUIRE(mg && level >= mg->min_level && level <= mg->max_level, "bad level");
  return mg->L(level);
}
static void amg_check_level(const mfg_amg *mg, const mfg_amg::Level &l, const mfg_vec *v)
{
  MFG_REQUIRE(v && v->n == l.n && v->dt == mg->dt, "vector does not fit the level");
}
int mfg_amg_vcycle(mfg_amg *mg, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] { MFG_REQUIRE(mg, "null argument"); amg_check_active(mg, dst, src); mg->vmult(dst, src); });
}
int mfg_amg_active_operator(mfg_amg *mg, mfg_laplace **op) { return guarded([&] { MFG_REQUIRE(mg && op, "null argument"); *op = mg->active_op; }); }
int mfg_amg_level_operator(mfg_amg *mg, int level, mfg_laplace **op)
{
  return guarded([&] { MFG_REQUIRE(op, "null argument"); *op = amg_level(mg, level).op; });
}
int mfg_amg_vmult_interface_down(mfg_amg *mg, int level, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    mfg_amg::Level &l = amg_level(mg, level);
    amg_check_level(mg, l, dst); amg_check_level(mg, l, src);
    MFG_REQUIRE(dst != src, "aliased arguments");
    mg->interface_down(l, dst, src);
  });
}
int mfg_amg_vmult_interface_up(mfg_amg *mg, int level, mfg_vec *dst, const mfg_vec *src)
{
  return guarded([&] {
    mfg_amg::Level &l = amg_level(mg, level);
    amg_check_level(mg, l, dst); amg_check_level(mg, l, src);
    MFG_REQUIRE(dst != src, "aliased arguments");
    mg->interface_up(l, dst, src);
  });
}
int mfg_amg_prolongate(mfg_amg *mg, int level, mfg_vec *dst_fine, const mfg_vec *src_coarse)
{
  return guarded([&] {
    mfg_amg::Level &l = amg_level(mg, level);
    MFG_REQUIRE(level > mg->min_level, "no transfer into the coarsest level");
    MG_CHECK(mfg_mgt_prolongate(l.transfer, dst_fine, src_coarse));
  });
}
int mfg_amg_restrict_and_add(mfg_amg *mg, int level, mfg_vec *dst_coarse, const mfg_vec *src_fine)
{
  return guarded([&] {
    mfg_amg::Level &l = amg_level(mg, level);
    MFG_REQUIRE(level > mg->min_level, "no transfer into the coarsest level");
    MG_CHECK(mfg_mgt_restrict_and_add(l.transfer, dst_coarse, src_fine));
  });
}
int mfg_amg_copy_to_level(mfg_amg *mg, int level, mfg_vec *dst_level, const mfg_vec *src_active)
{
  return guarded([&] {
    mfg_amg::Level &l = amg_level(mg, level);
    amg_check_level(mg, l, dst_level);
    MFG_REQUIRE(src_active && src_active->n == mg->active_op->mf->n_dofs, "source does not fit the active mesh");
    vec_fill(dst_level, 0.0);
    vec_copy_with_indices(dst_level, src_active, l.copy_level.p, l.copy_global.p, l.copy_global.n);
  });
}
int mfg_amg_copy_from_level(mfg_amg *mg, int level, mfg_vec *dst_active, const mfg_vec *src_level)
{
  return guarded([&] {
    mfg_amg::Level &l = amg_level(mg, level);
    amg_check_level(mg, l, src_level);
    MFG_REQUIRE(dst_active && dst_active->n == mg->active_op->mf->n_dofs, "destination does not fit the active mesh");
    vec_copy_with_indices(dst_active, src_level, l.copy_global.p, l.copy_level.p, l.copy_global.n);
  });
}
int mfg_amg_info(const mfg_amg *mg, int level, double *lambda_max, long *coarse_iterations, size_t *n_dofs, size_t *n_edge)
{
  return guarded([&] {
    MFG_REQUIRE(mg && level >= mg->min_level && level <= mg->max_level, "bad level");
    const mfg_amg::Level &l = mg->lv[level - mg->min_level];
    if (lambda_max) *lambda_max = l.smoother ? l.smoother->lambda_max : 0.0;
    if (coarse_iterations) *coarse_iterations = mg->coarse_iterations;
    if (n_dofs) *n_dofs = l.n;
    if (n_edge) *n_edge = l.n_edge;
  });
}
int mfg_amg_solve_cg(mfg_amg *mg, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int *iters, double *last_residual, double *history)
{
  return guarded([&] {
    MFG_REQUIRE(mg, "null argument");
    amg_check_active(mg, x, b);
    solve_cg_preconditioned(mg->ctx, mg->dt, mg->active_op, [&](mfg_vec *h, const mfg_vec *g) { mg->vmult(h, g); }, x, b, abs_tol, max_iter, iters,
                            last_residual, history);
  });
}

}  // extern "C"
