// partitioned_mg.h -- geometric multigrid over the box partition, host code in C++ on the facade classes (the Python form is
// dealii_cuda_b200/partitioned_mg.py; algorithm: poisson_mg.cu:430-552 / bmop_mg.cu:60-82 with the data layout of distributed.h).
//
// Every level l = min_level..max_level is partitioned like the finest one: a rank's box on level l - 1 is its box on level l coarsened
// once, so the transfer between two levels is local to a box (MGTransferMatrixFreeGpu on the box's two meshes) and the only
// communication is the interface exchange of the operator:
//   level operator = local cell loop + exchange_add;   smoother = PreconditionChebyshev's recurrence with the exchanged diagonal
//   (mfg_vec_chebyshev_update, one kernel per product and box), eigenvalue estimate by CG / Lanczos with owned-DoF inner products;
//   restriction = fine residual x 1 / (number of boxes holding the DoF), local restrict_and_add, exchange_add of the coarse result;
//   prolongation = local;   coarse problem = CG over the partition (MGCoarseIterative, bmop_mg.cu:65-82).
// LocalWorldLevel holds ALL boxes of a level in one process and stages the exchange through the host: the form that runs on one GPU
// (and in the CPU emulation build of the tests).  With one box per process the same algorithm runs with the exchange of
// InterfaceExchange::push over peer memory; that transport is the caller's (distributed.py does it with torch.distributed).
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <memory>
#include <vector>
#include "distributed.h"

namespace dealii_cuda_b200 {

template <typename Number> using Field = std::vector<GpuVector<Number>>;   // one vector per box

template <int dim, int fe_degree, typename Number> class LocalWorldLevel
{
public:
  struct Part
  {
    std::unique_ptr<BoxPartition>                              partition;
    std::unique_ptr<HyperCubeMesh<dim>>                        mesh;
    std::unique_ptr<LaplaceOperatorGpu<dim, fe_degree, Number>> op;
    std::unique_ptr<ExchangePlan>                              plan;
    std::unique_ptr<InterfaceExchange<Number>>                 exchange;
    GpuVector<Number>                                          send, recv, owned;
    unsigned int                                               n = 0;
  };

  LocalWorldLevel(int world, int level, bool strong = false, double left = -1., double right = 1.) : world_(world), level_(level)
  {
    parts_.resize(world);
    for (int rank = 0; rank < world; ++rank)
      {
        Part &p = parts_[rank];
        p.partition.reset(new BoxPartition(rank, world, dim, fe_degree, level, strong, left, right));
        p.mesh.reset(new HyperCubeMesh<dim>(fe_degree, p.partition->box()));
        p.op.reset(new LaplaceOperatorGpu<dim, fe_degree, Number>());
        p.op->reinit(*p.mesh);
        p.n = p.mesh->n_dofs();
        mfg_mesh *m = p.mesh->handle();
        p.plan.reset(new ExchangePlan(*p.partition, p.n, [m](const std::vector<uint32_t> &xyz) {
          std::vector<uint32_t> dofs(xyz.size() / 3);
          if (!dofs.empty()) check(mfg_mesh_lattice_to_dof(m, dofs.size(), xyz.data(), dofs.data()));
          return dofs;
        }));
        p.exchange.reset(new InterfaceExchange<Number>(*p.plan));
        p.send.reinit((unsigned int)std::max<size_t>(1, p.plan->n_send()));
        p.recv.reinit((unsigned int)std::max<size_t>(1, p.plan->n_send()));
        std::vector<Number> own(p.n);
        for (unsigned int i = 0; i < p.n; ++i) own[i] = p.plan->owned_mask[i] ? Number(1) : Number(0);
        p.owned = own;
      }
    n_global_ = parts_[0].partition->global_n_dofs();
    // 1 / (number of boxes holding the DoF), and the inverse of the exchanged diagonal
    inv_mult_ = new_field();
    for (auto &v : inv_mult_) v = Number(1);
    exchange_add(inv_mult_);
    for (auto &v : inv_mult_) v.invert();
    inv_diag_ = new_field();
    for (int a = 0; a < world; ++a)
      {
        parts_[a].op->compute_diagonal();
        inv_diag_[a] = parts_[a].op->get_diagonal_inverse()->get_vector();
        inv_diag_[a].invert();
      }
    exchange_add(inv_diag_);
    for (auto &v : inv_diag_) v.invert();
    tmp_ = new_field();
  }

  Field<Number> new_field() const
  {
    Field<Number> f;
    for (const Part &p : parts_) f.emplace_back(p.n);
    return f;
  }
  // compress(add) + update_ghost_values: every replica of an interface DoF ends with the sum of all partial sums (rank order)
  void exchange_add(Field<Number> &f) const
  {
    if (world_ == 1) return;
    std::vector<std::vector<Number>> sends(world_);
    for (int a = 0; a < world_; ++a)
      {
        const Part &p = parts_[a];
        if (p.plan->n_send()) p.exchange->pack(f[a], const_cast<GpuVector<Number> &>(p.send).getData());
        sends[a] = p.send.toVector();
      }
    for (int a = 0; a < world_; ++a)
      {
        const Part &p = parts_[a];
        if (!p.plan->n_send()) continue;
        // receive buffer of box a: the neighbours' blocks for a in ascending rank order (what an all-to-all delivers)
        std::vector<Number> recv;
        for (int b : p.plan->neighbors)
          {
            const ExchangePlan &pb = *parts_[b].plan;
            size_t              off = 0;
            for (int q : pb.neighbors) if (q < a) off += pb.splits[q];
            recv.insert(recv.end(), sends[b].begin() + off, sends[b].begin() + off + pb.splits[a]);
          }
        if (recv.size() != p.plan->n_send()) throw std::runtime_error("partitioned_mg: exchange plan of the boxes is not symmetric");
        GpuVector<Number> &r = const_cast<GpuVector<Number> &>(p.recv);
        r.fromHost(recv.data(), (unsigned int)recv.size());
        p.exchange->accumulate(f[a], r.getDataRO());
      }
  }
  void vmult(Field<Number> &dst, const Field<Number> &src) const
  {
    for (int a = 0; a < world_; ++a) parts_[a].op->vmult(dst[a], src[a]);
    exchange_add(dst);
  }
  // global inner product: every DoF counted in the box that owns it
  double dot(const Field<Number> &x, const Field<Number> &y) const
  {
    double s = 0;
    for (int a = 0; a < world_; ++a)
      {
        GpuVector<Number> &t = const_cast<GpuVector<Number> &>(tmp_[a]);
        t = x[a];
        t.scale(parts_[a].owned);
        double r = 0;
        check(mfg_vec_dot(t.handle(), y[a].handle(), &r));
        s += r;
      }
    return s;
  }
  const std::vector<Part> &parts() const { return parts_; }
  const Field<Number>     &inv_diag() const { return inv_diag_; }
  const Field<Number>     &inv_mult() const { return inv_mult_; }
  unsigned long long       n_global() const { return n_global_; }
  int                      world() const { return world_; }
  int                      level() const { return level_; }

private:
  int                world_, level_;
  std::vector<Part>  parts_;
  Field<Number>      inv_mult_, inv_diag_, tmp_;
  unsigned long long n_global_ = 0;
};

// SolverCG on fields (deal.II's control flow as in csrc/multigrid.cu); precond(dst, src) may be empty.  Returns the iteration count.
template <typename Level, typename Number>
int field_cg(const Level &L, Field<Number> &x, const Field<Number> &b, double abs_tol, int max_iter,
             const std::function<void(Field<Number> &, const Field<Number> &)> &precond = nullptr, double *last_residual = nullptr)
{
  Field<Number> g = L.new_field(), h = L.new_field(), d = L.new_field();
  const size_t  np = x.size();
  L.vmult(g, x);
  for (size_t a = 0; a < np; ++a) g[a].add(Number(-1), b[a]);
  double res = std::sqrt(L.dot(g, g));
  int    it = 0;
  if (res > abs_tol)
    {
      auto apply_precond = [&]() {
        if (precond) precond(h, g);
        else for (size_t a = 0; a < np; ++a) h[a] = g[a];
      };
      apply_precond();
      for (size_t a = 0; a < np; ++a) d[a].equ(Number(-1), h[a]);
      double gh = L.dot(g, h);
      for (it = 1; it <= max_iter; ++it)
        {
          L.vmult(h, d);
          const double alpha = gh / L.dot(d, h);
          for (size_t a = 0; a < np; ++a) { x[a].add((Number)alpha, d[a]); g[a].add((Number)alpha, h[a]); }
          res = std::sqrt(L.dot(g, g));
          if (res <= abs_tol) break;
          apply_precond();
          const double gh_new = L.dot(g, h), beta = gh_new / gh;
          gh = gh_new;
          for (size_t a = 0; a < np; ++a) d[a].sadd((Number)beta, Number(-1), h[a]);
        }
      if (it > max_iter) it = max_iter;
    }
  if (last_residual) *last_residual = res;
  return it;
}

// PreconditionChebyshev on Dinv A of a partitioned level (mfg_cheb of csrc/multigrid.cu on fields)
template <typename Level, typename Number> class PartitionedChebyshev
{
public:
  PartitionedChebyshev(const Level &L, int degree = 5, double smoothing_range = 15., int eig_cg_n_iterations = 15) : L_(L), degree_(degree)
  {
    d_ = L.new_field();
    t_ = L.new_field();
    if (eig_cg_n_iterations > 0) estimate(eig_cg_n_iterations, 1e-2);
    const double beta = 1.2 * lambda_max_;
    const double alpha = smoothing_range > 1. ? lambda_max_ / smoothing_range : std::min(0.9 * lambda_max_, lambda_min_);
    delta_ = 0.5 * (beta - alpha);
    theta_ = 0.5 * (beta + alpha);
  }
  // PreconditionChebyshev::vmult (zero_start) / ::step
  void apply(Field<Number> &x, const Field<Number> &b, bool zero_start)
  {
    double       rhok = delta_ / theta_;
    const double sigma = theta_ / delta_;
    if (!zero_start) L_.vmult(t_, x);
    update(x, b, 0., 1. / theta_, zero_start, true);
    for (int k = 0; k < degree_; ++k)
      {
        L_.vmult(t_, x);
        const double rhokp = 1. / (2. * sigma - rhok), f1 = rhokp * rhok, f2 = 2. * rhokp / delta_;
        rhok = rhokp;
        update(x, b, f1, f2, false, false);
      }
  }
  double lambda_max() const { return lambda_max_; }

private:
  void update(Field<Number> &x, const Field<Number> &b, double f1, double f2, bool zero_start, bool first)
  {
    for (size_t a = 0; a < x.size(); ++a)
      check(mfg_vec_chebyshev_update(default_context(), x[a].handle(), d_[a].handle(), t_[a].handle(), b[a].handle(), L_.inv_diag()[a].handle(), f1, f2,
                                     zero_start ? 1 : 0, first ? 1 : 0));
  }
  // deal.II's estimate: eig_cg_n_iterations steps of CG on Dinv A from the right-hand side 1 / sqrt(n) (entry 0 zeroed), eigenvalues of the
  // Lanczos matrix of the CG coefficients (largest one by bisection on the Sturm count)
  void estimate(int n_iterations, double eig_cg_residual)
  {
    Field<Number> g = L_.new_field(), h = L_.new_field(), dd = L_.new_field();
    const size_t  np = g.size();
    for (size_t a = 0; a < np; ++a)
      {
        std::vector<Number> v(g[a].size(), (Number)(-1. / std::sqrt((double)L_.n_global())));
        if (a == 0 && !v.empty()) v[0] = 0;
        g[a] = v;
      }
    auto precondition = [&]() { for (size_t a = 0; a < np; ++a) { h[a] = g[a]; h[a].scale(L_.inv_diag()[a]); } };
    precondition();
    for (size_t a = 0; a < np; ++a) dd[a].equ(Number(-1), h[a]);
    double              gh = L_.dot(g, h), res = std::sqrt(L_.dot(g, g));
    std::vector<double> alphas, betas;
    for (int it = 1; it <= n_iterations && res > eig_cg_residual; ++it)
      {
        L_.vmult(h, dd);
        const double alpha = gh / L_.dot(dd, h);
        alphas.push_back(alpha);
        for (size_t a = 0; a < np; ++a) g[a].add((Number)alpha, h[a]);
        res = std::sqrt(L_.dot(g, g));
        precondition();
        const double gh_new = L_.dot(g, h), beta = gh_new / gh;
        gh = gh_new;
        betas.push_back(beta);
        for (size_t a = 0; a < np; ++a) dd[a].sadd((Number)beta, Number(-1), h[a]);
      }
    if (alphas.empty()) return;
    const size_t        k = alphas.size();
    std::vector<double> diag(k), off(k > 1 ? k - 1 : 0);
    for (size_t j = 0; j < k; ++j)
      {
        diag[j] = 1. / alphas[j] + (j ? betas[j - 1] / alphas[j - 1] : 0.);
        if (j + 1 < k) off[j] = std::sqrt(betas[j]) / alphas[j];
      }
    auto count_below = [&](double x) {  // Sturm count: eigenvalues < x
      int    cnt = 0;
      double q = diag[0] - x;
      if (q < 0) ++cnt;
      for (size_t j = 1; j < k; ++j)
        {
          q = diag[j] - x - off[j - 1] * off[j - 1] / (q == 0. ? 1e-300 : q);
          if (q < 0) ++cnt;
        }
      return cnt;
    };
    auto kth = [&](int which) {
      double lo = diag[0], hi = diag[0];
      for (size_t j = 0; j < k; ++j)
        {
          const double r = (j ? std::fabs(off[j - 1]) : 0.) + (j + 1 < k ? std::fabs(off[j]) : 0.);
          lo = std::min(lo, diag[j] - r);
          hi = std::max(hi, diag[j] + r);
        }
      for (int i = 0; i < 200; ++i)
        {
          const double mid = 0.5 * (lo + hi);
          (count_below(mid) <= which ? lo : hi) = mid;
        }
      return 0.5 * (lo + hi);
    };
    lambda_max_ = kth((int)k - 1);
    lambda_min_ = kth(0);
  }

  const Level  &L_;
  int           degree_;
  double        lambda_max_ = 1., lambda_min_ = 1., theta_ = 1., delta_ = 1.;
  Field<Number> d_, t_;
};

// Multigrid::level_v_step over partitioned levels + the CG it preconditions
template <int dim, int fe_degree, typename Number> class PartitionedMultigrid
{
public:
  typedef LocalWorldLevel<dim, fe_degree, Number> Level;

  PartitionedMultigrid(int world, int min_level, int max_level, bool strong = false, int smoother_degree = 5, double smoothing_range = 15.,
                       int eig_cg_n_iterations = 15)
    : min_level_(min_level), max_level_(max_level)
  {
    if (min_level < 1 || max_level < min_level) throw std::runtime_error("PartitionedMultigrid: need 1 <= min_level <= max_level");
    for (int l = min_level; l <= max_level; ++l) levels_.emplace_back(new Level(world, l, strong));
    for (int l = min_level + 1; l <= max_level; ++l)
      {
        smoothers_.emplace_back(new PartitionedChebyshev<Level, Number>(level(l), smoother_degree, smoothing_range, eig_cg_n_iterations));
        transfers_.emplace_back();
        for (int a = 0; a < world; ++a)
          {
            transfers_.back().emplace_back(new MGTransferMatrixFreeGpu<dim, Number>());
            transfers_.back().back()->build({level(l - 1).parts()[a].mesh.get(), level(l).parts()[a].mesh.get()}, (unsigned int)(l - 1));
          }
      }
    for (int l = min_level; l <= max_level; ++l)
      {
        x_.push_back(level(l).new_field());
        b_.push_back(level(l).new_field());
        t_.push_back(level(l).new_field());
      }
  }
  const Level &finest() const { return *levels_.back(); }
  const Level &level(int l) const { return *levels_[l - min_level_]; }
  double       lambda_max(int l) const { return smoothers_[l - min_level_ - 1]->lambda_max(); }
  long         coarse_iterations() const { return coarse_iterations_; }

  // PreconditionMG::vmult: one V-cycle on the right-hand side src
  void vmult(Field<Number> &dst, const Field<Number> &src)
  {
    Field<Number> &top = b_.back();
    for (size_t a = 0; a < top.size(); ++a) top[a] = src[a];
    cycle(max_level_);
    for (size_t a = 0; a < dst.size(); ++a) dst[a] = x_.back()[a];
  }
  int solve_cg(Field<Number> &x, const Field<Number> &b, double abs_tol, int max_iter = 1000, double *last_residual = nullptr)
  {
    return field_cg<Level, Number>(finest(), x, b, abs_tol, max_iter, [this](Field<Number> &d, const Field<Number> &s) { vmult(d, s); }, last_residual);
  }

private:
  void cycle(int l)
  {
    const int      i = l - min_level_;
    const Level   &L = level(l);
    Field<Number> &x = x_[i], &b = b_[i], &t = t_[i];
    if (l == min_level_)
      {
        for (auto &v : x) v = Number(0);
        const double bn = std::sqrt(L.dot(b, b));
        coarse_iterations_ += field_cg<Level, Number>(L, x, b, (sizeof(Number) == 8 ? 1e-10 : 1e-4) * std::max(bn, 1e-300), 10000);
        return;
      }
    PartitionedChebyshev<Level, Number> &S = *smoothers_[i - 1];
    S.apply(x, b, true);                                            // pre-smoothing from a zero guess
    L.vmult(t, x);
    for (size_t a = 0; a < t.size(); ++a) t[a].sadd(Number(-1), Number(1), b[a]);   // t = b - A x
    Field<Number> &bc = b_[i - 1];
    for (auto &v : bc) v = Number(0);
    for (size_t a = 0; a < t.size(); ++a)
      {
        t[a].scale(L.inv_mult()[a]);                                // every fine DoF counted once
        transfers_[i - 1][a]->restrict_and_add((unsigned int)l, bc[a], t[a]);
      }
    level(l - 1).exchange_add(bc);
    cycle(l - 1);
    for (size_t a = 0; a < t.size(); ++a)
      {
        transfers_[i - 1][a]->prolongate((unsigned int)l, t[a], x_[i - 1][a]);
        x[a].add(t[a]);
      }
    S.apply(x, b, false);                                           // post-smoothing
  }

  int                                                                        min_level_, max_level_;
  std::vector<std::unique_ptr<Level>>                                        levels_;
  std::vector<std::unique_ptr<PartitionedChebyshev<Level, Number>>>          smoothers_;
  std::vector<std::vector<std::unique_ptr<MGTransferMatrixFreeGpu<dim, Number>>>> transfers_;
  std::vector<Field<Number>>                                                 x_, b_, t_;
  long                                                                       coarse_iterations_ = 0;
};

}  // namespace dealii_cuda_b200
