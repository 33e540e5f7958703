// solver.cu -- conjugate gradients on GpuVectors with the Laplace operator: the control flow of deal.II's
// SolverCG<GpuVector> as the reference instantiates it (poisson.cu:233-260).  The reference runs
// five BLAS-1 kernels per iteration (operator*, add, add_and_dot, DiagonalMatrix::vmult -> scale, sadd;
// gpu_vec.cu:306-617), each of which cudaMallocs and blocks on a D2H copy.  Here alpha, beta and the residual stay on the
// device, the host never waits inside the loop, and an iteration is THREE kernels:
//   cell kernel   h = A d (constrained rows included) and, in the same pass, the per-warp partial sums of d . h
//   cg_residual   alpha = g.z / d.h ; g += alpha h ; z = Minv g (into h) ; |g|^2, g.z ; convergence ; beta
//   cg_advance    x += alpha d ; d = beta d - z ; h = 0 -- the zero pass of the NEXT operator application
// 11 vector passes beside the operator's own (the reference: 15 + the operator's zero pass and two constraint kernels).
// Operators whose kernel cannot emit the dot product (column kernel, hanging nodes) use vmult + cg_dot instead.  Preconditioner: the inverse diagonal (PreconditionChebyshev with
// its default degree 0 is a scaled Jacobi step; the scaling does not change the CG iterates).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "operators.cuh"

namespace mfg {
namespace {

constexpr int TH = 256;

__device__ inline double warp_sum(double x)
{
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;
}
// block-wide sum of two values; result valid in thread 0
__device__ inline void block_sum2(double &a, double &b)
{
  __shared__ double sa[TH / 32], sb[TH / 32];
  a = warp_sum(a); b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32)
    {
      a = threadIdx.x < TH / 32 ? sa[threadIdx.x] : 0.0;
      b = threadIdx.x < TH / 32 ? sb[threadIdx.x] : 0.0;
      a = warp_sum(a); b = warp_sum(b);
    }
}

// Device-resident state of one solve: the scalars never travel to the host inside the loop.  Every kernel of an
// iteration reads what it needs from here, the last block of a reduction kernel (ticket counter) finishes the sum in
// a fixed order and updates it.  Once `converged_at` is set, the kernels of later iterations return at once, so the
// host may enqueue a few iterations ahead of what it has seen of the residual without changing the result.
struct CgState
{
  double gh;            // g . h of the current iterate (h = Minv g)
  double alpha, beta;
  double res;           // |g|
  double tol;
  int    converged_at;  // iteration at which |g| <= tol was met, -1 before
  unsigned ticket;
};

__device__ inline bool last_block_done(unsigned *ticket)
{
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  return last;
}
// sum of the per-block partials in block order by the calling (last) block; result valid in thread 0
__device__ inline void sum_partials2(const volatile double *partial, double &a, double &b)
{
  a = 0; b = 0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
  __syncthreads();
  block_sum2(a, b);
}

// alpha = g.h / d.h
template <typename T>
__global__ void cg_dot(const T *__restrict__ d, const T *__restrict__ h, size_t n, double *__restrict__ partial, CgState *st)
{
  if (st->converged_at >= 0) return;
  double s0 = 0, s1 = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride)  // four independent loads per vector in flight
    {
      const T d0 = d[i], d1 = d[i + stride], d2 = d[i + 2 * stride], d3 = d[i + 3 * stride];
      const T h0 = h[i], h1 = h[i + stride], h2 = h[i + 2 * stride], h3 = h[i + 3 * stride];
      s0 += (double)d0 * (double)h0 + (double)d1 * (double)h1;
      s1 += (double)d2 * (double)h2 + (double)d3 * (double)h3;
    }
  for (; i < n; i += stride) s0 += (double)d[i] * (double)h[i];
  s0 += s1; s1 = 0;
  block_sum2(s0, s1);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s0; partial[2 * blockIdx.x + 1] = 0; }
  if (last_block_done(&st->ticket))
    {
      sum_partials2(partial, s0, s1);
      if (threadIdx.x == 0) { st->alpha = st->gh / s0; st->ticket = 0; }
    }
}
// g += alpha h ; z = Minv .* g (z overwrites h) ; |g|, g.z ; convergence test ; beta = g.z / gh_old
// (first = true: alpha = 0, h holds nothing yet: only z and the sums, the start of the iteration)
// dotp != nullptr: alpha = gh / (sum of the n_dot partial sums of d . h the cell kernel left), summed by every block in the
// same fixed order, so that all blocks use the same bits
template <typename T>
__global__ void cg_residual(T *__restrict__ g, T *__restrict__ h, const T *__restrict__ minv, size_t n, double *__restrict__ partial, CgState *st,
                            int it, double *__restrict__ history, const double *__restrict__ dotp, unsigned n_dot)
{
  if (st->converged_at >= 0) return;
  const bool first = it == 0;
  __shared__ double s_alpha;
  if (dotp != nullptr && !first)
    {
      double a = 0, b = 0;
      for (unsigned i = threadIdx.x; i < n_dot; i += blockDim.x) a += dotp[i];
      block_sum2(a, b);
      if (threadIdx.x == 0) s_alpha = st->gh / a;
      __syncthreads();
    }
  const T alpha = first ? T(0) : (T)(dotp != nullptr ? s_alpha : st->alpha);
  double gg = 0, gz = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride)  // four independent elements per thread in flight
    {
      T gv[4], hv[4], mv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { gv[k] = g[i + k * stride]; hv[k] = first ? T(0) : h[i + k * stride]; mv[k] = minv ? minv[i + k * stride] : T(1); }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        {
          const T gi = first ? gv[k] : gv[k] + alpha * hv[k];
          if (!first) g[i + k * stride] = gi;
          const T zi = mv[k] * gi;
          h[i + k * stride] = zi;
          gg += (double)gi * (double)gi;
          gz += (double)gi * (double)zi;
        }
    }
  for (; i < n; i += stride)
    {
      const T gi = first ? g[i] : g[i] + alpha * h[i];
      if (!first) g[i] = gi;
      const T zi = minv ? minv[i] * gi : gi;
      h[i] = zi;
      gg += (double)gi * (double)gi;
      gz += (double)gi * (double)zi;
    }
  block_sum2(gg, gz);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = gg; partial[2 * blockIdx.x + 1] = gz; }
  if (last_block_done(&st->ticket))
    {
      sum_partials2(partial, gg, gz);
      if (threadIdx.x == 0)
        {
          const double res = sqrt(gg);
          st->res = res;
          if (history) history[it] = res;
          if (res <= st->tol) st->converged_at = it;
          st->beta = first ? 0.0 : gz / st->gh;
          if (dotp != nullptr && !first) st->alpha = s_alpha;  // (for cg_advance)
          st->gh = gz;
          st->ticket = 0;
        }
    }
}
// x += alpha d (also in the iteration that converged, as SolverCG updates the solution before the check) ;
// d = beta d - z ; zero_z: z = 0, the zero pass of the next operator application, whose cell kernel is launched as the
// programmatic dependent of this kernel
template <typename T> __global__ void cg_advance(T *__restrict__ x, T *__restrict__ d, T *__restrict__ z, size_t n, const CgState *st, int it, bool zero_z)
{
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int c = st->converged_at;
  if (c >= 0 && c < it) return;
  const bool first = it == 0, done = c == it;
  const T alpha = (T)st->alpha, beta = (T)st->beta;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride)
    {
      T dv[4], xv[4], zv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { dv[k] = first ? T(0) : d[i + k * stride]; xv[k] = first ? T(0) : x[i + k * stride]; zv[k] = done ? T(0) : z[i + k * stride]; }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        {
          if (!first) x[i + k * stride] = xv[k] + alpha * dv[k];
          if (!done) d[i + k * stride] = beta * dv[k] - zv[k];
          if (zero_z) z[i + k * stride] = T(0);
        }
    }
  for (; i < n; i += stride)
    {
      const T di = first ? T(0) : d[i];
      if (!first) x[i] += alpha * di;
      if (!done) d[i] = beta * di - z[i];
      if (zero_z) z[i] = T(0);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Building blocks of the CG loop over a partition (one process per GPU): the same three vector kernels with
// sums over the OWNED DoFs only; the scalars live in a block of 8 doubles on the device that the caller all-reduces in
// stream order between the kernels (NCCL), so the host never reads a scalar inside the loop.
//   scal[0] = d.h   scal[1] = g.g   scal[2] = g.z   scal[4] = g.z of the previous iterate   scal[5] = alpha
//   scal[6] = beta  scal[7] = iteration at which |g| <= tol was met (-1 before)   scal[3] = |g|   scal[9] = iteration counter
// (it < 0 in cgd_beta / cgd_advance: the iteration number is the device counter, so that one captured CUDA graph serves every iteration)
// ---------------------------------------------------------------------------------------------------------------
__device__ inline void block_partial_store(double a, double b, double *partial)
{
  block_sum2(a, b);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b; }
}
template <typename T>
__global__ void cgd_dot(const T *__restrict__ d, const T *__restrict__ h, const uint8_t *__restrict__ owned, size_t n, double *__restrict__ partial,
                        double *scal, unsigned *ticket)
{
  if (scal[7] >= 0) return;
  double s0 = 0, s1 = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride)  // four independent elements per thread in flight
    {
      T dv[4], hv[4];
      uint8_t ov[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { dv[k] = d[i + k * stride]; hv[k] = h[i + k * stride]; ov[k] = owned[i + k * stride]; }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ov[k]) s0 += (double)dv[k] * (double)hv[k];
    }
  for (; i < n; i += stride)
    if (owned[i]) s0 += (double)d[i] * (double)h[i];
  block_partial_store(s0, s1, partial);
  if (last_block_done(ticket))
    {
      sum_partials2(partial, s0, s1);
      if (threadIdx.x == 0) { scal[0] = s0; *ticket = 0; }
    }
}
// one thread: alpha = g.z / d.h (after the all-reduce of scal[0])
__global__ void cgd_alpha(double *scal) { if (scal[7] < 0) scal[5] = scal[4] / scal[0]; }
template <typename T>
__global__ void cgd_residual(T *__restrict__ g, T *__restrict__ h, const T *__restrict__ minv, const uint8_t *__restrict__ owned, size_t n,
                             double *__restrict__ partial, double *scal, unsigned *ticket, int first)
{
  if (scal[7] >= 0) return;
  const T alpha = first ? T(0) : (T)scal[5];
  double gg = 0, gz = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride)
    {
      T gv[4], hv[4], mv[4];
      uint8_t ov[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        {
          gv[k] = g[i + k * stride]; hv[k] = first ? T(0) : h[i + k * stride]; mv[k] = minv ? minv[i + k * stride] : T(1);
          ov[k] = owned[i + k * stride];
        }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        {
          const T gi = first ? gv[k] : gv[k] + alpha * hv[k];
          if (!first) g[i + k * stride] = gi;
          const T zi = mv[k] * gi;
          h[i + k * stride] = zi;
          if (ov[k]) { gg += (double)gi * (double)gi; gz += (double)gi * (double)zi; }
        }
    }
  for (; i < n; i += stride)
    {
      const T gi = first ? g[i] : g[i] + alpha * h[i];
      if (!first) g[i] = gi;
      const T zi = minv ? minv[i] * gi : gi;
      h[i] = zi;
      if (owned[i]) { gg += (double)gi * (double)gi; gz += (double)gi * (double)zi; }
    }
  block_partial_store(gg, gz, partial);
  if (last_block_done(ticket))
    {
      sum_partials2(partial, gg, gz);
      if (threadIdx.x == 0) { scal[1] = gg; scal[2] = gz; *ticket = 0; }
    }
}
// one thread: |g|, convergence, beta (after the all-reduce of scal[1..2])
__global__ void cgd_beta(double *scal, double tol, int it, double *history)
{
  if (scal[7] >= 0) return;
  if (it < 0) it = (int)scal[9] + 1;
  scal[9] = (double)it;
  const double res = sqrt(scal[1]);
  scal[3] = res;
  if (history) history[it] = res;
  scal[6] = it == 0 ? 0.0 : scal[2] / scal[4];
  scal[4] = scal[2];
  if (res <= tol) scal[7] = (double)it;
}
template <typename T> __global__ void cgd_advance(T *__restrict__ x, T *__restrict__ d, const T *__restrict__ z, size_t n, const double *scal, int it)
{
  if (it < 0) it = (int)scal[9];
  const double c = scal[7];
  if (c >= 0 && c < it) return;
  const bool first = it == 0, done = c == (double)it;
  const T alpha = (T)scal[5], beta = (T)scal[6];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride)
    {
      T dv[4], xv[4], zv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { dv[k] = first ? T(0) : d[i + k * stride]; xv[k] = first ? T(0) : x[i + k * stride]; zv[k] = done ? T(0) : z[i + k * stride]; }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        {
          if (!first) x[i + k * stride] = xv[k] + alpha * dv[k];
          if (!done) d[i + k * stride] = beta * dv[k] - zv[k];
        }
    }
  for (; i < n; i += stride)
    {
      const T di = first ? T(0) : d[i];
      if (!first) x[i] += alpha * di;
      if (!done) d[i] = beta * di - z[i];
    }
}

template <typename T>
void cg_solve(mfg_laplace *op, mfg_vec *x, const mfg_vec *b, double tol, int max_iter, bool jacobi, int *iters, double *last_res,
              double *history)
{
  mfg_ctx *ctx = op->ctx;
  const size_t n = op->mf->n_dofs;
  cudaStream_t s = ctx->stream;
  const int nb = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((RED_SCRATCH_DOUBLES - 8) / 2, (size_t)ctx->sm_count * 8), (n + TH * 8 - 1) / (TH * 8)));
  // work vectors live in the operator and are reused by later solves: cudaMalloc / cudaFree of vectors this size cost
  // tens of milliseconds per solve once the process holds large pinned host buffers (measured)
  const size_t wbytes = ((3 * n * sizeof(T) + 63) / 64) * 64 + 256;
  if (op->solver_work.n != wbytes) op->solver_work.alloc(wbytes);
  struct View { T *p; } g{(T *)op->solver_work.p}, d{g.p + n}, h{d.p + n};
  struct SView { CgState *p; } st{reinterpret_cast<CgState *>(op->solver_work.p + ((3 * n * sizeof(T) + 63) / 64) * 64)};
  DevBuf<double> hist(history ? (size_t)max_iter + 1 : 0);
  const T *minv = nullptr;
  if (jacobi)
    {
      if (!op->diagonal_is_available) laplace_compute_diagonal(op);
      minv = (const T *)op->inv_diag->p;
    }
  mfg_vec vg;
  vg.ctx = ctx; vg.dt = x->dt; vg.n = n; vg.owns = false; vg.p = g.p;
  // g = A x - b   (g = -b if x is zero)
  if (vec_all_zero(x)) vec_equ(&vg, -1.0, b);
  else { laplace_vmult(op, g.p, x->p, false); vec_sadd(&vg, 1.0, -1.0, b); }
  CgState init;
  init.gh = 0; init.alpha = 0; init.beta = 0; init.res = 0; init.tol = tol; init.converged_at = -1; init.ticket = 0;
  MFG_CUDA(cudaMemcpyAsync(st.p, &init, sizeof(init), cudaMemcpyHostToDevice, s));
  double *partial = ctx->red_dev + 8;
  // pinned mirror of the state, refreshed after every iteration; the host looks at the copy of iteration it - LAG
  // (already complete in practice) and stops enqueueing once it shows convergence
  constexpr int LAG = 3;
  static_assert((LAG + 1) * sizeof(CgState) <= RED_HOST_DOUBLES * sizeof(double), "pinned scratch too small");
  CgState *mirror = reinterpret_cast<CgState *>(ctx->red_host);
  cudaEvent_t ev[LAG + 1];
  for (auto &e : ev) MFG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto snapshot = [&](int it) {
    MFG_CUDA(cudaMemcpyAsync(mirror + it % (LAG + 1), st.p, sizeof(CgState), cudaMemcpyDeviceToHost, s));
    MFG_CUDA(cudaEventRecord(ev[it % (LAG + 1)], s));
  };
  // fused loop: the cell kernel emits d . (A d) and runs behind cg_advance, which zeroes its destination
  uint32_t n_dot = 0;
  if (op->solver_dot.n < 8192) op->solver_dot.alloc(8192);
  MFG_CUDA(cudaMemsetAsync(op->solver_dot.p, 0, op->solver_dot.bytes(), s));
  const bool fused = op->cg_fused && laplace_active_variant(op) == 50 && op->mf->hn_mask.n == 0 && op->mf->n_cells > 0 &&
                     (size_t)ctx->sm_count * 3 * 4 <= op->solver_dot.n;
  // iteration 0: h = Minv g, |g|, g.h ; d = -h
  cg_residual<T><<<nb, TH, 0, s>>>(g.p, h.p, minv, n, partial, st.p, 0, hist.p, nullptr, 0u);
  MFG_CUDA_LAST();
  cg_advance<T><<<nb, TH, 0, s>>>((T *)x->p, d.p, h.p, n, st.p, 0, fused);
  MFG_CUDA_LAST();
  snapshot(0);
  for (int it = 1; it <= max_iter; ++it)
    {
      if (it > LAG)
        {
          MFG_CUDA(cudaEventSynchronize(ev[(it - LAG) % (LAG + 1)]));
          if (mirror[(it - LAG) % (LAG + 1)].converged_at >= 0) break;
        }
      bool with_dot = false;
      if (fused) with_dot = laplace_cell_dot(op, h.p, d.p, op->solver_dot.p, &n_dot);   // h = A d and the partial sums of d . h
      if (!with_dot) laplace_vmult(op, h.p, d.p, false);                                  // h = A d
      if (!with_dot) { cg_dot<T><<<nb, TH, 0, s>>>(d.p, h.p, n, partial, st.p); MFG_CUDA_LAST(); }  // alpha
      cg_residual<T><<<nb, TH, 0, s>>>(g.p, h.p, minv, n, partial, st.p, it, hist.p, with_dot ? op->solver_dot.p : nullptr, n_dot);
      MFG_CUDA_LAST();
      cg_advance<T><<<nb, TH, 0, s>>>((T *)x->p, d.p, h.p, n, st.p, it, fused);
      MFG_CUDA_LAST();
      snapshot(it);
    }
  CgState fin;
  MFG_CUDA(cudaMemcpyAsync(&fin, st.p, sizeof(fin), cudaMemcpyDeviceToHost, s));
  MFG_CUDA(cudaStreamSynchronize(s));
  const int it = fin.converged_at >= 0 ? fin.converged_at : max_iter;
  if (history) MFG_CUDA(cudaMemcpy(history, hist.p, ((size_t)it + 1) * sizeof(double), cudaMemcpyDeviceToHost));
  for (auto &e : ev) cudaEventDestroy(e);
  if (iters) *iters = it;
  if (last_res) *last_res = fin.res;
}

}  // namespace
}  // namespace mfg

using namespace mfg;

extern "C" int mfg_solver_cg(mfg_laplace *op, mfg_vec *x, const mfg_vec *b, double abs_tol, int max_iter, int use_jacobi, int *iters,
                             double *last_residual, double *residual_history)
{
  return guarded([&] {
    MFG_REQUIRE(op && x && b, "null argument");
    MFG_REQUIRE(x->dt == op->mf->dt && b->dt == op->mf->dt, "vector dtype differs from operator dtype");
    MFG_REQUIRE(x->n == op->mf->n_dofs && b->n == op->mf->n_dofs, "vector size differs from operator size");
    MFG_REQUIRE(max_iter >= 0, "max_iter must be non-negative");
    if (op->mf->dt == MFG_F64) cg_solve<double>(op, x, b, abs_tol, max_iter, use_jacobi != 0, iters, last_residual, residual_history);
    else cg_solve<float>(op, x, b, abs_tol, max_iter, use_jacobi != 0, iters, last_residual, residual_history);
  });
}

// ---- CG over a partition: kernels on device pointers; `scal` = 8 doubles on the device (+ 1 unsigned ticket behind them) ----
static int cgd_blocks(const mfg_ctx *ctx, size_t n)
{
  return (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((RED_SCRATCH_DOUBLES - 8) / 2, (size_t)ctx->sm_count * 8), (n + 256 * 8 - 1) / (256 * 8)));
}
extern "C" int mfg_cgd_init(mfg_ctx *ctx, double *scal_dev)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && scal_dev, "null argument");
    const double init[10] = {0, 0, 0, 0, 0, 0, 0, -1.0, 0, 0};  // ([8] holds the ticket counter: zero bits)
    MFG_CUDA(cudaMemcpyAsync(scal_dev, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    MFG_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}
extern "C" int mfg_cgd_dot(mfg_ctx *ctx, mfg_dtype dt, const void *d, const void *h, const uint8_t *owned_dev, size_t n, double *scal_dev)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && d && h && owned_dev && scal_dev, "null argument");
    unsigned *ticket = reinterpret_cast<unsigned *>(scal_dev + 8);
    const int nb = cgd_blocks(ctx, n);
    if (dt == MFG_F64) cgd_dot<double><<<nb, 256, 0, ctx->stream>>>((const double *)d, (const double *)h, owned_dev, n, ctx->red_dev + 8, scal_dev, ticket);
    else cgd_dot<float><<<nb, 256, 0, ctx->stream>>>((const float *)d, (const float *)h, owned_dev, n, ctx->red_dev + 8, scal_dev, ticket);
    MFG_CUDA_LAST();
  });
}
extern "C" int mfg_cgd_alpha(mfg_ctx *ctx, double *scal_dev)
{
  return guarded([&] { MFG_REQUIRE(ctx && scal_dev, "null argument"); cgd_alpha<<<1, 1, 0, ctx->stream>>>(scal_dev); MFG_CUDA_LAST(); });
}
extern "C" int mfg_cgd_residual(mfg_ctx *ctx, mfg_dtype dt, void *g, void *h, const void *minv, const uint8_t *owned_dev, size_t n, double *scal_dev, int first)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && g && h && owned_dev && scal_dev, "null argument");
    unsigned *ticket = reinterpret_cast<unsigned *>(scal_dev + 8);
    const int nb = cgd_blocks(ctx, n);
    if (dt == MFG_F64) cgd_residual<double><<<nb, 256, 0, ctx->stream>>>((double *)g, (double *)h, (const double *)minv, owned_dev, n, ctx->red_dev + 8, scal_dev, ticket, first);
    else cgd_residual<float><<<nb, 256, 0, ctx->stream>>>((float *)g, (float *)h, (const float *)minv, owned_dev, n, ctx->red_dev + 8, scal_dev, ticket, first);
    MFG_CUDA_LAST();
  });
}
extern "C" int mfg_cgd_beta(mfg_ctx *ctx, double *scal_dev, double tol, int it)
{
  return guarded([&] { MFG_REQUIRE(ctx && scal_dev, "null argument"); cgd_beta<<<1, 1, 0, ctx->stream>>>(scal_dev, tol, it, nullptr); MFG_CUDA_LAST(); });
}
extern "C" int mfg_cgd_advance(mfg_ctx *ctx, mfg_dtype dt, void *x, void *d, const void *z, size_t n, const double *scal_dev, int it)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && x && d && z && scal_dev, "null argument");
    const int nb = cgd_blocks(ctx, n);
    if (dt == MFG_F64) cgd_advance<double><<<nb, 256, 0, ctx->stream>>>((double *)x, (double *)d, (const double *)z, n, scal_dev, it);
    else cgd_advance<float><<<nb, 256, 0, ctx->stream>>>((float *)x, (float *)d, (const float *)z, n, scal_dev, it);
    MFG_CUDA_LAST();
  });
}
