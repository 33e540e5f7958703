// solver.cu -- conjugate gradients on GpuVectors with the Laplace operator: the control flow of deal.II's
// SolverCG<GpuVector> as the reference instantiates it (poisson.cu:233-260), with the
// BLAS-1 of every iteration fused into two reduction kernels instead of the reference's five
// (operator*, add, add_and_dot, DiagonalMatrix::vmult -> scale, sadd; gpu_vec.cu:306-617), each of which
// cudaMallocs and blocks on a D2H copy there.  Preconditioner: the inverse diagonal (PreconditionChebyshev with
// its default degree 0 is a scaled Jacobi step; the scaling does not change the CG iterates).
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

// x += alpha d ; g += alpha h ; z = Minv .* g (or z = g) ; partial[2b] = g.g, partial[2b+1] = g.z
template <typename T>
__global__ void cg_update(T *__restrict__ x, T *__restrict__ g, T *__restrict__ z, const T *__restrict__ d, const T *__restrict__ h,
                          const T *__restrict__ minv, T alpha, size_t n, double *__restrict__ partial)
{
  double gg = 0, gz = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      x[i] += alpha * d[i];
      const T gi = g[i] + alpha * h[i];
      g[i] = gi;
      const T zi = minv ? minv[i] * gi : gi;
      z[i] = zi;
      gg += (double)gi * (double)gi;
      gz += (double)gi * (double)zi;
    }
  block_sum2(gg, gz);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = gg; partial[2 * blockIdx.x + 1] = gz; }
}
// d = beta d - z ; (fused with nothing else: the next operation is the operator apply)
template <typename T> __global__ void cg_direction(T *__restrict__ d, const T *__restrict__ z, T beta, size_t n)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = beta * d[i] - z[i];
}
// partial[2b] = a.b, partial[2b+1] = c.c
template <typename T>
__global__ void dot2(const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ c, size_t n, double *__restrict__ partial)
{
  double s0 = 0, s1 = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      s0 += (double)a[i] * (double)b[i];
      if (c) s1 += (double)c[i] * (double)c[i];
    }
  block_sum2(s0, s1);
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s0; partial[2 * blockIdx.x + 1] = s1; }
}
__global__ void finish2(const double *partial, int nb, double *out)
{
  double a = 0, b = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
  block_sum2(a, b);
  if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

template <typename T> struct Cg
{
  mfg_laplace *op; mfg_ctx *ctx; size_t n; int nb;
  void reduce2(double &a, double &b)
  {
    finish2<<<1, TH, 0, ctx->stream>>>(ctx->red_dev + 8, nb, ctx->red_dev);
    MFG_CUDA_LAST();
    MFG_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MFG_CUDA(cudaStreamSynchronize(ctx->stream));
    a = ctx->red_host[0]; b = ctx->red_host[1];
  }
};

template <typename T>
void cg_solve(mfg_laplace *op, mfg_vec *x, const mfg_vec *b, double tol, int max_iter, bool jacobi, int *iters, double *last_res,
              double *history)
{
  mfg_ctx *ctx = op->ctx;
  const size_t n = op->mf->n_dofs;
  cudaStream_t s = ctx->stream;
  Cg<T> cg{op, ctx, n, (int)std::max<size_t>(1, std::min<size_t>((RED_SCRATCH_DOUBLES - 8) / 2, (n + TH * 8 - 1) / (TH * 8)))};
  const int nb = cg.nb;
  DevBuf<T> g(n), d(n), h(n);
  MFG_CUDA(cudaMemsetAsync(d.p, 0, n * sizeof(T), s));
  MFG_CUDA(cudaMemsetAsync(h.p, 0, n * sizeof(T), s));
  const T *minv = nullptr;
  if (jacobi)
    {
      if (!op->diagonal_is_available) laplace_compute_diagonal(op);
      minv = (const T *)op->inv_diag->p;
    }
  mfg_vec vg, vd, vh;
  vg.ctx = vd.ctx = vh.ctx = ctx; vg.dt = vd.dt = vh.dt = x->dt; vg.n = vd.n = vh.n = n; vg.owns = vd.owns = vh.owns = false;
  vg.p = g.p; vd.p = d.p; vh.p = h.p;
  // g = A x - b   (g = -b if x is zero)
  if (vec_all_zero(x)) vec_equ(&vg, -1.0, b);
  else { laplace_vmult(op, g.p, x->p, false); vec_sadd(&vg, 1.0, -1.0, b); }
  // h = Minv g (or g) ; d = -h ; gh = g.h ; res = |g|
  double gg, gh;
  cg_update<T><<<nb, TH, 0, s>>>((T *)x->p, g.p, h.p, d.p, h.p, minv, T(0), n, ctx->red_dev + 8);  // alpha = 0: only z and the sums
  MFG_CUDA_LAST();
  cg.reduce2(gg, gh);
  double res = std::sqrt(gg);
  int it = 0;
  if (history) history[0] = res;
  if (res > tol)
    {
      vec_equ(&vd, -1.0, &vh);
      for (it = 1; it <= max_iter; ++it)
        {
          laplace_vmult(op, h.p, d.p, false);                               // h = A d
          dot2<T><<<nb, TH, 0, s>>>(d.p, h.p, (const T *)nullptr, n, ctx->red_dev + 8);
          MFG_CUDA_LAST();
          double dh, dummy;
          cg.reduce2(dh, dummy);
          const double alpha = gh / dh;
          // x += alpha d ; g += alpha h ; h = Minv g ; |g|^2 ; g.h
          cg_update<T><<<nb, TH, 0, s>>>((T *)x->p, g.p, h.p, d.p, h.p, minv, (T)alpha, n, ctx->red_dev + 8);
          MFG_CUDA_LAST();
          double gh_new;
          cg.reduce2(gg, gh_new);
          res = std::sqrt(gg);
          if (history) history[it] = res;
          if (res <= tol) break;
          const double beta = gh_new / gh;
          gh = gh_new;
          cg_direction<T><<<nb, TH, 0, s>>>(d.p, h.p, (T)beta, n);          // d = beta d - h
          MFG_CUDA_LAST();
        }
      if (it > max_iter) it = max_iter;
    }
  MFG_CUDA(cudaStreamSynchronize(s));
  if (iters) *iters = it;
  if (last_res) *last_res = res;
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
