// vector.cu -- GpuVector<Number> storage and BLAS-1 (gpu_vec.h:22-176,
// gpu_vec.cu:221-644), type-erased over float/double behind the C ABI.
//
// Differences from the reference: grid-stride kernels sized to the SM count
// instead of one block per 4096 entries with `int` indexing (gpu_vec.cu:311);
// reductions are two-pass with a fixed block order (deterministic, the
// reference's one atomicAdd per block is not, gpu_vec.cu:430-433) and reuse a
// context-owned scratch buffer + pinned result instead of
// cudaMalloc/cudaMemset/cudaMemcpy/cudaFree per call (gpu_vec.cu:543-557).
#include "vector.cuh"

namespace mfg {

namespace {

constexpr int EW_THREADS = 256;
inline unsigned ew_blocks(const mfg_ctx *ctx, size_t n)
{
  const size_t want = (n + EW_THREADS * 4 - 1) / (EW_THREADS * 4);
  const size_t cap  = (size_t)ctx->sm_count * 16;
  return (unsigned)std::max<size_t>(1, std::min(want, cap));
}

template <typename T, typename F> __global__ void ew1(T *v, size_t n, F f)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = f(v[i]);
}
template <typename T, typename S, typename F> __global__ void ew2(T *v, const S *x, size_t n, F f)
{
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) v[i] = f(v[i], x[i]);
}

template <typename T> struct FillOp { T a; __device__ T operator()(T) const { return a; } };
template <typename T> struct ScalOp { T a; __device__ T operator()(T v) const { return a * v; } };
template <typename T> struct InvOp { __device__ T operator()(T v) const { return T(1) / v; } };
template <typename T> struct SaddOp { T s, a; __device__ T operator()(T v, T x) const { return s * v + a * x; } };
template <typename T, typename S> struct EquOp { T a; __device__ T operator()(T, S x) const { return a * (T)x; } };
template <typename T> struct MulOp { __device__ T operator()(T v, T x) const { return v * x; } };
template <typename T> struct DivOp { __device__ T operator()(T v, T x) const { return v / x; } };

// ---- reductions -------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1024;

__device__ inline double block_sum(double x)
{
  __shared__ double sh[RED_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32)
    {
      x = threadIdx.x < RED_THREADS / 32 ? sh[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    }
  return x;  // valid in thread 0
}

// mode 0: a.b   mode 1: v += alpha*x, then v.w   mode 2: count of non-zeros
template <typename T, int MODE>
__global__ void red_pass1(T *v, const T *x, const T *w, T alpha, size_t n, double *partial)
{
  double       acc    = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      if (MODE == 0) acc += (double)v[i] * (double)x[i];
      else if (MODE == 1) { const T nv = v[i] + alpha * x[i]; v[i] = nv; acc += (double)nv * (double)w[i]; }
      else acc += (v[i] != T(0)) ? 1.0 : 0.0;
    }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void red_pass2(const double *partial, int nb, double *out)
{
  double acc = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  acc = block_sum(acc);
  if (threadIdx.x == 0) *out = acc;
}

template <typename T, int MODE> double reduce(mfg_ctx *ctx, T *v, const T *x, const T *w, T alpha, size_t n)
{
  const int nb = (int)std::max<size_t>(1, std::min<size_t>(RED_MAX_BLOCKS, (n + RED_THREADS * 8 - 1) / (RED_THREADS * 8)));
  red_pass1<T, MODE><<<nb, RED_THREADS, 0, ctx->stream>>>(v, x, w, alpha, n, ctx->red_dev + 8);
  MFG_CUDA_LAST();
  red_pass2<<<1, RED_THREADS, 0, ctx->stream>>>(ctx->red_dev + 8, nb, ctx->red_dev);
  MFG_CUDA_LAST();
  MFG_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  MFG_CUDA(cudaStreamSynchronize(ctx->stream));
  return ctx->red_host[0];
}

template <typename D, typename S, typename I>
__global__ void copy_idx(D *dst, const S *src, const I *di, const I *si, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[di[i]] = (D)src[si[i]];
}

}  // namespace

#define DISPATCH(v, CALL)                                  \
  do {                                                     \
    if ((v)->dt == MFG_F64) { typedef double T; CALL; }    \
    else { typedef float T; CALL; }                        \
  } while (0)

static void same(const mfg_vec *a, const mfg_vec *b)
{
  MFG_REQUIRE(a->dt == b->dt, "vector dtypes differ");
  MFG_REQUIRE(a->n == b->n, "vector sizes differ");
}

void vec_fill(mfg_vec *v, double a)
{
  if (!v->n) return;
  DISPATCH(v, (ew1<T><<<ew_blocks(v->ctx, v->n), EW_THREADS, 0, v->ctx->stream>>>((T *)v->p, v->n, FillOp<T>{(T)a})));
  MFG_CUDA_LAST();
}
void vec_scal(mfg_vec *v, double a)
{
  if (!v->n) return;
  DISPATCH(v, (ew1<T><<<ew_blocks(v->ctx, v->n), EW_THREADS, 0, v->ctx->stream>>>((T *)v->p, v->n, ScalOp<T>{(T)a})));
  MFG_CUDA_LAST();
}
void vec_invert(mfg_vec *v)
{
  if (!v->n) return;
  DISPATCH(v, (ew1<T><<<ew_blocks(v->ctx, v->n), EW_THREADS, 0, v->ctx->stream>>>((T *)v->p, v->n, InvOp<T>{})));
  MFG_CUDA_LAST();
}
void vec_sadd(mfg_vec *v, double s, double a, const mfg_vec *x)
{
  same(v, x); if (!v->n) return;
  DISPATCH(v, (ew2<T, T><<<ew_blocks(v->ctx, v->n), EW_THREADS, 0, v->ctx->stream>>>((T *)v->p, (const T *)x->p, v->n, SaddOp<T>{(T)s, (T)a})));
  MFG_CUDA_LAST();
}
void vec_equ(mfg_vec *v, double a, const mfg_vec *x)
{
  MFG_REQUIRE(v->n == x->n, "vector sizes differ"); if (!v->n) return;
  const unsigned nb = ew_blocks(v->ctx, v->n); cudaStream_t s = v->ctx->stream;
  if (v->dt == MFG_F64 && x->dt == MFG_F64) ew2<double, double><<<nb, EW_THREADS, 0, s>>>((double *)v->p, (const double *)x->p, v->n, EquOp<double, double>{a});
  else if (v->dt == MFG_F64) ew2<double, float><<<nb, EW_THREADS, 0, s>>>((double *)v->p, (const float *)x->p, v->n, EquOp<double, float>{a});
  else if (x->dt == MFG_F64) ew2<float, double><<<nb, EW_THREADS, 0, s>>>((float *)v->p, (const double *)x->p, v->n, EquOp<float, double>{(float)a});
  else ew2<float, float><<<nb, EW_THREADS, 0, s>>>((float *)v->p, (const float *)x->p, v->n, EquOp<float, float>{(float)a});
  MFG_CUDA_LAST();
}
void vec_scale(mfg_vec *v, const mfg_vec *x)
{
  same(v, x); if (!v->n) return;
  DISPATCH(v, (ew2<T, T><<<ew_blocks(v->ctx, v->n), EW_THREADS, 0, v->ctx->stream>>>((T *)v->p, (const T *)x->p, v->n, MulOp<T>{})));
  MFG_CUDA_LAST();
}
void vec_divide(mfg_vec *v, const mfg_vec *x)
{
  same(v, x); if (!v->n) return;
  DISPATCH(v, (ew2<T, T><<<ew_blocks(v->ctx, v->n), EW_THREADS, 0, v->ctx->stream>>>((T *)v->p, (const T *)x->p, v->n, DivOp<T>{})));
  MFG_CUDA_LAST();
}
double vec_dot(const mfg_vec *a, const mfg_vec *b)
{
  same(a, b); if (!a->n) return 0.0;
  double r = 0;
  DISPATCH(a, (r = reduce<T, 0>(a->ctx, (T *)a->p, (const T *)b->p, (const T *)nullptr, T(0), a->n)));
  return r;
}
double vec_add_and_dot(mfg_vec *v, double alpha, const mfg_vec *x, const mfg_vec *w)
{
  same(v, x); same(v, w); if (!v->n) return 0.0;
  double r = 0;
  DISPATCH(v, (r = reduce<T, 1>(v->ctx, (T *)v->p, (const T *)x->p, (const T *)w->p, (T)alpha, v->n)));
  return r;
}
bool vec_all_zero(const mfg_vec *v)
{
  if (!v->n) return true;
  double r = 0;
  DISPATCH(v, (r = reduce<T, 2>(v->ctx, (T *)v->p, (const T *)nullptr, (const T *)nullptr, T(0), v->n)));
  return r == 0.0;
}
void vec_copy_with_indices(mfg_vec *dst, const mfg_vec *src, const uint32_t *di, const uint32_t *si, size_t n)
{
  if (!n) return;
  const unsigned nb = (unsigned)((n + 255) / 256); cudaStream_t s = dst->ctx->stream;
  if (dst->dt == MFG_F64 && src->dt == MFG_F64) copy_idx<<<nb, 256, 0, s>>>((double *)dst->p, (const double *)src->p, di, si, n);
  else if (dst->dt == MFG_F64) copy_idx<<<nb, 256, 0, s>>>((double *)dst->p, (const float *)src->p, di, si, n);
  else if (src->dt == MFG_F64) copy_idx<<<nb, 256, 0, s>>>((float *)dst->p, (const double *)src->p, di, si, n);
  else copy_idx<<<nb, 256, 0, s>>>((float *)dst->p, (const float *)src->p, di, si, n);
  MFG_CUDA_LAST();
}

}  // namespace mfg
