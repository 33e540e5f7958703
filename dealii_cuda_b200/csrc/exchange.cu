// exchange.cu -- interface-DoF exchange kernels for the multi-GPU partition (the reference has no
// multi-GPU code).  pack gathers this rank's partial sums of the interface DoFs into the send buffer;
// accumulate adds the contributions of all sharers in ascending rank order so every replica ends up with the
// bit-identical sum.
#include "operators.cuh"

struct mfg_exchange
{
  mfg_ctx  *ctx = nullptr;
  mfg_dtype dt = MFG_F64;
  mfg::DevBuf<uint32_t> pack_idx, shared_dofs, offsets;
  mfg::DevBuf<int32_t>  slots;
};

namespace mfg {
namespace {
template <typename T> __global__ void k_pack(const T *__restrict__ v, const uint32_t *__restrict__ pidx, size_t n, T *__restrict__ send)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) send[i] = v[pidx[i]];
}
// pack + send in one step over peer memory: entry i of this rank's send order belongs to chunk c (one chunk per
// neighbour, ascending rank) and is stored straight into that neighbour's receive buffer through NVLink
struct PushTable { unsigned long long dst[32]; uint32_t start[33]; int n; };
template <typename T> __global__ void k_push(const T *__restrict__ v, const uint32_t *__restrict__ pidx, size_t n, const PushTable tb)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  while (c + 1 < tb.n && i >= tb.start[c + 1]) ++c;
  reinterpret_cast<T *>(tb.dst[c])[i - tb.start[c]] = v[pidx[i]];
}
template <typename T>
__global__ void k_accumulate(T *__restrict__ v, const T *__restrict__ recv, const uint32_t *__restrict__ dofs,
                             const uint32_t *__restrict__ offsets, const int32_t *__restrict__ slots, size_t n)
{
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n) return;
  const uint32_t d = dofs[u];
  const T mine = v[d];
  T acc = 0;
  for (uint32_t j = offsets[u]; j < offsets[u + 1]; ++j)
    {
      const int32_t s = slots[j];
      acc += s < 0 ? mine : recv[s];
    }
  v[d] = acc;
}
template <typename T>
__global__ void k_dot_masked(const T *__restrict__ a, const T *__restrict__ b, const uint8_t *__restrict__ m, size_t n, double *partial)
{
  __shared__ double sh[8];
  double acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    if (m[i]) acc += (double)a[i] * (double)b[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32)
    {
      acc = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (threadIdx.x == 0) partial[blockIdx.x] = acc;
    }
}
__global__ void k_sum_partials(const double *partial, int nb, double *out)
{
  __shared__ double sh[8];
  double acc = 0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) acc += partial[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32)
    {
      acc = threadIdx.x < 8 ? sh[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (threadIdx.x == 0) *out = acc;
    }
}
}  // namespace
}  // namespace mfg

using namespace mfg;

extern "C" {

int mfg_exchange_create(mfg_ctx *ctx, mfg_dtype dt, const uint32_t *pack_idx_host, size_t n_send, const uint32_t *shared_dofs_host,
                        size_t n_shared, const uint32_t *offsets_host, const int32_t *slots_host, size_t n_slots, mfg_exchange **out)
{
  return guarded([&] {
    MFG_REQUIRE(ctx && out, "null argument");
    MFG_REQUIRE((pack_idx_host || !n_send) && (shared_dofs_host || !n_shared) && (slots_host || !n_slots), "null array");
    MFG_REQUIRE(offsets_host != nullptr, "offsets required (n_shared + 1 entries)");
    MFG_REQUIRE(offsets_host[0] == 0 && offsets_host[n_shared] == n_slots, "offsets do not span the slot array");
    std::unique_ptr<mfg_exchange> ex(new mfg_exchange);
    ex->ctx = ctx; ex->dt = dt;
    ex->pack_idx.upload(pack_idx_host, n_send, ctx->stream);
    ex->shared_dofs.upload(shared_dofs_host, n_shared, ctx->stream);
    ex->offsets.upload(offsets_host, n_shared + 1, ctx->stream);
    ex->slots.upload(slots_host, n_slots, ctx->stream);
    *out = ex.release();
  });
}
int mfg_exchange_destroy(mfg_exchange *ex) { return guarded([&] { delete ex; }); }
// tb = threads per block.  The overlapped form runs next to the persistent interior cell kernel, which leaves about
// 1000 registers per SM free: only one-warp blocks with few registers find a slot there (pack and accumulate use < 24)
static void pack_on(mfg_exchange *ex, const void *vec_dev, void *send_dev, cudaStream_t st, unsigned tb = 256)
{
  MFG_REQUIRE(ex && vec_dev && (send_dev || !ex->pack_idx.n), "null argument");
  const size_t n = ex->pack_idx.n; if (!n) return;
  const unsigned nb = (unsigned)((n + tb - 1) / tb);
  if (ex->dt == MFG_F64) k_pack<double><<<nb, tb, 0, st>>>((const double *)vec_dev, ex->pack_idx.p, n, (double *)send_dev);
  else k_pack<float><<<nb, tb, 0, st>>>((const float *)vec_dev, ex->pack_idx.p, n, (float *)send_dev);
  MFG_CUDA_LAST();
}
int mfg_exchange_pack(mfg_exchange *ex, const void *vec_dev, void *send_dev)
{
  return guarded([&] { MFG_REQUIRE(ex, "null argument"); pack_on(ex, vec_dev, send_dev, ex->ctx->stream); });
}
int mfg_exchange_pack_stream(mfg_exchange *ex, const void *vec_dev, void *send_dev, void *cuda_stream)
{
  return guarded([&] { pack_on(ex, vec_dev, send_dev, (cudaStream_t)cuda_stream, 32); });
}
static void accumulate_on(mfg_exchange *ex, void *vec_dev, const void *recv_dev, cudaStream_t st, unsigned tb = 256)
{
  MFG_REQUIRE(ex && vec_dev, "null argument");
  const size_t n = ex->shared_dofs.n; if (!n) return;
  const unsigned nb = (unsigned)((n + tb - 1) / tb);
  if (ex->dt == MFG_F64)
    k_accumulate<double><<<nb, tb, 0, st>>>((double *)vec_dev, (const double *)recv_dev, ex->shared_dofs.p, ex->offsets.p, ex->slots.p, n);
  else
    k_accumulate<float><<<nb, tb, 0, st>>>((float *)vec_dev, (const float *)recv_dev, ex->shared_dofs.p, ex->offsets.p, ex->slots.p, n);
  MFG_CUDA_LAST();
}
int mfg_exchange_accumulate(mfg_exchange *ex, void *vec_dev, const void *recv_dev)
{
  return guarded([&] { MFG_REQUIRE(ex, "null argument"); accumulate_on(ex, vec_dev, recv_dev, ex->ctx->stream); });
}
int mfg_exchange_accumulate_stream(mfg_exchange *ex, void *vec_dev, const void *recv_dev, void *cuda_stream)
{
  return guarded([&] { accumulate_on(ex, vec_dev, recv_dev, (cudaStream_t)cuda_stream, 32); });
}
int mfg_exchange_push_stream(mfg_exchange *ex, const void *vec_dev, const uint64_t *peer_dst, const uint32_t *chunk_start, int n_chunks,
                             void *cuda_stream)
{
  return guarded([&] {
    MFG_REQUIRE(ex && vec_dev && peer_dst && chunk_start, "null argument");
    MFG_REQUIRE(n_chunks >= 1 && n_chunks <= 32, "between 1 and 32 neighbours");
    const size_t n = ex->pack_idx.n; if (!n) return;
    MFG_REQUIRE(chunk_start[0] == 0 && chunk_start[n_chunks] == n, "chunk starts do not span the send list");
    PushTable tb;
    tb.n = n_chunks;
    for (int c = 0; c < n_chunks; ++c) { tb.dst[c] = peer_dst[c]; tb.start[c] = chunk_start[c]; }
    tb.start[n_chunks] = chunk_start[n_chunks];
    const unsigned tbk = 32, nb = (unsigned)((n + tbk - 1) / tbk);  // one-warp blocks: they run beside the persistent cell kernel
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ex->ctx->stream;
    if (ex->dt == MFG_F64) k_push<double><<<nb, tbk, 0, st>>>((const double *)vec_dev, ex->pack_idx.p, n, tb);
    else k_push<float><<<nb, tbk, 0, st>>>((const float *)vec_dev, ex->pack_idx.p, n, tb);
    MFG_CUDA_LAST();
  });
}
int mfg_vec_dot_masked(const mfg_vec *a, const mfg_vec *b, const uint8_t *owned_mask_dev, double *out)
{
  return guarded([&] {
    MFG_REQUIRE(a && b && owned_mask_dev && out, "null argument");
    MFG_REQUIRE(a->dt == b->dt && a->n == b->n, "vectors differ in dtype or size");
    mfg_ctx *ctx = a->ctx;
    if (!a->n) { *out = 0; return; }
    const int nb = (int)std::max<size_t>(1, std::min<size_t>(1024, (a->n + 2047) / 2048));
    if (a->dt == MFG_F64) k_dot_masked<double><<<nb, 256, 0, ctx->stream>>>((const double *)a->p, (const double *)b->p, owned_mask_dev, a->n, ctx->red_dev + 8);
    else k_dot_masked<float><<<nb, 256, 0, ctx->stream>>>((const float *)a->p, (const float *)b->p, owned_mask_dev, a->n, ctx->red_dev + 8);
    MFG_CUDA_LAST();
    k_sum_partials<<<1, 256, 0, ctx->stream>>>(ctx->red_dev + 8, nb, ctx->red_dev);
    MFG_CUDA_LAST();
    MFG_CUDA(cudaMemcpyAsync(ctx->red_host, ctx->red_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    MFG_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = ctx->red_host[0];
  });
}

}  // extern "C"
