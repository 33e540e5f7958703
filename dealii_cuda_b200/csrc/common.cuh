// common.cuh -- error handling, context and small device-array helpers shared by
// all translation units of libmfgpu.so.
// Replaces matrix_free_gpu/cuda_utils.cuh:15-27 (CUDA_CHECK_SUCCESS / CUDA_CHECK_LAST),
// gpu_list.{h,cu} (immutable device index array) and utils.h (ipow).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/mfgpu.h"

namespace mfg {

struct Error : std::runtime_error
{
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &msg);

#define MFG_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      throw ::mfg::Error(MFG_ERR_CUDA, std::string("CUDA error '") + cudaGetErrorString(e__) + \
                                           "' at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)
#define MFG_CUDA_LAST() MFG_CUDA(cudaGetLastError())
#define MFG_REQUIRE(cond, msg)                                                              \
  do {                                                                                      \
    if (!(cond)) throw ::mfg::Error(MFG_ERR_INVALID, std::string(msg) + " (" #cond ")");    \
  } while (0)

// wraps the body of every extern "C" entry point
template <typename F>
int guarded(F &&f) noexcept
{
  try { f(); return MFG_OK; }
  catch (const Error &e) { set_last_error(e.what()); return e.code; }
  catch (const std::bad_alloc &) { set_last_error("out of host memory"); return MFG_ERR_NOMEM; }
  catch (const std::exception &e) { set_last_error(e.what()); return MFG_ERR_INVALID; }
  catch (...) { set_last_error("unknown error"); return MFG_ERR_INVALID; }
}

constexpr __host__ __device__ inline unsigned ipow(unsigned b, int e) { return e <= 0 ? 1u : b * ipow(b, e - 1); }

template <typename T> struct DevBuf  // RAII device array (GpuList<T> / raw cudaMalloc in the reference)
{
  T     *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t n_)
  {
    release();
    n = n_;
    if (n) { cudaError_t e = cudaMalloc(&p, n * sizeof(T)); if (e != cudaSuccess) { p = nullptr; n = 0; throw Error(MFG_ERR_NOMEM, std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); } }
  }
  void upload(const T *host, size_t n_, cudaStream_t s)
  {
    if (n != n_) alloc(n_);
    if (n) { MFG_CUDA(cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, s)); MFG_CUDA(cudaStreamSynchronize(s)); }
  }
  void download(T *host, cudaStream_t s) const
  {
    if (n) { MFG_CUDA(cudaMemcpyAsync(host, p, n * sizeof(T), cudaMemcpyDeviceToHost, s)); MFG_CUDA(cudaStreamSynchronize(s)); }
  }
  size_t bytes() const { return n * sizeof(T); }
};

}  // namespace mfg

struct mfg_ctx
{
  int          device   = 0;
  cudaStream_t stream   = nullptr;
  int          sm_count = 0, cc_major = 0, cc_minor = 0;
  size_t       l2_bytes = 0;
  // device scratch for reductions (the reference cudaMallocs per call, gpu_vec.cu:543-557)
  double      *red_dev  = nullptr;   // [8]
  double      *red_host = nullptr;   // pinned [RED_HOST_DOUBLES]
};
