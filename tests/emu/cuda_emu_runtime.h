// cuda_emu_runtime.h -- TEST INFRASTRUCTURE: stand-in for <cuda_runtime.h> when the library's sources are compiled for the CPU
// emulation (tests/emu/build_emu_lib.py): "device" memory is host memory, streams and events do nothing (every launch completes
// before emu_launch returns), kernels run on the thread-per-CUDA-thread executor of cuda_emu.h.
#pragma once
#include <chrono>
#include <cstdlib>
#include "cuda_emu.h"

typedef int   cudaError_t;
typedef void *cudaStream_t;
struct emu_event { double t; };
typedef emu_event *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount, major, minor, l2CacheSize; char name[64]; };
struct float4 { float x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

enum { cudaErrorInvalidConfiguration = 9, cudaErrorIllegalAddress = 700 };
inline const char *cudaGetErrorString(cudaError_t e)
{
  return e == cudaSuccess ? "no error" : e == cudaErrorInvalidConfiguration ? "emulation: invalid launch configuration (block size / dynamic shared memory without opt-in)" :
         e == cudaErrorIllegalAddress ? "emulation: a block wrote behind its dynamic shared memory" : "emulated CUDA error";
}
inline cudaError_t cudaGetLastError()
{
  const int e = emu_launch_error;
  emu_launch_error = 0;
  return e == 1 ? cudaErrorInvalidConfiguration : e == 2 ? cudaErrorIllegalAddress : cudaSuccess;
}
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMalloc(reinterpret_cast<void **>(p), n); }
inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaGetLastError(); }   // (a failed launch surfaces at the next synchronisation, as on the device)
inline cudaError_t cudaDeviceSynchronize() { return cudaGetLastError(); }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = emu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
  std::memset(p, 0, sizeof(*p));
  p->multiProcessorCount = 2; p->major = 10; p->minor = 0; p->l2CacheSize = 1 << 20;
  std::strcpy(p->name, "CPU emulation");
  return cudaSuccess;
}
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute a, int v)
{
  if (a == cudaFuncAttributeMaxDynamicSharedMemorySize) emu_max_dyn_smem_optin = std::max(emu_max_dyn_smem_optin, (size_t)v);
  return cudaSuccess;
}
template <typename F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *b, F, int, size_t) { *b = 1; return cudaSuccess; }

#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __forceinline__ inline
template <typename T> inline T __ldg(const T *p) { return *p; }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned __cvta_generic_to_shared(const void *) { return 0; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_arrive_and_wait(emu_blk.current->warp->sync); }   // (full mask: every lane of the warp calls)

// __shfl_sync with a full mask: every lane publishes its value, reads the source lane's (warp rendezvous on both sides)
template <typename T> inline T __shfl_sync(unsigned, T v, int src_lane)
{
  emu_warp_ctx *w = emu_blk.current->warp;
  const int     lane = (int)(threadIdx.x & 31u);
  w->slot[lane] = (double)v;
  emu_arrive_and_wait(w->sync);
  const T r = (T)w->slot[src_lane & 31];
  emu_arrive_and_wait(w->sync);
  return r;
}

// cudaLaunchKernelEx with launch attributes (programmatic dependent launch): every emulated launch completes before it returns, so
// the attributes have nothing to order
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute *attrs; unsigned numAttrs; };
template <typename Kernel, typename... Args> inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, Kernel kernel, Args... args)
{
  emu_launch4(cfg->gridDim.x, cfg->blockDim.x, cfg->dynamicSmemBytes, cfg->stream, [&](auto... a) { kernel(a...); }, args...);
  return cudaGetLastError();
}
