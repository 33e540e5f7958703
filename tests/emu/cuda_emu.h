// cuda_emu.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the device code of the header-only generic path
// (include/dealii_cuda_b200/fee_gpu.cuh: FEEvaluationGpu, apply_kernel_shmem) on the CPU, so that its gather / interpolation /
// contraction / scatter logic is exercised in the CPU test suite.  One OS thread per CUDA thread of a block, blocks one after the
// other; __syncthreads is a barrier over the block; atomicAdd a compare-and-swap loop; dynamic shared memory is one static buffer
// defined by the user of this header (only one block is alive at a time); __shfl_down_sync through a per-warp slot array.  No streams.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __grid_constant__
#define __shared__
#define __align__(n) __attribute__((aligned(n)))

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
inline thread_local std::barrier<> *emu_block_barrier = nullptr;

inline void __syncthreads() { emu_block_barrier->arrive_and_wait(); }

// warp shuffles: the 32 lanes of a warp exchange through a slot array, a barrier over the warp on both sides (every lane of the
// warp must call, as on the device with a full mask)
struct emu_warp_ctx
{
  explicit emu_warp_ctx(int lanes) : bar(lanes), n(lanes) {}
  std::barrier<> bar;
  int            n;
  double         slot[32];
};
inline thread_local emu_warp_ctx *emu_warp = nullptr;
template <typename T> inline T __shfl_down_sync(unsigned, T v, int delta)
{
  const int lane = (int)(threadIdx.x & 31u);
  emu_warp->slot[lane] = (double)v;
  emu_warp->bar.arrive_and_wait();
  const T r = lane + delta < emu_warp->n ? (T)emu_warp->slot[lane + delta] : v;
  emu_warp->bar.arrive_and_wait();
  return r;
}

template <typename T> inline T atomicAdd(T *addr, T val)
{
  std::atomic_ref<T> a(*addr);
  T old = a.load();
  while (!a.compare_exchange_weak(old, old + val)) {}
  return old;
}

// dynamic shared memory of the running block for sources transformed by build_emu_lib.py (`extern __shared__ T x[]` becomes a pointer to it)
alignas(16) inline unsigned char emu_dyn_smem[256 * 1024];

// worker pool: OS threads are created once and reused by every launch (a CG solve launches thousands of small kernels)
class emu_pool
{
public:
  static emu_pool &get() { static emu_pool p; return p; }
  // run fn(t) for t = 0 .. n-1 on n workers concurrently, return when all are done
  void run(unsigned n, const std::function<void(unsigned)> &fn)
  {
    grow(n);
    task_ = &fn;
    remaining_.store((int)n);
    {
      std::lock_guard<std::mutex> lk(m_);
      want_ = n; ++generation_;
    }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return remaining_.load() == 0; });
  }
  ~emu_pool()
  {
    { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++generation_; }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
  }
private:
  void grow(unsigned n)
  {
    while (workers_.size() < n)
      {
        const unsigned id = (unsigned)workers_.size();
        unsigned long   seen;
        { std::lock_guard<std::mutex> lk(m_); seen = generation_; }
        workers_.emplace_back([this, id, seen]() mutable {
          for (;;)
            {
              std::unique_lock<std::mutex> lk(m_);
              cv_.wait(lk, [&] { return generation_ != seen; });
              seen = generation_;
              if (stop_) return;
              const bool mine = id < want_;
              lk.unlock();
              if (!mine) continue;
              (*task_)(id);
              if (remaining_.fetch_sub(1) == 1) { std::lock_guard<std::mutex> g(m_); done_cv_.notify_all(); }
            }
        });
      }
  }
  std::vector<std::thread> workers_;
  std::mutex               m_;
  std::condition_variable  cv_, done_cv_;
  unsigned long            generation_ = 0;
  unsigned                 want_ = 0;
  bool                     stop_ = false;
  std::atomic<int>         remaining_{0};
  const std::function<void(unsigned)> *task_ = nullptr;
};

// kernel<<<grid, block>>>(args...): blocks sequentially, the threads of a block concurrently
template <typename Kernel, typename... Args> void emu_launch(unsigned grid, unsigned block, Kernel kernel, Args... args)
{
  for (unsigned b = 0; b < grid; ++b)
    {
      std::barrier<>           bar((std::ptrdiff_t)block);
      std::vector<std::unique_ptr<emu_warp_ctx>> warps;
      for (unsigned w = 0; w * 32 < block; ++w) warps.emplace_back(new emu_warp_ctx((int)std::min(32u, block - w * 32)));
      emu_pool::get().run(block, [&](unsigned t) {
        threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
        emu_block_barrier = &bar;
        emu_warp = warps[t / 32].get();
        kernel(args...);
        // a thread that has left the kernel no longer takes part in the barriers of its block / warp (as on the device)
        emu_warp->bar.arrive_and_drop();
        bar.arrive_and_drop();
      });
    }
}
// the transformed form of kernel<<<grid, block, smem, stream>>>(args...).  The launch limits of the device are enforced: at most
// 1024 threads per block, dynamic shared memory above 48 KB only after an opt-in (cudaFuncSetAttribute records the largest value
// asked for; the emulation cannot tell the kernels apart), never above 227 KB; a canary behind the `smem` bytes of the launch
// catches a block that writes past what the host code reserved for it.
inline size_t emu_max_dyn_smem_optin = 0;
inline int    emu_launch_error = 0;   // sticky, read by cudaGetLastError of cuda_emu_runtime.h: 1 = invalid configuration, 2 = smem overrun
template <typename Kernel, typename... Args> void emu_launch4(unsigned grid, unsigned block, size_t smem, void *, Kernel kernel, Args... args)
{
  if (block == 0 || block > 1024 || grid == 0 || smem > 227 * 1024 || (smem > 48 * 1024 && smem > emu_max_dyn_smem_optin)) { emu_launch_error = 1; return; }
  const size_t guard = 4096;
  std::memset(emu_dyn_smem + smem, 0xA5, std::min(guard, sizeof(emu_dyn_smem) - smem));
  emu_launch(grid, block, kernel, args...);
  for (size_t i = 0; i < std::min(guard, sizeof(emu_dyn_smem) - smem); ++i)
    if (emu_dyn_smem[smem + i] != 0xA5) { emu_launch_error = 2; break; }
}
