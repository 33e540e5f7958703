// cuda_emu.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run the device code of the header-only generic path
// (include/dealii_cuda_b200/fee_gpu.cuh: FEEvaluationGpu, apply_kernel_shmem) on the CPU, so that its gather / interpolation /
// contraction / scatter logic is exercised in the CPU test suite.  One OS thread per CUDA thread of a block, blocks one after the
// other; __syncthreads is a barrier over the block; atomicAdd a compare-and-swap loop; dynamic shared memory is one static buffer
// defined by the user of this header (only one block is alive at a time); __shfl_down_sync through a per-warp slot array.  No streams.
#pragma once
#include <algorithm>
#include <atomic>
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

// kernel<<<grid, block>>>(args...): blocks sequentially, the threads of a block concurrently
template <typename Kernel, typename... Args> void emu_launch(unsigned grid, unsigned block, Kernel kernel, Args... args)
{
  for (unsigned b = 0; b < grid; ++b)
    {
      std::barrier<>           bar((std::ptrdiff_t)block);
      std::vector<std::unique_ptr<emu_warp_ctx>> warps;
      for (unsigned w = 0; w * 32 < block; ++w) warps.emplace_back(new emu_warp_ctx((int)std::min(32u, block - w * 32)));
      std::vector<std::thread> threads;
      threads.reserve(block);
      for (unsigned t = 0; t < block; ++t)
        threads.emplace_back([&, t] {
          threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
          emu_block_barrier = &bar;
          emu_warp = warps[t / 32].get();
          kernel(args...);
        });
      for (auto &th : threads) th.join();
    }
}
