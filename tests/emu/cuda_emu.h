// cuda_emu.h -- TEST INFRASTRUCTURE: just enough of the CUDA execution model to run device code on the CPU (the header-only generic
// path, and -- through tests/emu/build_emu_lib.py -- the library's own kernels), so that gather / interpolation / contraction / scatter
// logic, barriers and launch arithmetic are exercised in the CPU test suite.  Blocks one after the other; the CUDA threads of a
// block are fibers of the calling OS thread, switched at __syncthreads / warp shuffles; atomicAdd is a plain read-modify-write
// (one OS thread); dynamic shared memory is one static buffer (only one block is alive at a time).  No streams.
#pragma once
#include <sys/mman.h>
#include <ucontext.h>
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
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

// Execution model: blocks one after the other; the CUDA threads of a block are FIBERS (ucontext) of the calling OS thread, switched
// only at synchronisation points.  A barrier counts arrivals and yields until its generation changes; a fiber that leaves the kernel
// no longer takes part (as on the device).  One OS thread runs everything, so atomics are plain read-modify-writes and nothing
// spends time in futexes (the first version used one OS thread per CUDA thread and std::barrier: correct, but 2/3 of the run time of
// the emulation tests was system time).
struct emu_sync
{
  int           count = 0, alive = 0;
  unsigned long generation = 0;
};
struct emu_warp_ctx
{
  emu_sync sync;
  double   slot[32];
};
struct emu_fiber
{
  ucontext_t    ctx;
  char         *stack = nullptr;
  bool          done = true;
  unsigned      tid = 0;
  emu_warp_ctx *warp = nullptr;
};
struct emu_block
{
  ucontext_t                   scheduler;
  std::vector<emu_fiber>       fibers;
  std::vector<emu_warp_ctx>    warps;
  emu_sync                     sync;
  emu_fiber                   *current = nullptr;
  const std::function<void()> *body = nullptr;
};
inline emu_block emu_blk;   // (one block is alive at a time)

inline void emu_yield() { swapcontext(&emu_blk.current->ctx, &emu_blk.scheduler); }
inline void emu_arrive_and_wait(emu_sync &s)
{
  const unsigned long g = s.generation;
  if (++s.count >= s.alive) { s.count = 0; ++s.generation; }
  else while (s.generation == g) emu_yield();
}
inline void emu_leave(emu_sync &s)
{
  --s.alive;
  if (s.alive > 0 && s.count >= s.alive) { s.count = 0; ++s.generation; }
}
inline void __syncthreads() { emu_arrive_and_wait(emu_blk.sync); }

// warp shuffles: the lanes of a warp exchange through a slot array, a warp-wide rendezvous on both sides (every lane of the warp
// must call, as on the device with a full mask)
template <typename T> inline T __shfl_down_sync(unsigned, T v, int delta)
{
  emu_warp_ctx *w = emu_blk.current->warp;
  const int     lane = (int)(threadIdx.x & 31u), n = (int)std::min<unsigned>(32u, blockDim.x - (threadIdx.x & ~31u));
  w->slot[lane] = (double)v;
  emu_arrive_and_wait(w->sync);
  const T r = lane + delta < n ? (T)w->slot[lane + delta] : v;
  emu_arrive_and_wait(w->sync);
  return r;
}

template <typename T> inline T atomicAdd(T *addr, T val)
{
  const T old = *addr;
  *addr = old + val;
  return old;
}

// dynamic shared memory of the running block for sources transformed by build_emu_lib.py (`extern __shared__ T x[]` becomes a pointer to it)
alignas(16) inline unsigned char emu_dyn_smem[256 * 1024];

inline void emu_fiber_entry()
{
  emu_fiber *f = emu_blk.current;
  (*emu_blk.body)();
  emu_leave(f->warp->sync);
  emu_leave(emu_blk.sync);
  f->done = true;
  swapcontext(&f->ctx, &emu_blk.scheduler);
}

// kernel<<<grid, block>>>(args...)
template <typename Kernel, typename... Args> void emu_launch(unsigned grid, unsigned block, Kernel kernel, Args... args)
{
  constexpr size_t STACK = 256 * 1024;   // (mapped without reserve: only what a fiber touches becomes resident)
  emu_block &B = emu_blk;
  if (B.fibers.size() < block) B.fibers.resize(block);
  for (unsigned t = 0; t < block; ++t)
    if (!B.fibers[t].stack)
      {
        void *p = mmap(nullptr, STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
        if (p == MAP_FAILED) std::abort();
        B.fibers[t].stack = static_cast<char *>(p);
      }
  const std::function<void()> body = [&]() { kernel(args...); };
  B.body = &body;
  const unsigned n_warps = (block + 31) / 32;
  if (B.warps.size() < n_warps) B.warps.resize(n_warps);
  for (unsigned b = 0; b < grid; ++b)
    {
      blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
      B.sync = emu_sync();
      B.sync.alive = (int)block;
      for (unsigned w = 0; w < n_warps; ++w)
        {
          B.warps[w].sync = emu_sync();
          B.warps[w].sync.alive = (int)std::min(32u, block - w * 32);
        }
      for (unsigned t = 0; t < block; ++t)
        {
          emu_fiber &f = B.fibers[t];
          f.done = false; f.tid = t; f.warp = &B.warps[t / 32];
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = STACK;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, emu_fiber_entry, 0);
        }
      unsigned remaining = block;
      while (remaining)
        for (unsigned t = 0; t < block; ++t)
          {
            emu_fiber &f = B.fibers[t];
            if (f.done) continue;
            B.current = &f;
            threadIdx.x = t;
            swapcontext(&B.scheduler, &f.ctx);
            if (f.done) --remaining;
          }
    }
  B.body = nullptr;
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
