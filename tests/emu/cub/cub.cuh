// cub/cub.cuh -- TEST INFRASTRUCTURE: the few CUB entry points mesh.cu uses (DeviceScan::ExclusiveSum, DeviceSelect::Flagged over a
// counting / transform iterator), sequential on the host, for the CPU emulation build of the library (tests/emu/build_emu_lib.py).
// Same calling convention as CUB: a first call with a null workspace returns the workspace size.
#pragma once
#include <cstddef>
#include "cuda_emu_runtime.h"

namespace cub {

template <typename T> struct CountingInputIterator
{
  T base;
  explicit CountingInputIterator(T b) : base(b) {}
  T operator[](size_t i) const { return (T)(base + i); }
};
template <typename Value, typename Op, typename Input> struct TransformInputIterator
{
  Input in;
  Op    op;
  TransformInputIterator(Input i, Op o) : in(i), op(o) {}
  Value operator[](size_t i) const { return op(in[i]); }
};

struct DeviceScan
{
  template <typename In, typename Out> static cudaError_t ExclusiveSum(void *tmp, size_t &tmp_bytes, In in, Out out, int n, cudaStream_t = nullptr)
  {
    if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
    auto acc = decltype(in[0] + in[0])(0);
    for (int i = 0; i < n; ++i) { const auto v = in[i]; out[i] = acc; acc += v; }
    return cudaSuccess;
  }
};
struct DeviceSelect
{
  template <typename In, typename Flag, typename Out, typename Num>
  static cudaError_t Flagged(void *tmp, size_t &tmp_bytes, In in, Flag flags, Out out, Num n_selected, int n, cudaStream_t = nullptr)
  {
    if (!tmp) { tmp_bytes = 1; return cudaSuccess; }
    size_t k = 0;
    for (int i = 0; i < n; ++i) if (flags[i]) out[k++] = in[i];
    *n_selected = (decltype(*n_selected + 0))k;
    return cudaSuccess;
  }
};

}  // namespace cub
