// vector.cuh -- GpuVector<Number> storage (gpu_vec.h:22-176) behind the C ABI.
#pragma once
#include <algorithm>
#include "common.cuh"

struct mfg_vec
{
  mfg_ctx  *ctx  = nullptr;
  mfg_dtype dt   = MFG_F64;
  size_t    n    = 0;     // logical size
  size_t    cap  = 0;     // allocated elements (owning vectors only)
  void     *p    = nullptr;
  bool      owns = true;
  size_t    esize() const { return dt == MFG_F64 ? 8 : 4; }
};

namespace mfg {
constexpr int RED_SCRATCH_DOUBLES = 8 + 1024;
constexpr int RED_HOST_DOUBLES = 64;  // pinned host scratch of a context (reduction results, solver state mirror)
void   vec_fill(mfg_vec *v, double a);
void   vec_scal(mfg_vec *v, double a);
void   vec_invert(mfg_vec *v);
void   vec_sadd(mfg_vec *v, double s, double a, const mfg_vec *x);
void   vec_equ(mfg_vec *v, double a, const mfg_vec *x);
void   vec_scale(mfg_vec *v, const mfg_vec *x);
void   vec_divide(mfg_vec *v, const mfg_vec *x);
double vec_dot(const mfg_vec *a, const mfg_vec *b);
double vec_add_and_dot(mfg_vec *v, double alpha, const mfg_vec *x, const mfg_vec *w);
bool   vec_all_zero(const mfg_vec *v);
void   vec_copy_with_indices(mfg_vec *dst, const mfg_vec *src, const uint32_t *di, const uint32_t *si, size_t n);
}  // namespace mfg
