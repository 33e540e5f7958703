// gpu_vec.h -- header-only C++ facade over the C ABI (mfgpu.h) with the reference's class names:
//   GpuVector<Number>   matrix_free_gpu/gpu_vec.h:22-176       GpuList<T>   matrix_free_gpu/gpu_list.h:6-34
// Ownership and semantics follow the reference: RAII, deep copies, swap exchanges pointers, GpuVector(n)
// zero-fills, resize(n) does not.  CUDA / library errors throw std::runtime_error (the reference throws
// dealii::ExcMessage, cuda_utils.cuh:15-23).
#pragma once
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>
#include "../mfgpu.h"

namespace dealii_cuda_b200 {

inline void check(int rc)
{
  if (rc != MFG_OK) throw std::runtime_error(std::string("mfgpu: ") + mfg_last_error());
}

// process-wide default context (device 0, legacy default stream: what the reference uses everywhere)
inline mfg_ctx *&default_context()
{
  static mfg_ctx *ctx = nullptr;
  if (!ctx) check(mfg_ctx_create(0, nullptr, &ctx));
  return ctx;
}

template <typename Number> constexpr mfg_dtype dtype_of()
{
  static_assert(std::is_same<Number, float>::value || std::is_same<Number, double>::value, "Number must be float or double");
  return std::is_same<Number, double>::value ? MFG_F64 : MFG_F32;
}

template <typename Number> class GpuVector
{
public:
  typedef Number       value_type;
  typedef unsigned int size_type;

  GpuVector() { check(mfg_vec_create(default_context(), dtype_of<Number>(), 0, &v_)); }
  explicit GpuVector(unsigned int n) { check(mfg_vec_create(default_context(), dtype_of<Number>(), n, &v_)); }
  GpuVector(const GpuVector &o) : GpuVector() { check(mfg_vec_copy(v_, o.v_)); }
  template <typename Other> GpuVector(const GpuVector<Other> &o) : GpuVector() { check(mfg_vec_copy(v_, o.handle())); }
  explicit GpuVector(const std::vector<Number> &h) : GpuVector((unsigned int)h.size()) { fromHost(h.data(), (unsigned int)h.size()); }
  // non-owning view of a vector that lives inside a library object (e.g. the operator's inverse diagonal): same interface,
  // the handle is not destroyed with the view
  static GpuVector borrowed(mfg_vec *v) { return GpuVector(v, false); }
  GpuVector(GpuVector &&o) noexcept : v_(o.v_), owned_(o.owned_) { o.v_ = nullptr; }
  ~GpuVector() { if (v_ && owned_) mfg_vec_destroy(v_); }

  GpuVector &operator=(const GpuVector &o) { check(mfg_vec_copy(v_, o.v_)); return *this; }
  template <typename Other> GpuVector &operator=(const GpuVector<Other> &o) { check(mfg_vec_copy(v_, o.handle())); return *this; }
  GpuVector &operator=(const std::vector<Number> &h) { resize((unsigned int)h.size()); fromHost(h.data(), (unsigned int)h.size()); return *this; }
  GpuVector &operator=(const Number a) { check(mfg_vec_fill(v_, (double)a)); return *this; }

  unsigned int  size() const { return (unsigned int)mfg_vec_size(v_); }
  Number       *getData() { return static_cast<Number *>(mfg_vec_data(v_)); }
  const Number *getDataRO() const { return static_cast<const Number *>(mfg_vec_data(v_)); }
  mfg_vec      *handle() const { return v_; }

  void resize(unsigned int n) { check(mfg_vec_resize(v_, n)); }
  void reinit(unsigned int n, bool leave_elements_uninitialized = false) { resize(n); if (!leave_elements_uninitialized) *this = Number(0); }
  void reinit(const GpuVector &o, bool leave_elements_uninitialized = false) { reinit(o.size(), leave_elements_uninitialized); }
  void fromHost(const Number *buf, unsigned int n) { check(mfg_vec_from_host(v_, buf, n)); }
  void copyToHost(std::vector<Number> &dst) const { dst.resize(size()); check(mfg_vec_to_host(v_, dst.data(), dst.size())); }
  std::vector<Number> toVector() const { std::vector<Number> h; copyToHost(h); return h; }

  Number operator*(const GpuVector &o) const { double r; check(mfg_vec_dot(v_, o.v_, &r)); return (Number)r; }
  void   add(const GpuVector &x) { sadd(1, 1, x); }
  void   add(const Number a, const GpuVector &x) { sadd(1, a, x); }
  void   sadd(const Number s, const GpuVector &x) { sadd(s, 1, x); }
  void   sadd(const Number s, const Number a, const GpuVector &x) { check(mfg_vec_sadd(v_, (double)s, (double)a, x.v_)); }
  GpuVector &operator+=(const GpuVector &x) { sadd(1, 1, x); return *this; }
  GpuVector &operator-=(const GpuVector &x) { sadd(1, -1, x); return *this; }
  Number add_and_dot(const Number a, const GpuVector &x, const GpuVector &w)
  {
    double r; check(mfg_vec_add_and_dot(v_, (double)a, x.v_, w.v_, &r)); return (Number)r;
  }
  void       scale(const GpuVector &x) { check(mfg_vec_scale(v_, x.v_)); }
  GpuVector &operator/=(const GpuVector &x) { check(mfg_vec_divide(v_, x.v_)); return *this; }
  GpuVector &invert() { check(mfg_vec_invert(v_)); return *this; }
  void       equ(const Number a, const GpuVector &x) { check(mfg_vec_equ(v_, (double)a, x.v_)); }
  GpuVector &operator*=(const Number a) { check(mfg_vec_scal(v_, (double)a)); return *this; }
  Number     l2_norm() const { double r; check(mfg_vec_l2_norm(v_, &r)); return (Number)r; }
  bool       all_zero() const { int z; check(mfg_vec_all_zero(v_, &z)); return z != 0; }
  unsigned int memory_consumption() const { return size() * sizeof(Number); }
  void       swap(GpuVector &o) { check(mfg_vec_swap(v_, o.v_)); }
  void       compress() const {}  // gpu_vec.h:175

private:
  GpuVector(mfg_vec *v, bool owned) : v_(v), owned_(owned) {}
  mfg_vec *v_ = nullptr;
  bool     owned_ = true;
};

// DiagonalMatrix<GpuVector<Number>> as the reference hands it to PreconditionChebyshev (laplace_operator_gpu.h:79, 423-429):
// vmult = pointwise product with the stored vector
template <typename VectorType> class DiagonalMatrix   // (VectorType = GpuVector<Number>, as deal.II's DiagonalMatrix<VectorType>)
{
public:
  explicit DiagonalMatrix(VectorType &&diag) : diag_(std::move(diag)) {}
  const VectorType &get_vector() const { return diag_; }
  void vmult(VectorType &dst, const VectorType &src) const { dst = src; dst.scale(diag_); }
  unsigned int m() const { return diag_.size(); }

private:
  VectorType diag_;
};

// GpuList<T>: immutable device index array.  Only what the facade needs: the index lists live inside the
// library objects, GpuList keeps the host copy that is handed to them.
template <typename T> class GpuList
{
public:
  GpuList() {}
  GpuList(const std::vector<T> &h) : host_(h) {}
  GpuList &operator=(const std::vector<T> &h) { host_ = h; return *this; }
  void               clear() { host_.clear(); }
  unsigned int       size() const { return (unsigned int)host_.size(); }
  const std::vector<T> &host() const { return host_; }
  std::size_t        memory_consumption() const { return host_.size() * sizeof(T); }

private:
  std::vector<T> host_;
};

}  // namespace dealii_cuda_b200
