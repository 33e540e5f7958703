// slab_common.cuh -- pieces shared by the slab-type Laplace cell kernels (kernels_slab3.cuh, kernels_stage.cuh):
//   * lane <-> (cell of the warp group, index) map and the conflict-free padded shared-memory layouts of the three
//     array transposes A (owns y,z) -> B (owns x,y) -> C (owns x,z) -> A (found by tools/slab2_layout_search.py);
//   * even-odd decomposition of the 1-D contractions (TensorOpsShmem::contraction, tensor_ops.cuh:25-117: the Gauss and
//     Gauss-Lobatto points are symmetric about the cell centre, so the interpolation matrix is centro-symmetric and the
//     collocation derivative centro-antisymmetric: 21 resp. 20 operations per line of 5 instead of 25);
//   * mbarrier / bulk-async-copy (TMA) helpers for the coefficient image of a group.
#pragma once
#include <cmath>
#include "kernels_v0.cuh"

namespace mfg {

struct Slab2Lay { int SL, SH, SI, SJ, SK; };

// merge mask of a group: bit (10 d + c) = cell c of the group hands the contributions of its upper face in direction d
// to cell c + 2^d of the same group (whose lower face has the same DoF indices) and does not scatter them itself
constexpr int SLAB2_MERGE_MAX_CW = 10;

// Lane map ("half split"): a warp group holds CW = 32/n cells, cell c = cl + HC*ch with HC = CW/2; the cells with ch = 0
// live in lanes 0..15, the others in lanes 16..31, lane = 16 ch + n cl + x.  A 64-bit gather / scatter instruction is
// processed per half warp, and this way each half touches the DoFs of HC cells only (tools/gather_line_model.py).
// For odd CW (n = 6) there is no split: lane = n c + x.
// Strides found by tools/slab2_layout_search.py: element (c,i,j,k) of a group at SL*cl + SH*ch + SI*i + SJ*j + SK*k.
// AB: conflict free for lanes (c,i) and (c,k); BC: lanes (c,k) and (c,j); CA: lanes (c,j) and (c,i).
// F = elements per buffer, a multiple of 16 bytes; the coefficient image uses the BC layout.
template <int n, int WB> struct Slab2Tab;
#define MFG_SLAB2_TAB(n_, WB_, AB_, BC_, CA_, F_)                                                  \
  template <> struct Slab2Tab<n_, WB_>                                                             \
  {                                                                                                \
    static constexpr Slab2Lay AB() { return Slab2Lay AB_; }                                        \
    static constexpr Slab2Lay BC() { return Slab2Lay BC_; }                                        \
    static constexpr Slab2Lay CA() { return Slab2Lay CA_; }                                        \
    static constexpr int F = F_;                                                                   \
  }
#define MFG_L(...) {__VA_ARGS__}
MFG_SLAB2_TAB(2, 8, MFG_L(2, 16, 1, 32, 65), MFG_L(2, 16, 32, 1, 65), MFG_L(2, 16, 1, 65, 32), 130);
MFG_SLAB2_TAB(3, 8, MFG_L(3, 45, 1, 90, 15), MFG_L(9, 135, 1, 3, 45), MFG_L(3, 45, 1, 15, 90), 270);
MFG_SLAB2_TAB(4, 8, MFG_L(4, 16, 1, 32, 129), MFG_L(4, 16, 32, 1, 129), MFG_L(4, 16, 1, 129, 32), 516);
MFG_SLAB2_TAB(5, 8, MFG_L(5, 75, 1, 150, 15), MFG_L(25, 375, 1, 5, 75), MFG_L(5, 75, 1, 15, 150), 750);
MFG_SLAB2_TAB(6, 8, MFG_L(6, 0, 1, 30, 185), MFG_L(6, 0, 30, 1, 185), MFG_L(6, 0, 1, 185, 30), 1106);
MFG_SLAB2_TAB(2, 4, MFG_L(4, 1, 2, 32, 66), MFG_L(4, 1, 32, 2, 66), MFG_L(4, 1, 2, 66, 32), 132);
MFG_SLAB2_TAB(3, 4, MFG_L(1, 15, 5, 30, 91), MFG_L(1, 15, 30, 5, 91), MFG_L(1, 15, 5, 91, 30), 272);
MFG_SLAB2_TAB(4, 4, MFG_L(4, 16, 1, 32, 129), MFG_L(4, 16, 32, 1, 129), MFG_L(4, 16, 1, 129, 32), 516);
MFG_SLAB2_TAB(5, 4, MFG_L(10, 1, 2, 150, 30), MFG_L(50, 1, 2, 10, 150), MFG_L(10, 1, 2, 30, 150), 752);
MFG_SLAB2_TAB(6, 4, MFG_L(1, 0, 5, 30, 187), MFG_L(1, 0, 30, 5, 187), MFG_L(1, 0, 5, 187, 30), 1116);
#undef MFG_L
#undef MFG_SLAB2_TAB

// lane <-> (cell in group, index x); idle lanes get cell = -1
struct Slab2Lane { int c, cl, ch, x; };
template <int n> __host__ __device__ inline Slab2Lane slab2_lane(int lane)
{
  constexpr int  CW = 32 / n, HC = CW % 2 == 0 ? CW / 2 : CW;
  constexpr bool SPLIT = CW % 2 == 0;
  const int      ch = SPLIT ? lane / 16 : 0, l16 = SPLIT ? lane % 16 : lane;
  if (l16 >= HC * n) return Slab2Lane{-1, 0, ch, 0};
  return Slab2Lane{HC * ch + l16 / n, l16 / n, ch, l16 % n};
}

// inverse of slab2_lane (any c, clamped into the warp: the result is only used where the merge mask says so)
template <int n> __host__ __device__ inline int slab2_lane_of(int c, int x)
{
  constexpr int  CW = 32 / n, HC = CW % 2 == 0 ? CW / 2 : CW;
  constexpr bool SPLIT = CW % 2 == 0;
  if (c < 0) c = 0;
  return (SPLIT ? 16 * (c / HC) + n * (c % HC) : n * c) + x;
}

// Even-odd tables of one 1-D matrix M (out[q] = sum_k M[k][q] in[k]) with M[n-1-k][n-1-q] = +-M[k][q]:
//   Ce[k][q] = (M[k][q] + M[n-1-k][q]) / 2   (k < n/2),   Ce[n/2][q] = M[n/2][q]  (n odd: the middle input)
//   Co[k][q] = (M[k][q] - M[n-1-k][q]) / 2   (k < n/2)
// rows have length m = (n+1)/2
template <typename Number, int n> struct EoTab
{
  static constexpr int h = n / 2, m = (n + 1) / 2;
  Number Ce[(h + 1) * m];
  Number Co[h * m];
};
template <typename Number, int n> struct EoMats { EoTab<Number, n> N, NT, D, DT; };

// even-odd tables of M (TR: of its transpose); sign = +1 centro-symmetric, -1 centro-antisymmetric
template <typename Number, int n> inline void make_eo(const double *M, bool TR, int sign, EoTab<Number, n> &T)
{
  constexpr int h = n / 2, m = (n + 1) / 2;
  auto at = [&](int k, int q) { return TR ? M[q * n + k] : M[k * n + q]; };
  double scale = 0;
  for (int i = 0; i < n * n; ++i) scale = std::max(scale, std::fabs(M[i]));
  for (int k = 0; k < n; ++k)
    for (int q = 0; q < n; ++q)
      if (std::fabs(at(k, q) - sign * at(n - 1 - k, n - 1 - q)) > 1e-12 * scale)
        throw Error(MFG_ERR_UNSUPPORTED, "slab2 kernel: 1-D shape matrices are not centro-(anti)symmetric");
  for (int i = 0; i < (h + 1) * m; ++i) T.Ce[i] = 0;
  for (int i = 0; i < h * m; ++i) T.Co[i] = 0;
  for (int k = 0; k < h; ++k)
    for (int q = 0; q < m; ++q)
      {
        T.Ce[k * m + q] = (Number)(0.5 * (at(k, q) + at(n - 1 - k, q)));
        T.Co[k * m + q] = (Number)(0.5 * (at(k, q) - at(n - 1 - k, q)));
      }
  if (n & 1)
    for (int q = 0; q < m; ++q) T.Ce[h * m + q] = (Number)at(h, q);
}

template <typename Number, int n> inline void make_eo_tables(const double *N, const double *D, EoMats<Number, n> &em)
{
  make_eo<Number, n>(N, false, +1, em.N);
  make_eo<Number, n>(N, true, +1, em.NT);
  make_eo<Number, n>(D, false, -1, em.D);
  make_eo<Number, n>(D, true, -1, em.DT);
}

// out = M^T-contraction of one line.  ANTI = false: centro-symmetric M (interpolation), true: centro-antisymmetric
// (collocation derivative)
template <int n, bool ANTI, typename Number, bool NOP = false>
__device__ __forceinline__ void eo_apply(const EoTab<Number, n> &T, const Number (&in)[n], Number (&out)[n])
{
  if (NOP)
    {
#pragma unroll
      for (int q = 0; q < n; ++q) out[q] = in[q];
      return;
    }
  constexpr int  h = n / 2, m = (n + 1) / 2;
  constexpr bool odd = n & 1;
  constexpr int  qe = ANTI ? h : m;  // outputs fed by the even part of the input (+ the middle input)
  constexpr int  qo = ANTI ? m : h;  // outputs fed by the odd part
  Number e[h], o[h], P[m], R[m];
#pragma unroll
  for (int k = 0; k < h; ++k) { e[k] = in[k] + in[n - 1 - k]; o[k] = in[k] - in[n - 1 - k]; }
#pragma unroll
  for (int q = 0; q < qe; ++q) P[q] = T.Ce[q] * e[0];
#pragma unroll
  for (int k = 1; k < h; ++k)
#pragma unroll
    for (int q = 0; q < qe; ++q) P[q] = fma(T.Ce[k * m + q], e[k], P[q]);
  if (odd)
    {
#pragma unroll
      for (int q = 0; q < qe; ++q) P[q] = fma(T.Ce[h * m + q], in[h], P[q]);
    }
#pragma unroll
  for (int q = 0; q < qo; ++q) R[q] = T.Co[q] * o[0];
#pragma unroll
  for (int k = 1; k < h; ++k)
#pragma unroll
    for (int q = 0; q < qo; ++q) R[q] = fma(T.Co[k * m + q], o[k], R[q]);
#pragma unroll
  for (int q = 0; q < h; ++q)
    {
      out[q]         = ANTI ? R[q] + P[q] : P[q] + R[q];
      out[n - 1 - q] = ANTI ? R[q] - P[q] : P[q] - R[q];
    }
  if (odd) out[h] = ANTI ? R[h] : P[h];
}

// contraction of every line of a slab held in registers: contracted index has register stride S, lines stride T
template <int n, int S, int T, bool ANTI, typename Number, bool NOP = false>
__device__ __forceinline__ void slab2_apply(const EoTab<Number, n> &M, Number (&v)[n * n])
{
  if (NOP) return;
#pragma unroll
  for (int l = 0; l < n; ++l)
    {
      Number in[n], out[n];
#pragma unroll
      for (int e = 0; e < n; ++e) in[e] = v[l * T + e * S];
      eo_apply<n, ANTI, Number, NOP>(M, in, out);
#pragma unroll
      for (int e = 0; e < n; ++e) v[l * T + e * S] = out[e];
    }
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(a), "r"(parity) : "memory");
}
// one bulk-async copy global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
  const unsigned b = (unsigned)__cvta_generic_to_shared(bar), d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b)
               : "memory");
}

// layout of the kernel's private arrays (for the builders in operators.cu)
struct Slab2Geom { int n, cw, hc, cwf; Slab2Lay bc; };
bool      slab2_supported(int dim, int degree, mfg_dtype dt);  // (kernels_slab3_inst.cu)
Slab2Geom slab2_geom(int degree, mfg_dtype dt);

}  // namespace mfg
