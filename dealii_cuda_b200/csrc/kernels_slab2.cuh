// kernels_slab2.cuh -- second-generation slab Laplace cell kernel for 3D, n = p+1 <= 6.
//
// Same operator as kernels_v0.cuh / kernels_slab.cuh (the reference's apply_kernel_shmem<LocalOperator>,
// matrix_free_gpu.h:318-341 + fee_gpu.cuh:197-365 + tensor_ops.cuh:179-261).  Built from the ncu profile of the
// first slab kernel (profiles/r01_slab_kernel_8warps_q4_f64_r6_ncu.txt): 224 L1-data-pipe wavefronts per cell
// (172 shared + 52 global) at 65 % of the pipe, FP64 pipe 41 %.  Changes:
//   * the kernel's own copies of the index map and of the merged coefficient are stored in the order the threads
//     consume them (built once at setup, operators.cu):
//       idxP [group][slot j + n k][32 lanes]     one fully coalesced 128-byte row per gather / scatter instruction,
//                                                every lane reads the DoF indices of its OWN slab -> the gathered
//                                                values land in registers, no staging round trip through shared memory;
//       cwP  [group][shared-memory image]        one bulk-async copy (TMA, cp.async.bulk + mbarrier) per group puts the
//                                                coefficient block into shared memory without touching registers or
//                                                the LSU store path; the image is conflict free for BOTH layouts that
//                                                read it;
//   * 4 transposes per cell tensor entry instead of 7 shared-memory round trips:
//       A (owns y,z) --T1--> B (owns x,y) --T2: u_q and the x,y part of the result--> C (owns x,z) --T3--> A
//   * even-odd decomposition of every 1-D contraction (the Gauss / Gauss-Lobatto points are symmetric about the cell
//     centre, so the interpolation matrix is centro-symmetric and the collocation derivative centro-antisymmetric):
//     21 resp. 20 FP64 operations per line of 5 instead of 25;
//   * results are scattered straight from registers with red.global.add;
//   * cells of a group that share a face (x-, y-, z-neighbours inside the group, detected at setup by comparing the
//     index rows, operators.cu build_slab2_merge) sum their face contributions with warp shuffles first and only one
//     of them issues the red: the scatter is bound by the L1 -> L2 request rate (one 32-byte sector per cycle,
//     profiles/r01_slab2_ablation.jsonl), and this removes 27 % of the sectors of a 6-cell group.
// Thread layouts (a warp = CW = 32/n cells, lane <-> (cell c, index x), see slab2_lane):
//   A: x = i, registers (j,k)     B: x = k, registers (i,j)     C: x = j, registers (i,k)
// Sequence per group of cells:
//   gather -> A: N_y N_z -> B: N_x, quadrature phases x and y -> C: quadrature phase z, sum, N_x^T N_z^T
//          -> A: N_y^T -> red.add
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

// CFG: 0 = 3 blocks x 4 warps per SM (168 registers), 2 transpose buffers per warp
//      1 = 2 blocks x 4 warps (255 registers), 2 buffers
//      2 = 4 blocks x 4 warps (128 registers), 1 buffer
//      3 = 3 blocks x 4 warps (168 registers), 1 buffer (more L1 left for the gather)
template <int n, typename Number, int CFG> struct Slab2Cfg
{
  static constexpr int WB  = (int)sizeof(Number);
  using Tab = Slab2Tab<n, WB>;
  static constexpr int CW  = 32 / n;   // cells per warp group
  static constexpr int NPC = n * n * n;
  static constexpr int NS  = n * n;
  static constexpr int WPB = 4;
  static constexpr int OCC  = CFG % 4;
  static constexpr bool TEX = (CFG / 4) % 2;  // gather src through the texture pipe (tex1Dfetch) instead of the LSU pipe
  // software pipeline: 0 = none; 1 = the index rows of the next group are loaded during the C phase of the current one
  // and the rows needed by the scatter are re-read before N_y^T; 2 = in addition the gather of the next group is issued
  // before the scatter of the current one
  static constexpr int PF   = (CFG % 32) / 8;
  // + 256: kernel with the in-group face merge of the scatter compiled in (costs registers: 24 bytes of spills at Q4 FP64
  // with 168 registers, so the default FP64 kernel is built without it)
  static constexpr bool MERGE = (CFG / 256) % 2 != 0;
  // + 512: gather and scatter in plane layouts (one cell per pass, lane <-> a point of the (i,j) resp. (i,k) plane, the
  // third index in registers): an instruction then covers contiguous runs of 9 and 3 DoFs instead of 3 and 1, 71 / 76
  // sectors per cell instead of 94.  Dense cell-major transpose buffers; conflict free for n = 5.
  static constexpr bool KGS = (CFG / 512) % 2 != 0;
  // measurement-only ablations (tools/ablate.py; results are wrong): 1 = no src gather, 2 = no scatter, 4 = no contractions
  static constexpr int ABL  = (CFG / 32) % 8;
  static constexpr int MINB = OCC == 1 ? 2 : OCC == 2 ? 4 : 3;
  static constexpr int NBUF = (OCC == 2 || OCC == 3) ? 1 : 2;
  static constexpr int F   = Tab::F;
  static constexpr int CWF = Tab::F;  // the coefficient image has the BC layout
  static constexpr int PER_WARP = CWF + NBUF * F;    // elements
  static constexpr size_t SMEM = 16 * WPB /* mbarriers */ + (size_t)WPB * PER_WARP * sizeof(Number);
  static constexpr uint32_t CW_BYTES = CWF * WB;
  static_assert(CW_BYTES % 16 == 0 && (F * WB) % 16 == 0, "bulk copy size must be a multiple of 16 bytes");
};

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

template <typename Number> __device__ __forceinline__ Number tex_fetch(cudaTextureObject_t tex, uint32_t i);
template <> __device__ __forceinline__ double tex_fetch<double>(cudaTextureObject_t tex, uint32_t i)
{
  const int2 t = tex1Dfetch<int2>(tex, (int)i);
  return __hiloint2double(t.y, t.x);
}
template <> __device__ __forceinline__ float tex_fetch<float>(cudaTextureObject_t tex, uint32_t i) { return tex1Dfetch<float>(tex, (int)i); }

#ifdef MFG_SLAB2_ABLATE
__constant__ int g_slab2_delay[2];  // experiment: start-up stagger {ns per step, mode}
#endif

template <int n, typename Number, int CFG>
__global__ void __launch_bounds__(Slab2Cfg<n, Number, CFG>::WPB * 32, Slab2Cfg<n, Number, CFG>::MINB)
laplace_cell_slab2(const uint32_t *__restrict__ idxP, const Number *__restrict__ cwP, const Number *__restrict__ src,
                   Number *__restrict__ dst, const uint32_t n_groups, const __grid_constant__ EoMats<Number, n> em,
                   const cudaTextureObject_t tex, const uint32_t *__restrict__ mergeP, const uint32_t *__restrict__ glist,
                   const int dep_wait, const uint32_t *__restrict__ idxLex, const uint32_t *__restrict__ idxJ, const uint32_t n_cells)
{
  using Cfg = Slab2Cfg<n, Number, CFG>;
  using Tab = typename Cfg::Tab;
  constexpr int NS = Cfg::NS;
  constexpr bool NOPC = (Cfg::ABL & 4) != 0;
  constexpr int HCn = (Cfg::CW % 2 == 0 ? Cfg::CW / 2 : Cfg::CW) * n * n * n;
  // KGS: dense buffers, element (c,i,j,k) at n^3 c + n^2 k + n j + i (written by the gather planes, read by B) and at
  // n^3 c + n^2 j + n k + i (written by C, read by the scatter planes)
  constexpr Slab2Lay AB = Cfg::KGS ? Slab2Lay{n * n * n, HCn, 1, n, n * n} : Tab::AB(), BC = Tab::BC(),
                     CA = Cfg::KGS ? Slab2Lay{n * n * n, HCn, 1, n * n, n} : Tab::CA();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw) + 2 * warp;
  Number   *W   = reinterpret_cast<Number *>(smem_raw + 16 * Cfg::WPB) + (size_t)warp * Cfg::PER_WARP;  // coefficient image
  Number   *P   = W + Cfg::CWF;
  Number   *Q   = Cfg::NBUF == 2 ? P + Cfg::F : P;
  const Slab2Lane lm = slab2_lane<n>(lane);
  const bool active = lm.c >= 0;
  // idle lanes shadow the first lane of their own half warp for loads (broadcast, no conflict) and never store
  const int cl = lm.cl, ch = lm.ch, x = lm.x;
  const uint32_t total_warps = gridDim.x * Cfg::WPB;
  // work list: groups glist[0 .. n_groups) (multi-GPU: interface groups first, interior groups while the exchange
  // runs) or simply 0 .. n_groups
  // a launch that follows with programmatic stream serialization (the interior groups after the interface groups of a
  // multi-GPU apply) may start right away: it reads nothing this one writes
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t k0 = blockIdx.x * Cfg::WPB + warp;
  if (k0 >= n_groups) return;
  const uint32_t g0 = glist ? __ldg(glist + k0) : k0;
#ifdef MFG_SLAB2_ABLATE
  if (g_slab2_delay[0] > 0)
    {
      const int mode = g_slab2_delay[1];
      const unsigned k = mode == 0 ? (blockIdx.x / 148) % Cfg::MINB : mode == 1 ? warp % 2 : mode == 2 ? (blockIdx.x % Cfg::MINB) : warp;
      for (unsigned t = 0; t < k; ++t) __nanosleep(g_slab2_delay[0]);
    }
#endif

  if (lane == 0)
    {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncwarp();
  if (lane == 0) bulk_load(W, cwP + (size_t)g0 * Cfg::CWF, Cfg::CW_BYTES, bar);
  unsigned phase = 0;

  // per-lane base addresses: layout XY written in X, read in Y
  const int cAB = AB.SL * cl + AB.SH * ch, cBC = BC.SL * cl + BC.SH * ch, cCA = CA.SL * cl + CA.SH * ch;
  const int bABw = cAB + AB.SI * x, bABr = cAB + AB.SK * x;  // A: x = i ; B: x = k
  const int bBCw = cBC + BC.SK * x, bBCr = cBC + BC.SJ * x;  // B: x = k ; C: x = j
  const int bCAw = cCA + CA.SJ * x, bCAr = cCA + CA.SI * x;  // C: x = j ; A: x = i

  auto load_ids = [&](uint32_t g, uint32_t (&id)[NS]) {
    const uint32_t *row = idxP + (size_t)g * NS * 32 + lane;
#pragma unroll
    for (int s = 0; s < NS; ++s) id[s] = __ldg(row + 32 * s);
  };
  // read_dof_values (fee_gpu.cuh:323-338): every lane gathers its own slab, u[j + n k]
  auto gather = [&](const uint32_t (&id)[NS], Number (&u)[NS]) {
#pragma unroll
    for (int s = 0; s < NS; ++s)
      {
        if (Cfg::ABL & 1) u[s] = Number(id[s] & 0xffu);
        else if (Cfg::TEX) u[s] = (id[s] & CONSTRAINED_BIT) ? Number(0) : tex_fetch<Number>(tex, id[s]);
        else u[s] = (id[s] & CONSTRAINED_BIT) ? Number(0) : __ldg(src + id[s]);
      }
  };
  uint32_t id[NS];   // PF >= 1: index rows of the group whose gather comes next
  Number   un[NS];   // PF == 2: gathered values of the next group
  if (Cfg::PF >= 1) load_ids(g0, id);
  if (Cfg::PF == 2) gather(id, un);

  // list entries k0, k0 + total_warps, ... (an atomic work counter instead of the fixed stride measured 27 % slower: the
  // 4 warps of a CTA lose the lines shared by 4 adjacent groups, profiles/r01_multigpu_overlap.txt)
  for (uint32_t k = k0; k < n_groups; k += total_warps)
    {
      const uint32_t  g    = glist ? __ldg(glist + k) : k;
      const bool      more = k + total_warps < n_groups;
      const uint32_t  gn   = more ? (glist ? __ldg(glist + k + total_warps) : k + total_warps) : 0;
      if (!Cfg::KGS && more && lane < NS) asm volatile("prefetch.global.L2 [%0];" ::"l"(idxP + ((size_t)gn * NS + lane) * 32));
      const uint32_t *irow = idxP + (size_t)g * NS * 32 + lane;
      Number u[NS], r[NS];
      if (Cfg::KGS)
        {
          // ---- K: one cell per pass, lane t <-> (i, j) = (t % n, t / n), k in registers: gather, N_z, store for B ----
          // (all index rows first, then all gathers: two exposed memory latencies per group, not two per cell)
          const bool pl = lane < NS;
          constexpr int HC = Cfg::CW % 2 == 0 ? Cfg::CW / 2 : Cfg::CW;
          uint32_t kid[Cfg::CW][n];
          Number   kv[Cfg::CW][n];
#pragma unroll
          for (int c = 0; c < Cfg::CW; ++c)
            {
              const uint32_t  cell = g * Cfg::CW + c;
              const bool      ok = pl && cell < n_cells;
              const uint32_t *row = idxLex + (size_t)cell * Cfg::NPC + lane;
#pragma unroll
              for (int k = 0; k < n; ++k) kid[c][k] = ok ? __ldg(row + NS * k) : CONSTRAINED_BIT;
            }
#pragma unroll
          for (int c = 0; c < Cfg::CW; ++c)
#pragma unroll
            for (int k = 0; k < n; ++k) kv[c][k] = (kid[c][k] & CONSTRAINED_BIT) ? Number(0) : __ldg(src + kid[c][k]);
#pragma unroll
          for (int c = 0; c < Cfg::CW; ++c)
            {
              Number out[n];
              eo_apply<n, false, Number, NOPC>(em.N, kv[c], out);
              if (pl)
                {
                  Number *pc = P + AB.SL * (c % HC) + AB.SH * (c / HC) + lane;  // + i + n j = + lane
#pragma unroll
                  for (int k = 0; k < n; ++k) pc[AB.SK * k] = out[k];
                }
            }
        }
      else
        {
          if (Cfg::PF == 0) load_ids(g, id);
          if (Cfg::PF == 2)
            {
#pragma unroll
              for (int s = 0; s < NS; ++s) u[s] = un[s];
            }
          else gather(id, u);
          // ---- A: N_y, N_z ----
          slab2_apply<n, 1, n, false, Number, NOPC>(em.N, u);
          slab2_apply<n, n, 1, false, Number, NOPC>(em.N, u);
          if (active)
            {
#pragma unroll
              for (int k = 0; k < n; ++k)
#pragma unroll
                for (int j = 0; j < n; ++j) P[bABw + AB.SJ * j + AB.SK * k] = u[j + n * k];
            }
        }
      __syncwarp();
      // ---- B: N_x -> u at the quadrature points, u[i + n j] ----
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) u[i + n * j] = P[bABr + AB.SI * i + AB.SJ * j];
      __syncwarp();  // P consumed
      slab2_apply<n, 1, n, false, Number, NOPC>(em.N, u);
      if (Cfg::KGS) slab2_apply<n, n, 1, false, Number, NOPC>(em.N, u);  // N_y (the plane gather only did N_z)
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) Q[bBCw + BC.SI * i + BC.SJ * j] = u[i + n * j];
        }
      if (Cfg::NBUF == 1) __syncwarp();
      mbar_wait(bar, phase);  // coefficient image of this group has landed
      phase ^= 1;
      // quadrature phases x and y: r = D_x^T (w .* D_x u) + D_y^T (w .* D_y u)
#pragma unroll
      for (int j = 0; j < n; ++j)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int i = 0; i < n; ++i) in[i] = u[i + n * j];
          eo_apply<n, true, Number, NOPC>(em.D, in, gq);
#pragma unroll
          for (int i = 0; i < n; ++i) gq[i] *= W[bBCw + BC.SI * i + BC.SJ * j];
          eo_apply<n, true, Number, NOPC>(em.DT, gq, t);
#pragma unroll
          for (int i = 0; i < n; ++i) r[i + n * j] = t[i];
        }
#pragma unroll
      for (int i = 0; i < n; ++i)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int j = 0; j < n; ++j) in[j] = u[i + n * j];
          eo_apply<n, true, Number, NOPC>(em.D, in, gq);
#pragma unroll
          for (int j = 0; j < n; ++j) gq[j] *= W[bBCw + BC.SI * i + BC.SJ * j];
          eo_apply<n, true, Number, NOPC>(em.DT, gq, t);
#pragma unroll
          for (int j = 0; j < n; ++j) r[i + n * j] += t[j];
        }
      if (Cfg::NBUF == 2)
        {
          if (active)
            {
#pragma unroll
              for (int j = 0; j < n; ++j)
#pragma unroll
                for (int i = 0; i < n; ++i) P[bBCw + BC.SI * i + BC.SJ * j] = r[i + n * j];
            }
          __syncwarp();
        }
      // ---- C: quadrature phase z on u[i + n k] ----
#pragma unroll
      for (int k = 0; k < n; ++k)
#pragma unroll
        for (int i = 0; i < n; ++i) u[i + n * k] = Q[bBCr + BC.SI * i + BC.SK * k];
      if (Cfg::PF >= 1)
        {
          if (more) load_ids(gn, id);
          else
            {
#pragma unroll
              for (int s = 0; s < NS; ++s) id[s] = CONSTRAINED_BIT;
            }
        }
      if (Cfg::NBUF == 1)
        {
          __syncwarp();  // u consumed by every lane: the buffer now carries r
          if (active)
            {
#pragma unroll
              for (int j = 0; j < n; ++j)
#pragma unroll
                for (int i = 0; i < n; ++i) P[bBCw + BC.SI * i + BC.SJ * j] = r[i + n * j];
            }
          __syncwarp();
        }
#pragma unroll
      for (int i = 0; i < n; ++i)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int k = 0; k < n; ++k) in[k] = u[i + n * k];
          eo_apply<n, true, Number, NOPC>(em.D, in, gq);
#pragma unroll
          for (int k = 0; k < n; ++k) gq[k] *= W[bBCr + BC.SI * i + BC.SK * k];
          eo_apply<n, true, Number, NOPC>(em.DT, gq, t);
#pragma unroll
          for (int k = 0; k < n; ++k) u[i + n * k] = t[k] + P[bBCr + BC.SI * i + BC.SK * k];
        }
      __syncwarp();  // P, Q and the coefficient image are consumed
      if (more && lane == 0) bulk_load(W, cwP + (size_t)gn * Cfg::CWF, Cfg::CW_BYTES, bar);
      // ---- C: N_x^T, N_z^T ----
      slab2_apply<n, 1, n, false, Number, NOPC>(em.NT, u);
      slab2_apply<n, n, 1, false, Number, NOPC>(em.NT, u);
      if (active)
        {
#pragma unroll
          for (int k = 0; k < n; ++k)
#pragma unroll
            for (int i = 0; i < n; ++i) Q[bCAw + CA.SI * i + CA.SK * k] = u[i + n * k];
        }
      __syncwarp();
      if (Cfg::KGS)
        {
          // ---- K': one cell per pass, lane t <-> (i, k) = (t % n, t / n), j in registers: N_y^T, red.add ----
          const bool pl = lane < NS;
          constexpr int HC = Cfg::CW % 2 == 0 ? Cfg::CW / 2 : Cfg::CW;
          uint32_t jid[Cfg::CW][n];
#pragma unroll
          for (int c = 0; c < Cfg::CW; ++c)
            {
              const uint32_t  cell = g * Cfg::CW + c;
              const bool      ok = pl && cell < n_cells;
              const uint32_t *row = idxJ + (size_t)cell * Cfg::NPC + lane;
#pragma unroll
              for (int j = 0; j < n; ++j) jid[c][j] = ok ? __ldg(row + NS * j) : CONSTRAINED_BIT;
            }
          if (dep_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
          for (int c = 0; c < Cfg::CW; ++c)
            {
              const Number *qc = Q + CA.SL * (c % HC) + CA.SH * (c / HC) + (pl ? lane : 0);  // + i + n k = + lane
              Number in[n], out[n];
#pragma unroll
              for (int j = 0; j < n; ++j) in[j] = qc[CA.SJ * j];
              eo_apply<n, false, Number, NOPC>(em.NT, in, out);
#pragma unroll
              for (int j = 0; j < n; ++j)
                {
                  if (Cfg::ABL & 2)
                    {
                      if (out[j] == Number(12345.678)) red_add(dst + jid[c][j], out[j]);
                    }
                  else if (!(jid[c][j] & CONSTRAINED_BIT)) red_add(dst + jid[c][j], out[j]);
                }
            }
          if (Cfg::NBUF == 1) __syncwarp();  // the next group's first store goes to the same memory
        }
      else
        {
      // ---- A: N_y^T ----
#pragma unroll
      for (int k = 0; k < n; ++k)
#pragma unroll
        for (int j = 0; j < n; ++j) u[j + n * k] = Q[bCAr + CA.SJ * j + CA.SK * k];
      if (Cfg::NBUF == 1) __syncwarp();  // the next group's first store goes to the same memory
      uint32_t idc[NS];  // index rows of this group for the scatter
      if (Cfg::PF >= 1)
        {
#pragma unroll
          for (int s = 0; s < NS; ++s) idc[s] = __ldg(irow + 32 * s);
        }
      slab2_apply<n, 1, n, false, Number, NOPC>(em.NT, u);
      // ---- face merge inside the group: x (lanes i = n-1 -> i = 0 of cell c+1), y (slots j = n-1 -> j = 0 of cell c+2),
      //      z (slots k = n-1 -> k = 0 of cell c+4); a handed-over value is zeroed so that a later merge does not move it twice
      bool xs = false, ys = false, zs = false;  // this lane hands over its i = n-1 entries / its j = n-1 slots / its k = n-1 slots
      if (Cfg::MERGE && Cfg::CW <= SLAB2_MERGE_MAX_CW)
        {
          const uint32_t mm = __ldg(mergeP + g);
          if (mm != 0)
            {
              const int c = lm.c;
              xs = active && x == n - 1 && ((mm >> c) & 1u);
              ys = active && ((mm >> (10 + c)) & 1u);
              zs = active && ((mm >> (20 + c)) & 1u);
              const bool xd = active && x == 0 && c >= 1 && ((mm >> (c - 1)) & 1u);
              const bool yd = active && c >= 2 && ((mm >> (10 + c - 2)) & 1u);
              const bool zd = active && c >= 4 && ((mm >> (20 + c - 4)) & 1u);
              const int  lx = slab2_lane_of<n>(c - 1, n - 1), ly = slab2_lane_of<n>(c - 2, x), lz = slab2_lane_of<n>(c - 4, x);
              if (mm & 0x3ffu)
                {
#pragma unroll
                  for (int s = 0; s < NS; ++s)
                    {
                      const Number t = __shfl_sync(0xffffffffu, u[s], lx);
                      u[s] = xd ? u[s] + t : (xs ? Number(0) : u[s]);
                    }
                }
              if (mm & (0x3ffu << 10))
                {
#pragma unroll
                  for (int k = 0; k < n; ++k)
                    {
                      const Number t = __shfl_sync(0xffffffffu, u[(n - 1) + n * k], ly);
                      if (yd) u[n * k] += t;
                      if (ys) u[(n - 1) + n * k] = Number(0);
                    }
                }
              if (mm & (0x3ffu << 20))
                {
#pragma unroll
                  for (int j = 0; j < n; ++j)
                    {
                      const Number t = __shfl_sync(0xffffffffu, u[j + n * (n - 1)], lz);
                      if (zd) u[j] += t;
                      if (zs) u[j + n * (n - 1)] = Number(0);
                    }
                }
            }
        }
      if (Cfg::PF == 2 && more) gather(id, un);
      // launched as programmatic dependent of the kernel that zeroes dst (single-GPU apply): everything above overlapped
      // with its tail, the first red has to wait for it
      if (dep_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
      // ---- distribute_local_to_global (fee_gpu.cuh:346-365): red.add straight from registers ----
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          const uint32_t ii = Cfg::PF >= 1 ? idc[s] : __ldg(irow + 32 * s);
          if (Cfg::ABL & 2)
            {
              if (u[s] == Number(12345.678)) red_add(dst + ii, u[s]);
            }
          else
            {
              const bool handed_over = xs || (s % n == n - 1 && ys) || (s / n == n - 1 && zs);
              if (!(ii & CONSTRAINED_BIT) && !handed_over) red_add(dst + ii, u[s]);
            }
        }
        }  // !KGS
    }
}

template <typename Number>
void launch_laplace_slab2(int degree, int cfg, const uint32_t *idxP, const Number *cwP, const Number *src, Number *dst, uint32_t n_groups,
                          const double *N, const double *D, int sm_count, cudaStream_t stream, cudaTextureObject_t tex = 0,
                          const uint32_t *mergeP = nullptr, const uint32_t *glist = nullptr, bool pdl = false, bool dep_wait = false,
                          const uint32_t *idxLex = nullptr, const uint32_t *idxJ = nullptr, uint32_t n_cells = 0);
// layout of the kernel's private arrays (for the builders in operators.cu)
struct Slab2Geom { int n, cw, hc, cwf; Slab2Lay bc; };
bool      slab2_supported(int dim, int degree, mfg_dtype dt);
Slab2Geom slab2_geom(int degree, mfg_dtype dt);

}  // namespace mfg
