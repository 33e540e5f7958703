// kernels_slab3.cuh -- third-generation slab Laplace cell kernel for 3D, n = p+1 <= 6 (round 2).
//
// Same operator and contraction core as kernels_slab2.cuh (apply_kernel_shmem<LocalOperator>, matrix_free_gpu.h:318-341
// + fee_gpu.cuh:197-365 + tensor_ops.cuh:179-261: a lane owns an n x n slab, layouts A -> B -> C -> A, even-odd 1-D
// contractions, merged weight a J^-2 JxW delivered by one bulk-async copy per group).  Two changes, both read off
// the ncu profiles of the slab2 kernel (profiles/r01_slab2_kernel_cfg3_q4_f64_r6_ncu.txt: a third of the stall cycles
// wait for the gather) and of the staged kernel (profiles/r02_stage_*: staging through tables costs 40 % more
// instructions than it saves in memory transactions):
//   read_dof_values (fee_gpu.cuh:323-338)  every lane still gathers the DoFs of its own slab through the coalesced
//     index rows idxP, but with cp.async into a row of shared memory of its own (element s of lane l at 32 s + l: no
//     bank conflicts, immediate offsets, no other lane involved), ONE GROUP AHEAD of the arithmetic: the index rows of
//     group g+1 are requested after the slab of group g has been read, the copies are issued after the first two
//     contractions of group g, and nothing waits for global memory any more.  Constrained DoFs are zero-filled by the copy.
//   distribute_local_to_global (fee_gpu.cuh:346-365, atomic.cuh:11-32)  faces shared by two cells of the group are
//     summed in registers as soon as the contraction ACROSS the face is done (x and z in layout C, y in layout A: n
//     shuffled values per lane and direction instead of n^2 for the x face at the end; the remaining contractions act
//     along the face and are the same for both cells), which removes a fifth of the red.global.add sectors.
#pragma once
#include "slab_common.cuh"

namespace mfg {

template <int n, typename Number, bool ASYNC> struct Slab3Cfg
{
  static constexpr int WB  = (int)sizeof(Number);
  using Tab = Slab2Tab<n, WB>;
  static constexpr int CW  = 32 / n;
  static constexpr int NS  = n * n;
  static constexpr int WPB = 4;
  static constexpr int F   = Tab::F;        // elements of the transpose buffer / of the coefficient image
  static constexpr int G   = ASYNC ? NS * 32 : 0;  // gather rows
  static constexpr int PER_WARP = 2 * F + G;
  static constexpr size_t SMEM = 16 * WPB + (size_t)WPB * PER_WARP * WB;
  // 3 CTAs x 4 warps x 168 registers per SM; n = 6: 2 CTAs x 255 registers (slabs of 36 values).  More resident warps
  // (5- and 7-warp CTAs at 136 / 144 registers) spill 90-150 bytes and measured 15-25 % slower (profiles/r02_slab3_flavours.txt)
  static constexpr int REGS = n >= 6 ? 255 : 168;
  static constexpr uint32_t CW_BYTES = F * WB;
  static constexpr bool MERGE = CW <= SLAB2_MERGE_MAX_CW;
};

template <int BYTES> __device__ __forceinline__ void slab3_cp_zfill(void *smem_dst, const void *gsrc, bool valid)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int      sz = valid ? BYTES : 0;  // fewer source bytes than the copy size: the rest is filled with zeros
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(d), "l"(gsrc), "n"(BYTES), "r"(sz) : "memory");
}

// ASYNC: gather through cp.async one group ahead (false: register gather at the start of the group);
// EARLY: face merges right after the contraction across the face (false: no merges);
// DOT:   the kernel also emits src . (A src) (conjugate gradients: the denominator of alpha without a pass of its own):
//        src . (A src) = sum over the quadrature points of w |grad src|^2 (+ src_c^2 on the constrained rows), which is at hand in
//        registers in the quadrature phases: every warp sums its share in a fixed order and stores one partial sum; dot_out[w] is
//        written by the same warp w in every launch
template <int n, typename Number, bool ASYNC, bool EARLY, bool DOT>
__global__ void __launch_bounds__(Slab3Cfg<n, Number, ASYNC>::WPB * 32) __maxnreg__((Slab3Cfg<n, Number, ASYNC>::REGS))
laplace_cell_slab3(const uint32_t *__restrict__ idxP, const Number *__restrict__ cwP, const Number *__restrict__ src, Number *__restrict__ dst,
                   const uint32_t n_groups, const __grid_constant__ EoMats<Number, n> em, const uint32_t *__restrict__ mergeP,
                   const uint32_t *__restrict__ glist, const int dep_wait, const uint32_t *__restrict__ clist, const uint32_t n_clist,
                   double *__restrict__ dot_out)
{
  using Cfg = Slab3Cfg<n, Number, ASYNC>;
  using Tab = typename Cfg::Tab;
  constexpr int NS = Cfg::NS, WB = Cfg::WB;
  constexpr Slab2Lay AB = Tab::AB(), BC = Tab::BC(), CA = Tab::CA();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw) + 2 * warp;
  Number   *W   = reinterpret_cast<Number *>(smem_raw + 16 * Cfg::WPB) + (size_t)warp * Cfg::PER_WARP;  // coefficient image
  Number   *P   = W + Cfg::F;                                                                            // transposes
  Number   *Gr  = P + Cfg::F + lane;                                                                     // gather rows, this lane's column
  const Slab2Lane lm = slab2_lane<n>(lane);
  const bool active = lm.c >= 0;
  const int  cl = lm.cl, ch = lm.ch, x = lm.x, cc = lm.c;
  const uint32_t total_warps = gridDim.x * Cfg::WPB;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // dep_wait: 1 = the kernel in front only writes dst (zero pass): wait before the first write; 2 = it also writes src
  // (fused CG: d = beta d - z): wait before the first read
  if (dep_wait == 2) asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t k0 = blockIdx.x * Cfg::WPB + warp;
  if (k0 >= n_groups) return;
  auto group_of = [&](uint32_t k) { return glist ? __ldg(glist + k) : k; };
  auto load_ids = [&](uint32_t g, uint32_t (&id)[NS]) {
    const uint32_t *row = idxP + (size_t)g * NS * 32 + lane;
#pragma unroll
    for (int s = 0; s < NS; ++s) id[s] = __ldg(row + 32 * s);
  };
  auto issue_gather = [&](const uint32_t (&id)[NS]) {
#pragma unroll
    for (int s = 0; s < NS; ++s) slab3_cp_zfill<WB>(Gr + 32 * s, src + (id[s] & ~CONSTRAINED_BIT), !(id[s] & CONSTRAINED_BIT));
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  if (lane == 0)
    {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncwarp();
  uint32_t g = group_of(k0);
  if (lane == 0) bulk_load(W, cwP + (size_t)g * Cfg::F, Cfg::CW_BYTES, bar);
  unsigned phase = 0;
  double   dacc = 0;  // DOT: this lane's share of src . (A src)
  if (ASYNC)
    {
      uint32_t id[NS];
      load_ids(g, id);
      issue_gather(id);
    }

  const int cAB = AB.SL * cl + AB.SH * ch, cBC = BC.SL * cl + BC.SH * ch, cCA = CA.SL * cl + CA.SH * ch;
  const int bABw = cAB + AB.SI * x, bABr = cAB + AB.SK * x;
  const int bBCw = cBC + BC.SK * x, bBCr = cBC + BC.SJ * x;
  const int bCAw = cCA + CA.SJ * x, bCAr = cCA + CA.SI * x;

  for (uint32_t k = k0; k < n_groups; k += total_warps)
    {
      const bool     more = k + total_warps < n_groups;
      const uint32_t gn = more ? group_of(k + total_warps) : 0;
      const uint32_t *irow = idxP + (size_t)g * NS * 32 + lane;
      if (more && lane < NS) asm volatile("prefetch.global.L2 [%0];" ::"l"(idxP + ((size_t)(more ? gn : g) * NS + lane) * 32));
      Number u[NS], r[NS];
      Number eacc = Number(0);  // DOT: sum of w |grad src|^2 over this lane's quadrature points of the group
      // ---- read_dof_values: the slab u[j + n k] arrived in this lane's rows one group ago ----
      uint32_t id[NS];
      if (ASYNC)
        {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
          for (int s = 0; s < NS; ++s) u[s] = Gr[32 * s];
          // the index rows of the next group are requested now and used after the first two contractions
          if (more) load_ids(gn, id);
        }
      else
        {
          load_ids(g, id);
#pragma unroll
          for (int s = 0; s < NS; ++s) u[s] = (id[s] & CONSTRAINED_BIT) ? Number(0) : __ldg(src + id[s]);
        }
      // ---- A: N_y, N_z ----
      slab2_apply<n, 1, n, false, Number>(em.N, u);
      slab2_apply<n, n, 1, false, Number>(em.N, u);
      if (ASYNC && more) issue_gather(id);
      if (active)
        {
#pragma unroll
          for (int kk = 0; kk < n; ++kk)
#pragma unroll
            for (int j = 0; j < n; ++j) P[bABw + AB.SJ * j + AB.SK * kk] = u[j + n * kk];
        }
      __syncwarp();
      // ---- B: N_x -> u at the quadrature points, u[i + n j] ----
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) u[i + n * j] = P[bABr + AB.SI * i + AB.SJ * j];
      __syncwarp();  // P consumed
      slab2_apply<n, 1, n, false, Number>(em.N, u);
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) P[bBCw + BC.SI * i + BC.SJ * j] = u[i + n * j];
        }
      __syncwarp();
      mbar_wait(bar, phase);  // coefficient image of this group has landed
      phase ^= 1;
      // quadrature phases x and y: r = D_x^T (w .* D_x u) + D_y^T (w .* D_y u)
#pragma unroll
      for (int j = 0; j < n; ++j)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int i = 0; i < n; ++i) in[i] = u[i + n * j];
          eo_apply<n, true, Number>(em.D, in, gq);
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              const Number wg = gq[i] * W[bBCw + BC.SI * i + BC.SJ * j];
              if (DOT) eacc = fma(gq[i], wg, eacc);
              gq[i] = wg;
            }
          eo_apply<n, true, Number>(em.DT, gq, t);
#pragma unroll
          for (int i = 0; i < n; ++i) r[i + n * j] = t[i];
        }
#pragma unroll
      for (int i = 0; i < n; ++i)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int j = 0; j < n; ++j) in[j] = u[i + n * j];
          eo_apply<n, true, Number>(em.D, in, gq);
#pragma unroll
          for (int j = 0; j < n; ++j)
            {
              const Number wg = gq[j] * W[bBCw + BC.SI * i + BC.SJ * j];
              if (DOT) eacc = fma(gq[j], wg, eacc);
              gq[j] = wg;
            }
          eo_apply<n, true, Number>(em.DT, gq, t);
#pragma unroll
          for (int j = 0; j < n; ++j) r[i + n * j] += t[j];
        }
      // ---- C: quadrature phase z on u[i + n k] ----
#pragma unroll
      for (int kk = 0; kk < n; ++kk)
#pragma unroll
        for (int i = 0; i < n; ++i) u[i + n * kk] = P[bBCr + BC.SI * i + BC.SK * kk];
      __syncwarp();  // u consumed by every lane: the buffer now carries r
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) P[bBCw + BC.SI * i + BC.SJ * j] = r[i + n * j];
        }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < n; ++i)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int kk = 0; kk < n; ++kk) in[kk] = u[i + n * kk];
          eo_apply<n, true, Number>(em.D, in, gq);
#pragma unroll
          for (int kk = 0; kk < n; ++kk)
            {
              const Number wg = gq[kk] * W[bBCr + BC.SI * i + BC.SK * kk];
              if (DOT) eacc = fma(gq[kk], wg, eacc);
              gq[kk] = wg;
            }
          eo_apply<n, true, Number>(em.DT, gq, t);
#pragma unroll
          for (int kk = 0; kk < n; ++kk) u[i + n * kk] = t[kk] + P[bBCr + BC.SI * i + BC.SK * kk];
        }
      __syncwarp();  // P and the coefficient image are consumed
      if (more && lane == 0) bulk_load(W, cwP + (size_t)gn * Cfg::F, Cfg::CW_BYTES, bar);
      // ---- face merges: bit (10 dir + c) of the mask = cell c hands its upper face to cell c + 2^dir of the group ----
      const uint32_t mm = (EARLY && Cfg::MERGE) ? __ldg(mergeP + g) : 0u;
      const bool xs = active && ((mm >> cc) & 1u), xd = active && cc >= 1 && ((mm >> (cc - 1)) & 1u);
      const bool zs = active && ((mm >> (20 + cc)) & 1u), zd = active && cc >= 4 && ((mm >> (20 + cc - 4)) & 1u);
      const bool ys = active && ((mm >> (10 + cc)) & 1u), yd = active && cc >= 2 && ((mm >> (10 + cc - 2)) & 1u);
      // ---- C: N_x^T, x merge (lane <-> j: entries i = n-1 of cell c go to i = 0 of cell c+1) ----
      slab2_apply<n, 1, n, false, Number>(em.NT, u);
      if (mm & 0x3ffu)
        {
          const int lx = slab2_lane_of<n>(cc - 1, x);
#pragma unroll
          for (int kk = 0; kk < n; ++kk)
            {
              const Number t = __shfl_sync(0xffffffffu, u[(n - 1) + n * kk], lx);
              if (xd) u[n * kk] += t;
              if (xs) u[(n - 1) + n * kk] = Number(0);
            }
        }
      // ---- C: N_z^T, z merge (entries k = n-1 of cell c go to k = 0 of cell c+4) ----
      slab2_apply<n, n, 1, false, Number>(em.NT, u);
      if (mm & (0x3ffu << 20))
        {
          const int lz = slab2_lane_of<n>(cc - 4, x);
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              const Number t = __shfl_sync(0xffffffffu, u[i + n * (n - 1)], lz);
              if (zd) u[i] += t;
              if (zs) u[i + n * (n - 1)] = Number(0);
            }
        }
      if (active)
        {
#pragma unroll
          for (int kk = 0; kk < n; ++kk)
#pragma unroll
            for (int i = 0; i < n; ++i) P[bCAw + CA.SI * i + CA.SK * kk] = u[i + n * kk];
        }
      __syncwarp();
      // ---- A: N_y^T, y merge (lane <-> i: entries j = n-1 of cell c go to j = 0 of cell c+2) ----
#pragma unroll
      for (int kk = 0; kk < n; ++kk)
#pragma unroll
        for (int j = 0; j < n; ++j) u[j + n * kk] = P[bCAr + CA.SJ * j + CA.SK * kk];
      __syncwarp();  // the next group's first store goes to the same memory
      uint32_t idc[NS];  // index rows of this group for the scatter: requested before the last contraction
#pragma unroll
      for (int s = 0; s < NS; ++s) idc[s] = __ldg(irow + 32 * s);
      slab2_apply<n, 1, n, false, Number>(em.NT, u);
      if (mm & (0x3ffu << 10))
        {
          const int ly = slab2_lane_of<n>(cc - 2, x);
#pragma unroll
          for (int kk = 0; kk < n; ++kk)
            {
              const Number t = __shfl_sync(0xffffffffu, u[(n - 1) + n * kk], ly);
              if (yd) u[n * kk] += t;
              if (ys) u[(n - 1) + n * kk] = Number(0);
            }
        }
      if (dep_wait == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
      // whole vmult: identity on the constrained rows (load_and_add_constrained_values, constraint_handler_gpu.cu:277-289):
      // dst is zero now, the cells never write these rows, so any warp may copy its share once
      if (clist != nullptr && k == k0)
        for (uint32_t t = k0 * 32 + lane; t < n_clist; t += (total_warps < n_groups ? total_warps : n_groups) * 32)  // (warps with work)
          {
            const uint32_t c = __ldg(clist + t);
            const Number   v = __ldg(src + c);
            dst[c] = v;
            if (DOT) dacc += (double)v * (double)v;
          }
      // ---- distribute_local_to_global: red.add straight from registers; what was handed over is not written ----
      const bool xdead = xs && x == n - 1;
#pragma unroll
      for (int s = 0; s < NS; ++s)
        {
          const bool handed_over = xdead || (s % n == n - 1 && ys) || (s / n == n - 1 && zs);
          if (!(idc[s] & CONSTRAINED_BIT) && !handed_over)
            red_add(dst + idc[s], u[s]);
        }
      if (DOT && active) dacc += (double)eacc;
      g = gn;
    }
  if (DOT)
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
      if (lane == 0) dot_out[k0] = dacc;
    }
}

template <typename Number>
void launch_laplace_slab3(int degree, const uint32_t *idxP, const Number *cwP, const Number *src, Number *dst, uint32_t n_groups, const double *N,
                          const double *D, int sm_count, cudaStream_t stream, const uint32_t *mergeP, const uint32_t *glist, bool pdl, int dep_wait,
                          int device, int flavour, const uint32_t *clist, uint32_t n_clist, double *dot_out = nullptr,
                          uint32_t *n_dot = nullptr);  // flavour: bit 0 = asynchronous gather, bit 1 = early face merges;
                                                       // dot_out: per-warp partial sums of src . (A src), *n_dot of them (register gather only)

}  // namespace mfg
