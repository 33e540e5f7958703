// kernels_stage.cuh -- staged Laplace cell kernel for 3D, n = p+1 in 3..6 (round 2; default for 3D degree 2..5).
//
// Same operator as the other cell kernels (the reference's apply_kernel_shmem<LocalOperator>, matrix_free_gpu.h:318-341
// + fee_gpu.cuh:197-365 + tensor_ops.cuh:179-261) and the same contraction core as kernels_slab2.cuh (a lane owns an
// n x n slab, layouts A -> B -> C -> A, even-odd 1-D contractions, merged weight a J^-2 JxW from a bulk-async copy).
// What changes is how the DoF values enter and leave the warp (ncu of the slab2 kernel, profiles/r01_*: 128 of its 220
// L1 wavefronts per cell and a third of its stall cycles were the per-entry gather and the per-entry red.add):
//   read_dof_values (fee_gpu.cuh:323-338)
//     The cells of a warp group own one contiguous range of the vector (deal.II numbers the DoFs a cell touches first
//     consecutively); that range and a short list of halo DoFs are copied with cp.async -- coalesced, no registers, one
//     group AHEAD of the arithmetic -- into a staging buffer, from which every lane reads its slab through a position
//     table shared by all groups of the same shape (stage_plan.cu).  Constrained DoFs read the zero slot.
//   distribute_local_to_global (fee_gpu.cuh:346-365, atomic.cuh:11-32)
//     Faces shared by two cells of the group are summed in registers as soon as the contraction ACROSS the face is
//     done (x and z in layout C, y in layout A: five shuffled values per lane and direction; the remaining
//     contractions act along the face and are the same for both cells).  Every DoF of the group then has one holder,
//     which writes it to the staging buffer; the warp writes the buffer out in memory order: plain coalesced stores
//     for DoFs no cell outside the group touches, red.add for the rest.
#pragma once
#include "slab_common.cuh"
#include "stage_plan.h"

namespace mfg {

template <int n, typename Number> struct StageCfg
{
  static constexpr int WB  = (int)sizeof(Number);
  using Tab = Slab2Tab<n, WB>;
  static constexpr int CW  = 32 / n;
  static constexpr int NS  = n * n, NS2 = (NS + 1) / 2;
  static constexpr int WPB = 4;
  static constexpr int F   = Tab::F;                  // elements of a transpose buffer / of the coefficient image
  static constexpr int PW  = (F * WB + 127) / 128 * 128 / WB;  // P and W buffers (128-byte multiples)
  static constexpr int XCAP = n == 3 ? 256 : n == 4 ? 480 : n == 5 ? 704 : 1024;  // staging slots; the last one is the zero / trash slot
  static constexpr int HREG = n == 6 ? 12 : 8;  // halo index registers per lane
  static constexpr int HMAX = 32 * HREG;
  static constexpr int OCAP = (CW * n * n * n + 31) / 32 * 32;  // capacity of the own range of a group
  static constexpr int LCAP = OCAP + HMAX;                      // capacity of the load list (own range, then halo)
  static constexpr int OBATCH = (CW * (n - 1) * (n - 1) * (n - 1) + 31) / 32;  // rows of 32 own entries handled at a time (a regular group owns CW (n-1)^3)
  // pattern table (uint16 units): header | pos, two rows per uint32 [NS2][32] | own range: slot | flag << 16 [OCAP] | halo slots [HMAX]
  static constexpr int PSTRIDE = (STAGE_PH + 2 * NS2 * 32 + 2 * OCAP + HMAX + 7) / 8 * 8;
  static constexpr int NCLASS = n == 5 ? 4 : n == 4 ? 1 : n == 6 ? 8 : 4;  // groups g and g + NCLASS start at the same place of an octet
  static constexpr int PER_WARP = XCAP + 2 * PW;      // X (staging in), P (transposes + staging out), W (coefficients)
  static constexpr size_t SMEM = (size_t)WPB * PER_WARP * WB + 16 * WPB + 2 * PSTRIDE;
  static constexpr int MINB = (SMEM + 1024) * 3 <= 227 * 1024 ? 3 : 2;
  static constexpr uint32_t CW_BYTES = F * WB;
  static_assert(XCAP <= PW, "the results are staged in P");
};

struct StageClasses { uint32_t pat[8]; };

template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void stage_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// mode: 0 = every group (class order: CTA b takes the groups g = NCLASS m + b % NCLASS, whose position tables are the ones
//           the CTA keeps in shared memory), 2 = the same without the groups flagged as interface groups (multi-GPU, the
//           part that overlaps the exchange), 1 = the groups of the work list glist (multi-GPU: the interface groups)
// add:  vmult_add -- the plain stores become additions
// SYNC: the warps of a CTA start every work item together, so that they run through the same part of the (long) loop body
// at the same time and share its instruction-cache lines
template <int n, typename Number, bool SYNC>
__global__ void __launch_bounds__(StageCfg<n, Number>::WPB * 32, StageCfg<n, Number>::MINB)
laplace_cell_stage(const uint4 *__restrict__ gdesc, const uint32_t *__restrict__ halo, const uint16_t *__restrict__ ptab, const Number *__restrict__ cwP,
                   const Number *__restrict__ src, Number *__restrict__ dst, const uint32_t n_groups, const __grid_constant__ EoMats<Number, n> em,
                   const uint32_t *__restrict__ glist, const uint32_t n_list, const int mode, const __grid_constant__ StageClasses cls,
                   const int dep_wait, const int add)
{
  using Cfg = StageCfg<n, Number>;
  using Tab = typename Cfg::Tab;
  constexpr int NS = Cfg::NS, NS2 = Cfg::NS2, XCAP = Cfg::XCAP, WB = Cfg::WB, PSTRIDE = Cfg::PSTRIDE;
  constexpr Slab2Lay AB = Tab::AB(), BC = Tab::BC(), CA = Tab::CA();
  extern __shared__ __align__(16) unsigned char smem_raw[];  // (dynamic shared memory starts 1024-byte aligned)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Number   *X   = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;  // staging in
  Number   *P   = X + XCAP;                                                             // transposes, staging out
  Number   *W   = P + Cfg::PW;                                                          // coefficient image
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)Cfg::WPB * Cfg::PER_WARP * WB) + 2 * warp;
  uint16_t *cache = reinterpret_cast<uint16_t *>(smem_raw + (size_t)Cfg::WPB * Cfg::PER_WARP * WB + 16 * Cfg::WPB);
  unsigned char *const Xb = reinterpret_cast<unsigned char *>(X), *const Pb = reinterpret_cast<unsigned char *>(P);
  const Slab2Lane lm = slab2_lane<n>(lane);
  const bool active = lm.c >= 0;
  const int  cl = lm.cl, ch = lm.ch, x = lm.x, cc = lm.c;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- the position tables of the CTA's class of groups stay in shared memory ----
  const uint32_t a_cls = mode == 1 ? 0u : blockIdx.x % Cfg::NCLASS;
  const uint32_t cached_pat = cls.pat[a_cls];
  if (cached_pat != STAGE_NOPAT)
    {
      const uint4 *g4 = reinterpret_cast<const uint4 *>(ptab + (size_t)cached_pat * PSTRIDE);
      uint4       *s4 = reinterpret_cast<uint4 *>(cache);
      for (int t = threadIdx.x; t < PSTRIDE / 8; t += Cfg::WPB * 32) s4[t] = __ldg(g4 + t);
    }
  __syncthreads();
  auto tables = [&](uint32_t pat) -> const uint16_t * { return pat == cached_pat ? cache : ptab + (size_t)pat * PSTRIDE; };

  // work items of this warp: t0, t0 + stride, ... < limit; item -> group
  const uint32_t stride = mode == 1 ? gridDim.x * Cfg::WPB : (gridDim.x / Cfg::NCLASS) * Cfg::WPB;
  const uint32_t t0     = mode == 1 ? blockIdx.x * Cfg::WPB + warp : (blockIdx.x / Cfg::NCLASS) * Cfg::WPB + warp;
  const uint32_t limit  = mode == 1 ? n_list : (n_groups + Cfg::NCLASS - 1 - a_cls) / Cfg::NCLASS;
  auto group_of = [&](uint32_t t) { return mode == 1 ? __ldg(glist + t) : Cfg::NCLASS * t + a_cls; };
  auto fetch = [&](uint32_t t, uint4 &d) {
    if (t < limit) d = __ldg(gdesc + group_of(t));
    else d.z = STAGE_NOPAT << 16;
  };
  // (the plan leaves some groups to the slab2 kernel; mode 2 leaves out the interface groups)
  auto usable = [&](const uint4 &d) { return (d.z >> 16) != STAGE_NOPAT && !(mode == 2 && (d.w >> 31)); };

  // ---- asynchronous copy of the DoF values of a group into X (read_dof_values, first half) ----
  // (table entries are loaded in batches BEFORE the copies / stores that use them: the tables may sit in shared memory,
  // and the compiler cannot move a load across a shared-memory store it has to assume might alias)
  auto issue_own = [&](const uint4 &d, const uint16_t *ph) {
    const uint32_t  own_total = ph[STAGE_H_OWN];
    const uint32_t *own32 = reinterpret_cast<const uint32_t *>(ph + STAGE_PH) + NS2 * 32 + lane;
    for (uint32_t e0 = 0; e0 < own_total; e0 += 32 * Cfg::OBATCH)
      {
        uint32_t sl[Cfg::OBATCH];
#pragma unroll
        for (int t = 0; t < Cfg::OBATCH; ++t) sl[t] = e0 + 32 * t < own_total ? own32[e0 + 32 * t] : 0u;
#pragma unroll
        for (int t = 0; t < Cfg::OBATCH; ++t)
          if (e0 + 32 * t + lane < own_total) cp_async_elem<WB>(Xb + (sl[t] & 0xffffu), src + d.x + e0 + 32 * t + lane);
      }
  };
  auto load_halo_ids = [&](const uint4 &d, uint32_t (&hid)[Cfg::HREG]) {
    const uint32_t nh = d.z & 0xffffu;
#pragma unroll
    for (int t = 0; t < Cfg::HREG; ++t) hid[t] = 32 * t < nh ? __ldg(halo + d.y + 32 * t + lane) : 0u;  // (rows are padded)
  };
  auto issue_halo = [&](const uint4 &d, const uint16_t *ph, const uint32_t (&hid)[Cfg::HREG]) {
    const uint32_t  nh = d.z & 0xffffu;
    const uint16_t *hs = ph + STAGE_PH + 2 * NS2 * 32 + 2 * Cfg::OCAP + lane;
    uint32_t        sl[Cfg::HREG];
#pragma unroll
    for (int t = 0; t < Cfg::HREG; ++t) sl[t] = 32 * t < nh ? hs[32 * t] : 0u;
#pragma unroll
    for (int t = 0; t < Cfg::HREG; ++t)
      if (32 * t + lane < nh) cp_async_elem<WB>(Xb + sl[t], src + hid[t]);
    stage_cp_commit();
  };

  if (t0 >= limit && !SYNC) return;
  if (lane == 0)
    {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      X[XCAP - 1] = Number(0);  // zero slot: never written by a copy
    }
  __syncwarp();
  unsigned phase = 0;
  // items of the CTA's first warp bound the trip count (SYNC: every warp runs the same number of iterations)
  const uint32_t t00 = mode == 1 ? blockIdx.x * Cfg::WPB : (blockIdx.x / Cfg::NCLASS) * Cfg::WPB;
  const int      n_iter = t00 < limit ? (int)((limit - t00 + stride - 1) / stride) : 0;
  uint4 d, dn, dq;
  d.z = STAGE_NOPAT << 16;
  fetch(t0, dn);

  const int cAB = AB.SL * cl + AB.SH * ch, cBC = BC.SL * cl + BC.SH * ch, cCA = CA.SL * cl + CA.SH * ch;
  const int bABw = cAB + AB.SI * x, bABr = cAB + AB.SK * x;
  const int bBCw = cBC + BC.SK * x, bBCr = cBC + BC.SJ * x;
  const int bCAw = cCA + CA.SJ * x, bCAr = cCA + CA.SI * x;

  // iteration it works on item t0 + it stride and requests the data of the item after it (it = -1 only requests)
  for (int it = -1; it < n_iter; ++it)
    {
      if (SYNC) __syncthreads();
      const uint32_t  tn = t0 + (uint32_t)(it + 1) * stride;
      const bool      cur = it >= 0 && usable(d), more = usable(dn);
      const uint16_t *ph = tables(d.z >> 16), *phn = tables(dn.z >> 16);
      const uint32_t *pp = reinterpret_cast<const uint32_t *>(ph + STAGE_PH) + lane;
      fetch(tn + stride, dq);  // descriptor of the item after the next: needed one iteration from now
      Number u[NS], r[NS];
      // ---- read_dof_values, second half: every lane reads its slab u[j + n k] from the staging buffer ----
      if (cur)
        {
          uint32_t pz[NS2];
#pragma unroll
          for (int s = 0; s < NS2; ++s) pz[s] = pp[32 * s];
          stage_cp_wait_all();
          __syncwarp();
#pragma unroll
          for (int s = 0; s < NS; ++s) u[s] = *reinterpret_cast<const Number *>(Xb + ((pz[s / 2] >> (16 * (s % 2))) & 0x7fffu));
          __syncwarp();  // X is free: the copy of the next group may overwrite it
        }
      uint32_t hid[Cfg::HREG];
      if (more)
        {
          load_halo_ids(dn, hid);
          issue_own(dn, phn);
        }
      // ---- A: N_y, N_z ----
      if (cur)
        {
          slab2_apply<n, 1, n, false, Number>(em.N, u);
          slab2_apply<n, n, 1, false, Number>(em.N, u);
        }
      if (more) issue_halo(dn, phn, hid);
      if (!cur)
        {
          if (more && lane == 0) bulk_load(W, cwP + (size_t)group_of(tn) * Cfg::F, Cfg::CW_BYTES, bar);
          d = dn; dn = dq;
          continue;
        }
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
          for (int i = 0; i < n; ++i) gq[i] *= W[bBCw + BC.SI * i + BC.SJ * j];
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
          for (int j = 0; j < n; ++j) gq[j] *= W[bBCw + BC.SI * i + BC.SJ * j];
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
          for (int kk = 0; kk < n; ++kk) gq[kk] *= W[bBCr + BC.SI * i + BC.SK * kk];
          eo_apply<n, true, Number>(em.DT, gq, t);
#pragma unroll
          for (int kk = 0; kk < n; ++kk) u[i + n * kk] = t[kk] + P[bBCr + BC.SI * i + BC.SK * kk];
        }
      __syncwarp();  // P and the coefficient image are consumed
      if (more && lane == 0) bulk_load(W, cwP + (size_t)group_of(tn) * Cfg::F, Cfg::CW_BYTES, bar);
      // ---- face merges: bit (10 dir + c) of the mask = cell c hands its upper face to cell c + 2^dir of the group ----
      const uint32_t mm = d.w & 0x3fffffffu;
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
      __syncwarp();  // P consumed: it now takes the results of the group
      // (the halo indices of the write-out are requested before the last contraction, not when they are needed)
      const uint32_t nh = d.z & 0xffffu;
      uint32_t wid[Cfg::HREG];
#pragma unroll
      for (int t = 0; t < Cfg::HREG; ++t) wid[t] = 32 * t < nh ? __ldg(halo + d.y + 32 * t + lane) : 0u;
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
      // ---- distribute_local_to_global: one holder per DoF writes to the staging buffer ... ----
      {
        uint32_t pw[NS2];
#pragma unroll
        for (int s = 0; s < NS2; ++s) pw[s] = pp[32 * s];
#pragma unroll
        for (int s = 0; s < NS; ++s)
          {
            const uint32_t pz = (pw[s / 2] >> (16 * (s % 2))) & 0xffffu;
            if (!(pz & STAGE_DEAD)) *reinterpret_cast<Number *>(Pb + pz) = u[s];
          }
      }
      __syncwarp();
      if (dep_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
      // ---- ... and the warp writes it out in memory order ----
      {
        const uint32_t  own_total = ph[STAGE_H_OWN];
        const uint32_t *own32 = reinterpret_cast<const uint32_t *>(ph + STAGE_PH) + NS2 * 32 + lane;
        const uint16_t *hs = ph + STAGE_PH + 2 * NS2 * 32 + 2 * Cfg::OCAP + lane;
        uint32_t hsl[Cfg::HREG];
#pragma unroll
        for (int t = 0; t < Cfg::HREG; ++t) hsl[t] = 32 * t < nh ? hs[32 * t] : 0u;
        for (uint32_t e0 = 0; e0 < own_total; e0 += 32 * Cfg::OBATCH)
          {
            uint32_t sl[Cfg::OBATCH];
#pragma unroll
            for (int t = 0; t < Cfg::OBATCH; ++t) sl[t] = e0 + 32 * t < own_total ? own32[e0 + 32 * t] : 0u;
#pragma unroll
            for (int t = 0; t < Cfg::OBATCH; ++t)
              {
                const uint32_t e = e0 + 32 * t + lane, f = sl[t] >> 16;
                if (e < own_total && f != 0u)
                  {
                    const Number v = *reinterpret_cast<const Number *>(Pb + (sl[t] & 0xffffu));
                    if (f == 1u && !add) dst[d.x + e] = v;
                    else red_add(dst + d.x + e, v);
                  }
              }
          }
#pragma unroll
        for (int t = 0; t < Cfg::HREG; ++t)
          if (32 * t + lane < nh) red_add(dst + wid[t], *reinterpret_cast<const Number *>(Pb + hsl[t]));
      }
      __syncwarp();  // the next group's transposes reuse P
      d = dn; dn = dq;
    }
}

struct StageGeom { int n, cw, hc, xcap, hmax, ocap, lcap, pstride, nclass; };
bool      stage_supported(int dim, int degree, mfg_dtype dt);
StageGeom stage_geom(int degree, mfg_dtype dt);
// mode 0 / 2: all groups / all but the interface groups in class order; mode 1: the n_list groups of glist
template <typename Number>
void launch_laplace_stage(int degree, const uint32_t *gdesc, const uint32_t *halo, const uint16_t *ptab, int pstride, const uint32_t *class_pat,
                          const Number *cwP, const Number *src, Number *dst, uint32_t n_groups, const double *N, const double *D, int sm_count,
                          cudaStream_t stream, const uint32_t *glist, uint32_t n_list, int mode, bool pdl, bool dep_wait, bool add, int device, bool sync = true);

}  // namespace mfg
