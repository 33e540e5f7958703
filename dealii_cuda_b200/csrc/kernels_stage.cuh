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
#include "kernels_slab2.cuh"
#include "stage_plan.h"

namespace mfg {

template <int n, typename Number> struct StageCfg
{
  static constexpr int WB  = (int)sizeof(Number);
  using Tab = Slab2Tab<n, WB>;
  static constexpr int CW  = 32 / n;
  static constexpr int NS  = n * n;
  static constexpr int WPB = 4;
  static constexpr int F   = Tab::F;                  // elements of a transpose buffer / of the coefficient image
  static constexpr int XCAP = (F + 31) / 32 * 32;     // staging slots; the last one is the zero / trash slot
  static constexpr int HREG = n == 5 ? 10 : n == 6 ? 12 : 8;  // halo index registers per lane
  static constexpr int HMAX = 32 * HREG;
  static constexpr int MINB = (3 * XCAP * WB * WPB + 64) * 3 <= 224 * 1024 ? 3 : 2;
  static constexpr int PER_WARP = 3 * XCAP;           // X (staging in), P (transposes + staging out), W (coefficients)
  static constexpr size_t SMEM = (size_t)WPB * PER_WARP * WB + 16 * WPB;
  static constexpr uint32_t CW_BYTES = F * WB;
  static constexpr int OCAP = (CW * n * n * n + 31) / 32 * 32;  // capacity of the own range of a group
  static constexpr int LCAP = OCAP + HMAX;                      // capacity of the load list (own range, then halo)
};

template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void stage_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void stage_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int n, typename Number>
__global__ void __launch_bounds__(StageCfg<n, Number>::WPB * 32, StageCfg<n, Number>::MINB)
laplace_cell_stage(const uint4 *__restrict__ gdesc, const uint32_t *__restrict__ halo, const uint16_t *__restrict__ ptab, const int pstride,
                   const Number *__restrict__ cwP, const Number *__restrict__ src, Number *__restrict__ dst, const uint32_t n_items,
                   const __grid_constant__ EoMats<Number, n> em, const uint32_t *__restrict__ glist, const int dep_wait)
{
  using Cfg = StageCfg<n, Number>;
  using Tab = typename Cfg::Tab;
  constexpr int NS = Cfg::NS, NS2 = (NS + 1) / 2, XCAP = Cfg::XCAP, WB = Cfg::WB;
  constexpr Slab2Lay AB = Tab::AB(), BC = Tab::BC(), CA = Tab::CA();
  extern __shared__ __align__(16) unsigned char smem_raw[];  // (dynamic shared memory starts 1024-byte aligned)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Number   *X   = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * Cfg::PER_WARP;  // staging in
  Number   *P   = X + XCAP;                                                             // transposes, staging out
  Number   *W   = P + XCAP;                                                             // coefficient image
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)Cfg::WPB * Cfg::PER_WARP * WB) + 2 * warp;
  const Slab2Lane lm = slab2_lane<n>(lane);
  const bool active = lm.c >= 0;
  const int  cl = lm.cl, ch = lm.ch, x = lm.x, cc = lm.c;
  const uint32_t total_warps = gridDim.x * Cfg::WPB;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // next staged item at or after k (groups the plan leaves to the slab2 kernel are skipped); returns n_items if none
  auto next_item = [&](uint32_t k, uint4 &d) {
    for (; k < n_items; k += total_warps)
      {
        const uint32_t g = glist ? __ldg(glist + k) : k;
        d = __ldg(gdesc + g);
        if ((d.z >> 16) != STAGE_NOPAT) return k;
      }
    return n_items;
  };
  auto group_of = [&](uint32_t k) { return glist ? __ldg(glist + k) : k; };

  // ---- asynchronous copy of the DoF values of a group into X (read_dof_values, first half) ----
  // (pattern tables: header | pos, two rows per uint32 [NS2][32] | own range: slot | flag << 16 [OCAP] | halo slots [HMAX])
  auto issue_own = [&](const uint4 &d, const uint16_t *ph) {
    const uint32_t  own_total = __ldg(ph + STAGE_H_OWN);
    const uint32_t *own32 = reinterpret_cast<const uint32_t *>(ph + STAGE_PH) + NS2 * 32;
#pragma unroll 4
    for (uint32_t e = lane; e < own_total; e += 32) cp_async_elem<WB>(X + (__ldg(own32 + e) & 0xffffu), src + d.x + e);
  };
  auto load_halo_ids = [&](const uint4 &d, uint32_t (&hid)[Cfg::HREG]) {
    const uint32_t nh = d.z & 0xffffu;
#pragma unroll
    for (int t = 0; t < Cfg::HREG; ++t) hid[t] = 32 * t < nh ? __ldg(halo + d.y + 32 * t + lane) : 0u;  // (rows are padded)
  };
  auto issue_halo = [&](const uint4 &d, const uint16_t *ph, const uint32_t (&hid)[Cfg::HREG]) {
    const uint32_t  nh = d.z & 0xffffu;
    const uint16_t *hs = ph + STAGE_PH + 2 * NS2 * 32 + 2 * Cfg::OCAP;
#pragma unroll
    for (int t = 0; t < Cfg::HREG; ++t)
      if (32 * t + lane < nh) cp_async_elem<WB>(X + __ldg(hs + 32 * t + lane), src + hid[t]);
    stage_cp_commit();
  };

  uint4    d;
  uint32_t k = next_item(blockIdx.x * Cfg::WPB + warp, d);
  if (k >= n_items) return;
  if (lane == 0)
    {
      mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      X[XCAP - 1] = Number(0);  // zero slot: never written by a copy
    }
  __syncwarp();
  if (lane == 0) bulk_load(W, cwP + (size_t)group_of(k) * Cfg::F, Cfg::CW_BYTES, bar);
  unsigned phase = 0;
  {
    const uint16_t *ph = ptab + (size_t)(d.z >> 16) * pstride;
    uint32_t hid[Cfg::HREG];
    load_halo_ids(d, hid);
    issue_own(d, ph);
    issue_halo(d, ph, hid);
  }

  const int cAB = AB.SL * cl + AB.SH * ch, cBC = BC.SL * cl + BC.SH * ch, cCA = CA.SL * cl + CA.SH * ch;
  const int bABw = cAB + AB.SI * x, bABr = cAB + AB.SK * x;
  const int bBCw = cBC + BC.SK * x, bBCr = cBC + BC.SJ * x;
  const int bCAw = cCA + CA.SJ * x, bCAr = cCA + CA.SI * x;

  while (k < n_items)
    {
      const uint16_t *ph = ptab + (size_t)(d.z >> 16) * pstride;
      const uint32_t *pp = reinterpret_cast<const uint32_t *>(ph + STAGE_PH) + lane;
      uint4           dn;
      const uint32_t  kn = next_item(k + total_warps, dn);
      const bool      more = kn < n_items;
      const uint16_t *phn = ptab + (size_t)(dn.z >> 16) * pstride;
      Number u[NS], r[NS];
      // ---- read_dof_values, second half: every lane reads its slab u[j + n k] from the staging buffer ----
      {
        uint32_t pz[NS2];
#pragma unroll
        for (int s = 0; s < NS2; ++s) pz[s] = __ldg(pp + 32 * s);
        stage_cp_wait_all();
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NS; ++s) u[s] = X[(pz[s / 2] >> (16 * (s % 2))) & 0x7fffu];
      }
      __syncwarp();  // X is free: the copy of the next group may overwrite it
      uint32_t hid[Cfg::HREG];
      if (more)
        {
          load_halo_ids(dn, hid);
          issue_own(dn, phn);
        }
      // ---- A: N_y, N_z ----
      slab2_apply<n, 1, n, false, Number>(em.N, u);
      slab2_apply<n, n, 1, false, Number>(em.N, u);
      if (more) issue_halo(dn, phn, hid);
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
      if (more && lane == 0) bulk_load(W, cwP + (size_t)group_of(kn) * Cfg::F, Cfg::CW_BYTES, bar);
      // ---- face merges: bit (10 dir + c) of the mask = cell c hands its upper face to cell c + 2^dir of the group ----
      const uint32_t mm = d.w;
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
#pragma unroll
        for (int s = 0; s < NS; ++s)
          {
            const uint32_t pz = (__ldg(pp + 32 * (s / 2)) >> (16 * (s % 2))) & 0xffffu;
            P[(pz & STAGE_DEAD) ? XCAP - 1 : pz] = u[s];
          }
      }
      __syncwarp();
      if (dep_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
      // ---- ... and the warp writes it out in memory order ----
      {
        const uint32_t  own_total = __ldg(ph + STAGE_H_OWN), nh = d.z & 0xffffu;
        const uint32_t *own32 = reinterpret_cast<const uint32_t *>(ph + STAGE_PH) + NS2 * 32;
        const uint16_t *hs = ph + STAGE_PH + 2 * NS2 * 32 + 2 * Cfg::OCAP;
#pragma unroll 4
        for (uint32_t e = lane; e < own_total; e += 32)
          {
            const uint32_t sf = __ldg(own32 + e), f = sf >> 16;
            const Number   v = P[sf & 0xffffu];
            if (f == 1u) dst[d.x + e] = v;
            else if (f == 2u) red_add(dst + d.x + e, v);
          }
#pragma unroll
        for (int t = 0; t < Cfg::HREG; ++t)
          if (32 * t + lane < nh) red_add(dst + __ldg(halo + d.y + 32 * t + lane), P[__ldg(hs + 32 * t + lane)]);
      }
      __syncwarp();  // the next group's transposes reuse P
      k = kn;
      d = dn;
    }
}

struct StageGeom { int n, cw, hc, xcap, hmax, ocap, lcap; };
bool      stage_supported(int dim, int degree, mfg_dtype dt);
StageGeom stage_geom(int degree, mfg_dtype dt);
template <typename Number>
void launch_laplace_stage(int degree, const uint32_t *gdesc, const uint32_t *halo, const uint16_t *ptab, int pstride, const Number *cwP,
                          const Number *src, Number *dst, uint32_t n_items, const double *N, const double *D, int sm_count, cudaStream_t stream,
                          const uint32_t *glist, bool pdl, bool dep_wait, int device);

}  // namespace mfg
