// kernels_slab.cuh -- "slab" Laplace cell kernel for 3D, n = p+1 <= 6.
//
// Same operator as kernels_v0.cuh (and as the reference's apply_kernel_shmem<LocalOperator>,
// matrix_free_gpu.h:318-341 + fee_gpu.cuh:197-365 + tensor_ops.cuh:179-261), different data flow,
// driven by the ncu profile of the column kernel (profiles/r01_v0_*): that kernel is bound by
// L1/shared-memory data-pipe wavefronts (97 %), 3/4 of them the 10 shared-memory round trips per
// cell tensor entry plus bank conflicts.  Here
//   * one thread owns an n x n slab of the cell tensor in registers, n threads per cell, so TWO of
//     the three directions are contracted in registers between shared-memory transposes
//     (7 round trips per entry instead of 10, gather/scatter staging included);
//   * a warp owns floor(32/n) cells and never synchronises with other warps (__syncwarp only);
//   * every transpose uses its own padded address function (tools/slab_layout_search.py) that is
//     bank-conflict free for both the writing and the reading layout;
//   * software pipeline with cp.async (the kernel is register-limited to 10-12 warps per SM, so latency
//     is hidden inside each warp, not by occupancy): the DoF values of the NEXT group are gathered
//     asynchronously (8-byte cp.async, zero-fill for constrained DoFs) behind the back half of the current
//     group; the merged coefficient block of a group (contiguous in global memory, L2-prefetched one group
//     ahead) is copied to shared memory with 16-byte cp.async behind the B phase; both share one buffer;
//   * a single transpose buffer per warp (phases separated by __syncwarp).
// Layouts (thread a of a cell, registers r):
//   L  lexicographic staging order, entry e = 32 q + lane   (coalesced gather / scatter)
//   A  S_xy: a = k, r = i + n j       B  S_yz: a = i, r = j + n k       C  S_xz: a = j, r = i + n k
// Sequence per group of cells:
//   gather -> L ->(LB) B: N_y N_z ->(AB) A: N_x, q-phase x,y ->(AC: G, W) C: q-phase z ->(AC) A: sum,
//   N_x^T N_y^T ->(AB) B: N_z^T ->(LB) L -> scatter (red.add)
#pragma once
#include "kernels_v0.cuh"

namespace mfg {

struct SlabStr { int RJ, RK, SC; };

// strides found by tools/slab_layout_search.py: element (c,i,j,k) at SC*c + i + RJ*j + RK*k
template <int n, int WB> constexpr SlabStr slab_str_LB()
{
  return WB == 8 ? (n == 2 ? SlabStr{4, 8, 18} : n == 3 ? SlabStr{3, 9, 35} : n == 4 ? SlabStr{4, 16, 68} : n == 5 ? SlabStr{5, 25, 133}
                                                                                                        : SlabStr{6, 36, 230})
                 : (n == 2 ? SlabStr{8, 16, 34} : n == 3 ? SlabStr{3, 9, 29} : n == 4 ? SlabStr{4, 16, 68} : n == 5 ? SlabStr{5, 25, 133}
                                                                                                         : SlabStr{6, 36, 218});
}
template <int n, int WB> constexpr SlabStr slab_str_AB()
{
  return WB == 8 ? (n == 2 ? SlabStr{2, 5, 10} : n == 3 ? SlabStr{3, 17, 51} : n == 4 ? SlabStr{4, 17, 68} : n == 5 ? SlabStr{5, 33, 165}
                                                                                                        : SlabStr{6, 41, 246})
                 : (n == 2 ? SlabStr{2, 5, 10} : n == 3 ? SlabStr{5, 30, 93} : n == 4 ? SlabStr{4, 17, 68} : n == 5 ? SlabStr{5, 30, 155}
                                                                                                         : SlabStr{6, 47, 282});
}
template <int n, int WB> constexpr SlabStr slab_str_AC()
{
  return WB == 8 ? (n == 2 ? SlabStr{3, 7, 14} : n == 3 ? SlabStr{3, 19, 57} : n == 4 ? SlabStr{4, 20, 81} : n == 5 ? SlabStr{5, 37, 185}
                                                                                                        : SlabStr{7, 47, 282})
                 : (n == 2 ? SlabStr{3, 7, 14} : n == 3 ? SlabStr{6, 19, 57} : n == 4 ? SlabStr{5, 21, 84} : n == 5 ? SlabStr{5, 27, 135}
                                                                                                         : SlabStr{7, 55, 330});
}
constexpr int cmax3(int a, int b, int c) { return a > b ? (a > c ? a : c) : (b > c ? b : c); }

template <int n, typename Number, int MINB_ = 0> struct SlabCfg
{
  static constexpr int     WB  = (int)sizeof(Number);
  static constexpr int     CW  = 32 / n;          // cells per warp
  static constexpr int     NA  = CW * n;          // active lanes
  static constexpr int     NPC = n * n * n;
  static constexpr int     NS  = n * n;
  static constexpr int     GE  = CW * NPC;        // tensor entries per group
  static constexpr int     Q   = (GE + 31) / 32;  // staging instructions per lane
  static constexpr SlabStr LB = slab_str_LB<n, WB>(), AB = slab_str_AB<n, WB>(), AC = slab_str_AC<n, WB>();
  static constexpr int     BUF = ((cmax3(LB.SC, AB.SC, AC.SC) * CW + 3) / 4) * 4;  // elements of the transpose buffer
  static constexpr int     XBUF = (((LB.SC * CW > GE ? LB.SC * CW : GE) + 3) / 4) * 4;    // elements of the staging / coefficient buffer
  // warps per block.  The register file is per SM sub-partition: 2 warps each -> 255 registers, 3 -> 168, 4 -> 128.
  // MINB_ == 15 selects the low-register variant: 3 blocks x 5 warps = 15 warps per SM at 128 registers.
  static constexpr bool    LOWREG = MINB_ == 15 || MINB_ == 12;  // 12: low-register code at 3 blocks x 4 warps (168 registers)
  static constexpr int     WPB = MINB_ == 15 ? 5 : 4;
  static constexpr size_t  SMEM = (size_t)WPB * (BUF + XBUF) * sizeof(Number);
  static constexpr int     MINB = (MINB_ == 15 || MINB_ == 12) ? 3 : MINB_ ? MINB_ : (n <= 4 ? 4 : 3);
  // cp.async chunk for the coefficient block of one group (GE*WB bytes, contiguous in global memory)
  static constexpr int     CHUNK = (GE * WB) % 16 == 0 ? 16 : 8;
  static constexpr int     NCHUNK = GE * WB / CHUNK;
};

__device__ __forceinline__ int slab_addr(const SlabStr s, int c, int i, int j, int k) { return s.SC * c + i + s.RJ * j + s.RK * k; }

// staging address of lexicographic group entry e
template <int n> __device__ __forceinline__ int slab_stage_addr(const SlabStr s, int e)
{
  constexpr int NPC = n * n * n;
  if (s.RJ == n && s.RK == n * n) return e + (s.SC - NPC) * (e / NPC);
  const int ce = e / NPC, r = e % NPC;
  return slab_addr(s, ce, r % n, (r / n) % n, r / (n * n));
}

template <int BYTES> __device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(sa), "l"(gsrc), "n"(BYTES) : "memory");
}
template <int BYTES> __device__ __forceinline__ void cp_async_zfill(void *smem_dst, const void *gsrc, int src_bytes)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(sa), "l"(gsrc), "n"(BYTES), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// contraction of every line of a slab held in registers: the contracted local index has stride S, lines are
// indexed with stride T (r = l*T + e*S)
template <int n, int S, int T, bool TR, typename Number>
__device__ __forceinline__ void slab_apply(const Number *__restrict__ M, Number (&v)[n * n])
{
#pragma unroll
  for (int l = 0; l < n; ++l)
    {
      Number in[n], out[n];
#pragma unroll
      for (int e = 0; e < n; ++e) in[e] = v[l * T + e * S];
      apply1d<n, TR>(M, in, out);
#pragma unroll
      for (int e = 0; e < n; ++e) v[l * T + e * S] = out[e];
    }
}

template <int n, typename Number, int MINB_>
__global__ void __launch_bounds__(SlabCfg<n, Number, MINB_>::WPB * 32, SlabCfg<n, Number, MINB_>::MINB)
laplace_cell_slab(const uint32_t *__restrict__ idx, const Number *__restrict__ cw, const Number *__restrict__ src,
                  Number *__restrict__ dst, const uint32_t n_cells, const uint32_t n_groups,
                  const __grid_constant__ ShapeMats<Number, n> sh)
{
  using Cfg = SlabCfg<n, Number, MINB_>;
  constexpr int CW = Cfg::CW, NA = Cfg::NA, NPC = Cfg::NPC, NS = Cfg::NS, GE = Cfg::GE, Q = Cfg::Q;
  constexpr SlabStr LB = Cfg::LB, AB = Cfg::AB, AC = Cfg::AC;
  constexpr bool LOWREG = Cfg::LOWREG;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Number   *buf  = reinterpret_cast<Number *>(smem_raw) + (size_t)warp * (Cfg::BUF + Cfg::XBUF);
  // X is time-shared: gathered DoF values of the NEXT group (staging layout LB) from the end of the
  // quadrature phase until the B loads, then the merged coefficient block of the CURRENT group
  // (dense [c][i + n j + n^2 k]) until the end of the quadrature phase.
  Number   *X    = buf + Cfg::BUF;
  const int  c = lane / n, a = lane % n;
  const bool active = lane < NA;
  const uint32_t total_warps = gridDim.x * Cfg::WPB;
  const uint32_t g0 = blockIdx.x * Cfg::WPB + warp;
  if (g0 >= n_groups) return;

  // index row entries of group g owned by this lane (CONSTRAINED_BIT for padding / cells beyond the mesh)
  auto load_ids = [&](uint32_t g, uint32_t (&ids)[Q]) {
    const size_t   ebase = (size_t)g * GE;
    const uint32_t cell0 = g * CW;
    if (cell0 + CW <= n_cells)
      {
#pragma unroll
        for (int q = 0; q < Q; ++q)
          {
            const int e = 32 * q + lane;
            ids[q] = (q < Q - 1 || e < GE) ? __ldg(idx + ebase + e) : CONSTRAINED_BIT;
          }
      }
    else
      {
#pragma unroll
        for (int q = 0; q < Q; ++q)
          {
            const int e = 32 * q + lane;
            ids[q] = (e < GE && cell0 + e / NPC < n_cells) ? __ldg(idx + ebase + e) : CONSTRAINED_BIT;
          }
      }
  };
  // read_dof_values, asynchronous: src[ids] -> X in staging layout; constrained entries are zero-filled
  auto issue_gather = [&](const uint32_t (&ids)[Q], const int lane_o) {
#pragma unroll
    for (int q = 0; q < Q; ++q)
      {
        const int e = 32 * q + lane_o;
        if (q < Q - 1 || e < GE)
          {
            const bool con = ids[q] & CONSTRAINED_BIT;
            cp_async_zfill<sizeof(Number)>(X + slab_stage_addr<n>(LB, e), src + (con ? 0u : ids[q]), con ? 0 : (int)sizeof(Number));
          }
      }
    cp_async_commit();
  };
  auto issue_w = [&](uint32_t g) {
    const char *gsrc = reinterpret_cast<const char *>(cw + (size_t)g * GE);
    char       *sdst = reinterpret_cast<char *>(X);
#pragma unroll
    for (int ch = lane; ch < Cfg::NCHUNK; ch += 32) cp_async<Cfg::CHUNK>(sdst + ch * Cfg::CHUNK, gsrc + ch * Cfg::CHUNK);
    cp_async_commit();
  };
  auto prefetch_l2 = [&](uint32_t g) {  // index rows and coefficient block of a later group
    const char *pi = reinterpret_cast<const char *>(idx + (size_t)g * GE);
    const char *pw = reinterpret_cast<const char *>(cw + (size_t)g * GE);
    for (int l = lane; l < (GE * 4 + 127) / 128; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pi + 128 * l));
    for (int l = lane; l < (GE * Cfg::WB + 127) / 128; l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + 128 * l));
  };

  {
    uint32_t ids[Q];
    load_ids(g0, ids);
    issue_gather(ids, lane);
  }
  const int bLB = slab_addr(LB, c, a, 0, 0), bAB_B = slab_addr(AB, c, a, 0, 0);
  const int bAB_A = slab_addr(AB, c, 0, 0, a), bAC_A = slab_addr(AC, c, 0, 0, a), bAC_C = slab_addr(AC, c, 0, a, 0);

  for (uint32_t g = g0; g < n_groups; g += total_warps)
    {
      const int lane_o = lane;
      const uint32_t gn       = g + total_warps;
      const bool     has_next = gn < n_groups;
      Number G[NS], R[NS];
      cp_async_wait_all();
      __syncwarp();  // gathered values of this group are in X
      // ---- B = S_yz (a = i): N_y, N_z ----
      if (active)
        {
#pragma unroll
          for (int k = 0; k < n; ++k)
#pragma unroll
            for (int j = 0; j < n; ++j) G[j + n * k] = X[bLB + LB.RJ * j + LB.RK * k];
        }
      else
        {
#pragma unroll
          for (int m = 0; m < NS; ++m) G[m] = 0;
        }
      __syncwarp();  // X consumed
      issue_w(g);    // coefficient block of this group (L2-prefetched one group ago) -> X
      if (has_next) prefetch_l2(gn);
      slab_apply<n, 1, n, false>(sh.N, G);
      slab_apply<n, n, 1, false>(sh.N, G);
      if (active)
        {
#pragma unroll
          for (int k = 0; k < n; ++k)
#pragma unroll
            for (int j = 0; j < n; ++j) buf[bAB_B + AB.RJ * j + AB.RK * k] = G[j + n * k];
        }
      __syncwarp();
      // ---- A = S_xy (a = k): N_x, quadrature phases x and y ----
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) G[i + n * j] = buf[bAB_A + i + AB.RJ * j];
        }
      slab_apply<n, 1, n, false>(sh.N, G);  // G = u at the quadrature points
      cp_async_wait_all();  // this lane's part of the coefficient block has landed ...
      __syncwarp();         // ... and everybody else's; buf (AB) has been consumed
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) buf[bAC_A + i + AC.RJ * j] = G[i + n * j];
        }
      {
        const Number *wA = X + NPC * c + NS * a;  // W(c; i, j, k = a) at wA[i + n j]
        // x lines  (LOWREG: operands re-read from this thread's own entries in buf, see below)
#pragma unroll
        for (int j = 0; j < n; ++j)
          {
            Number in[n], gq[n], t[n];
            if (LOWREG)
              {
                if (active)
                  {
#pragma unroll
                    for (int i = 0; i < n; ++i) in[i] = buf[bAC_A + i + AC.RJ * j];
                  }
                else
                  {
#pragma unroll
                    for (int i = 0; i < n; ++i) in[i] = 0;
                  }
              }
            else
              {
#pragma unroll
                for (int i = 0; i < n; ++i) in[i] = G[i + n * j];
              }
            apply1d<n, false>(sh.D, in, gq);
            if (active)
              {
#pragma unroll
                for (int i = 0; i < n; ++i) gq[i] *= wA[i + n * j];
              }
            apply1d<n, true>(sh.D, gq, t);
#pragma unroll
            for (int i = 0; i < n; ++i) R[i + n * j] = t[i];
          }
        // y lines.  LOWREG: re-read G from this thread's own entries in buf (written just above) so that G's registers
        // are dead after the x lines -- one slab less in the register peak (128 registers, 15 warps per SM)
#pragma unroll
        for (int i = 0; i < n; ++i)
          {
            Number in[n], gq[n], t[n];
            if (LOWREG)
              {
                if (active)
                  {
#pragma unroll
                    for (int j = 0; j < n; ++j) in[j] = buf[bAC_A + i + AC.RJ * j];
                  }
                else
                  {
#pragma unroll
                    for (int j = 0; j < n; ++j) in[j] = 0;
                  }
              }
            else
              {
#pragma unroll
                for (int j = 0; j < n; ++j) in[j] = G[i + n * j];
              }
            apply1d<n, false>(sh.D, in, gq);
            if (active)
              {
#pragma unroll
                for (int j = 0; j < n; ++j) gq[j] *= wA[i + n * j];
              }
            apply1d<n, true>(sh.D, gq, t);
#pragma unroll
            for (int j = 0; j < n; ++j) R[i + n * j] += t[j];
          }
      }
      __syncwarp();
      // ---- C = S_xz (a = j): quadrature phase z, streamed line by line; result replaces G in buf ----
      if (active)
        {
          const Number *wC = X + NPC * c + n * a;  // W(c; i, j = a, k) at wC[i + n^2 k]
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              Number in[n], gq[n], t[n];
#pragma unroll
              for (int k = 0; k < n; ++k) in[k] = buf[bAC_C + i + AC.RK * k];
              apply1d<n, false>(sh.D, in, gq);
#pragma unroll
              for (int k = 0; k < n; ++k) gq[k] *= wC[i + NS * k];
              apply1d<n, true>(sh.D, gq, t);
#pragma unroll
              for (int k = 0; k < n; ++k) buf[bAC_C + i + AC.RK * k] = t[k];
            }
        }
      __syncwarp();  // X is free again
      // index rows of the next group (L2 hits): in flight during the A phase below
      uint32_t ids[Q];
      if (has_next) load_ids(gn, ids);
      // ---- A: sum the three directions, N_x^T, N_y^T ----
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) R[i + n * j] += buf[bAC_A + i + AC.RJ * j];
        }
      __syncwarp();
      slab_apply<n, 1, n, true>(sh.N, R);
      slab_apply<n, n, 1, true>(sh.N, R);
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) buf[bAB_A + i + AB.RJ * j] = R[i + n * j];
        }
      // asynchronous gather of the next group's DoF values behind the rest of this group
      if (has_next) issue_gather(ids, lane_o);
      __syncwarp();
      // ---- B: N_z^T ----
      if (active)
        {
#pragma unroll
          for (int k = 0; k < n; ++k)
#pragma unroll
            for (int j = 0; j < n; ++j) R[j + n * k] = buf[bAB_B + AB.RJ * j + AB.RK * k];
        }
      __syncwarp();
      slab_apply<n, n, 1, true>(sh.N, R);
      if (active)
        {
#pragma unroll
          for (int k = 0; k < n; ++k)
#pragma unroll
            for (int j = 0; j < n; ++j) buf[bLB + LB.RJ * j + LB.RK * k] = R[j + n * k];
        }
      __syncwarp();
      // ---- distribute_local_to_global: lexicographic order, red.add ----
      {
        const size_t   ebase = (size_t)g * GE;
        const uint32_t cell0 = g * CW;
        if (cell0 + CW <= n_cells)
          {
#pragma unroll
            for (int q = 0; q < Q; ++q)
              {
                const int e = 32 * q + lane_o;
                if (q < Q - 1 || e < GE)
                  {
                    const uint32_t id = __ldg(idx + ebase + e);
                    if (!(id & CONSTRAINED_BIT)) red_add(dst + id, buf[slab_stage_addr<n>(LB, e)]);
                  }
              }
          }
        else
          {
#pragma unroll
            for (int q = 0; q < Q; ++q)
              {
                const int e = 32 * q + lane_o;
                if (e < GE && cell0 + e / NPC < n_cells)
                  {
                    const uint32_t id = __ldg(idx + ebase + e);
                    if (!(id & CONSTRAINED_BIT)) red_add(dst + id, buf[slab_stage_addr<n>(LB, e)]);
                  }
              }
          }
      }
      __syncwarp();
    }
  cp_async_wait_all();
}

template <typename Number>
void launch_laplace_slab(int degree, int min_blocks, const uint32_t *idx, const Number *cw, const Number *src, Number *dst, uint32_t n_cells,
                         const double *N, const double *D, int sm_count, cudaStream_t stream);
bool slab_supported(int dim, int degree, mfg_dtype dt);
int  slab_cells_per_group(int degree);
size_t slab_cw_padded_cells(uint32_t n_cells);

}  // namespace mfg
