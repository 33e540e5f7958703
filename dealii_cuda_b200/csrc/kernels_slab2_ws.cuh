// kernels_slab2_ws.cuh -- warp-specialised form of the slab2 Laplace cell kernel (3D, n = 5).
//
// The slab2 kernel at 12 warps per SM is latency bound: every warp serialises index rows -> gather -> contractions ->
// scatter, no unit is busier than 85 %.  Its plane-layout head (kernels_slab2.cuh, KGS) talks to the
// contraction phases only through a shared-memory buffer, so the head can run in other warps, groups ahead:
//   loader warps   : index rows, gather, N_z, store the (i,j)-planes into the consumer's buffer P[b]      (few registers)
//   contraction warps: wait full[b]; B: N_x N_y, quadrature phases x, y; C: phase z, N_x^T N_z^T; release P[b];
//                      plane-layout tail: N_y^T, red.add                                                  (never wait for a gather)
// Hand-off by mbarriers (full / empty per buffer, two buffers per contraction warp); the coefficient image still comes
// by one bulk-async copy per group.  Same arithmetic, layouts and arrays as the KGS configuration of slab2.
#pragma once
#include "kernels_slab2.cuh"

namespace mfg {

template <int n, typename Number, int NCW_ = 4, int NLW_ = 2, int MINB_ = 2, bool LS_ = false, int RC_ = 0, int RL_ = 0> struct Slab2WsCfg
{
  static constexpr int WB = (int)sizeof(Number);
  using Tab = Slab2Tab<n, WB>;
  static constexpr int CW = 32 / n, NPC = n * n * n, NS = n * n;
  static constexpr int NCW = NCW_;              // contraction warps per CTA
  static constexpr int NLW = NLW_;              // loader warps per CTA (each serves NCW / NLW contraction warps)
  static constexpr int MINB = MINB_;            // CTAs per SM
  static constexpr bool LS = LS_;               // the loader warps also run the scatter (N_y^T + red) of the groups they loaded
  // register reallocation between the warp groups (setmaxnreg, needs NCW = NLW = 4: one warp group each): the contraction
  // warps grow to RC registers, the loader warps shrink to RL; 0 = every warp keeps the launch allocation
  static constexpr int RC = RC_, RL = RL_;
  static_assert(RC_ == 0 || (NCW_ == 4 && NLW_ == 4), "setmaxnreg works on warp groups of 4 warps");
  static_assert(NCW % NLW == 0, "every loader serves the same number of contraction warps");
  static constexpr int THREADS = (NCW + NLW) * 32;
  static constexpr int F = Tab::F;              // elements of one transpose buffer / of the coefficient image
  static_assert(F >= CW * NPC, "dense buffers must fit");
  static constexpr int PER_WARP = 4 * F;        // W, P[0], P[1], Q
  static constexpr size_t SMEM = 64 * NCW /* 5 mbarriers per contraction warp, padded */ + (size_t)NCW * PER_WARP * sizeof(Number);
  static constexpr uint32_t CW_BYTES = F * WB;
};

// predicated red.global.add (no branch around the instruction): skipped when `skip` is set
__device__ __forceinline__ void red_add_unless(double *addr, double v, bool skip)
{
  asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %2, 0;\n@p red.global.add.f64 [%0], %1;\n}" ::"l"(addr), "d"(v), "r"((unsigned)skip) : "memory");
}
__device__ __forceinline__ void red_add_unless(float *addr, float v, bool skip)
{
  asm volatile("{\n.reg .pred p;\nsetp.eq.u32 p, %2, 0;\n@p red.global.add.f32 [%0], %1;\n}" ::"l"(addr), "f"(v), "r"((unsigned)skip) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int n, typename Number, int NCW_, int NLW_, int MINB_, bool LS_, int RC_, int RL_>
__global__ void __launch_bounds__((Slab2WsCfg<n, Number, NCW_, NLW_, MINB_, LS_, RC_, RL_>::THREADS), MINB_)
laplace_cell_slab2_ws(const uint32_t *__restrict__ idxLex, const uint32_t *__restrict__ idxJ, const Number *__restrict__ cwP,
                      const Number *__restrict__ src, Number *__restrict__ dst, const uint32_t n_groups, const uint32_t n_cells,
                      const __grid_constant__ EoMats<Number, n> em)
{
  using Cfg = Slab2WsCfg<n, Number, NCW_, NLW_, MINB_, LS_, RC_, RL_>;
  using Tab = typename Cfg::Tab;
  constexpr int NS = Cfg::NS, CW = Cfg::CW, NPC = Cfg::NPC, F = Cfg::F;
  constexpr int HC = CW % 2 == 0 ? CW / 2 : CW, HCn = HC * NPC;
  constexpr Slab2Lay KB = Slab2Lay{NPC, HCn, 1, n, n * n};   // dense: n^3 c + n^2 k + n j + i   (loader planes -> B)
  constexpr Slab2Lay CK = Slab2Lay{NPC, HCn, 1, n * n, n};   // dense: n^3 c + n^2 j + n k + i   (C -> scatter planes)
  constexpr Slab2Lay BC = Tab::BC();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t total_cw = gridDim.x * Cfg::NCW;  // contraction warps of the grid = stride of the work list
  auto bars_of = [&](int w) { return reinterpret_cast<uint64_t *>(smem_raw) + 8 * w; };  // [0,1] full, [2,3] empty, [4] W
  auto bufs_of = [&](int w) { return reinterpret_cast<Number *>(smem_raw + 64 * Cfg::NCW) + (size_t)w * Cfg::PER_WARP; };

  if (threadIdx.x < Cfg::NCW)
    {
      uint64_t *b = bars_of(threadIdx.x);
      for (int i = 0; i < 5; ++i) mbar_init(b + i, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();

  const bool pl = lane < NS;
  if (warp >= Cfg::NCW)
    {
      // ------------------------------------------------ loader ------------------------------------------------
      if (Cfg::RL > 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::RL));
      if (Cfg::RL > 0)
        {
          // ---- lean loader / scatter warp (fits the shrunk register allocation): it serves ONE contraction warp and has that
          //      warp's whole contraction time per group, so the group is handled in two halves of CW / 2 cells ----
          static_assert(Cfg::RL == 0 || (Cfg::LS && Cfg::NCW == Cfg::NLW && CW % 2 == 0), "lean helper: one warp each, LS mode");
          constexpr int HCELLS = CW / 2;
          const int      w = warp - Cfg::NCW;
          uint64_t      *bars = bars_of(w);
          Number        *Pb = bufs_of(w) + F;
          const uint32_t kw = blockIdx.x * Cfg::NCW + w;
          if (kw >= n_groups) return;
          const uint32_t n_it = (n_groups - kw + total_cw - 1) / total_cw;
          auto half_ids = [&](uint32_t g, int h, const uint32_t *base, uint32_t (&id)[HCELLS][n]) {
#pragma unroll
            for (int c = 0; c < HCELLS; ++c)
              {
                const uint32_t  cell = g * CW + h * HCELLS + c;
                const bool      ok = pl && cell < n_cells;
                const uint32_t *row = base + (size_t)cell * NPC + lane;
#pragma unroll
                for (int kk = 0; kk < n; ++kk) id[c][kk] = ok ? __ldg(row + NS * kk) : CONSTRAINED_BIT;
              }
          };
          auto tail = [&](uint32_t t) {  // scatter of iteration t from its buffer
            const uint32_t g = kw + t * total_cw;
            const Number  *P = Pb + F * (t & 1);
            bool waited = false;
#pragma unroll
            for (int h = 0; h < 2; ++h)
              {
                uint32_t jid[HCELLS][n];
                half_ids(g, h, idxJ, jid);
                if (!waited) { mbar_wait(bars + 2 + (t & 1), (t >> 1) & 1); waited = true; }
#pragma unroll
                for (int c = 0; c < HCELLS; ++c)
                  {
                    const int     cc = h * HCELLS + c;
                    const Number *qc = P + CK.SL * (cc % HC) + CK.SH * (cc / HC) + (pl ? lane : 0);
                    Number in[n], out[n];
#pragma unroll
                    for (int j = 0; j < n; ++j) in[j] = qc[CK.SJ * j];
                    eo_apply<n, false>(em.NT, in, out);
#pragma unroll
                    for (int j = 0; j < n; ++j)
                      if (!(jid[c][j] & CONSTRAINED_BIT)) red_add(dst + jid[c][j], out[j]);
                  }
              }
            __syncwarp();
          };
          for (uint32_t it = 0; it < n_it; ++it)
            {
              const uint32_t g = kw + it * total_cw;
              Number        *P = Pb + F * (it & 1);
              uint32_t kid[HCELLS][n];
              half_ids(g, 0, idxLex, kid);           // in flight during the scatter of the buffer's previous group
              if (it >= 2) tail(it - 2);
#pragma unroll
              for (int h = 0; h < 2; ++h)
                {
                  Number kv[HCELLS][n];
#pragma unroll
                  for (int c = 0; c < HCELLS; ++c)
#pragma unroll
                    for (int kk = 0; kk < n; ++kk) kv[c][kk] = (kid[c][kk] & CONSTRAINED_BIT) ? Number(0) : __ldg(src + kid[c][kk]);
                  if (h == 0) half_ids(g, 1, idxLex, kid);   // the second half's index rows travel with the first half's values
#pragma unroll
                  for (int c = 0; c < HCELLS; ++c)
                    {
                      const int cc = h * HCELLS + c;
                      Number out[n];
                      eo_apply<n, false>(em.N, kv[c], out);
                      if (pl)
                        {
                          Number *pc = P + KB.SL * (cc % HC) + KB.SH * (cc / HC) + lane;
#pragma unroll
                          for (int kk = 0; kk < n; ++kk) pc[KB.SK * kk] = out[kk];
                        }
                    }
                }
              __syncwarp();
              if (lane == 0) mbar_arrive(bars + (it & 1));  // full[b]
            }
          for (uint32_t t = n_it >= 2 ? n_it - 2 : 0; t < n_it; ++t) tail(t);
          return;
        }
      const int l = warp - Cfg::NCW;
      constexpr int SERVE = Cfg::NCW / Cfg::NLW;
      // work items q = 0, 1, ...: iteration it = q / SERVE of contraction warp w = l SERVE + q % SERVE.  The index rows of
      // item q + 1 are requested before the gathered values of item q are used: one exposed memory latency per item.
      auto item_group = [&](uint32_t q) { return blockIdx.x * Cfg::NCW + l * SERVE + q % SERVE + (q / SERVE) * total_cw; };
      auto load_ids = [&](uint32_t g, uint32_t (&id)[CW][n]) {
#pragma unroll
        for (int c = 0; c < CW; ++c)
          {
            const uint32_t  cell = g * CW + c;
            const bool      ok = pl && g < n_groups && cell < n_cells;
            const uint32_t *row = idxLex + (size_t)cell * NPC + lane;
#pragma unroll
            for (int kk = 0; kk < n; ++kk) id[c][kk] = ok ? __ldg(row + NS * kk) : CONSTRAINED_BIT;
          }
      };
      // LS: scatter of one group from its buffer (written by the contraction warp in the C -> K' layout), lane t <-> (i, k)
      auto scatter_tail = [&](int w, uint32_t g, uint32_t it) {
        const int b = it & 1;
        uint64_t *bars = bars_of(w);
        const Number *P = bufs_of(w) + F * (1 + b);
        uint32_t jid[CW][n];
#pragma unroll
        for (int c = 0; c < CW; ++c)
          {
            const uint32_t  cell = g * CW + c;
            const bool      ok = pl && cell < n_cells;
            const uint32_t *row = idxJ + (size_t)cell * NPC + lane;
#pragma unroll
            for (int j = 0; j < n; ++j) jid[c][j] = ok ? __ldg(row + NS * j) : CONSTRAINED_BIT;
          }
        mbar_wait(bars + 2 + b, (it >> 1) & 1);  // tail[b]: the contraction warp has stored the group's result planes
#pragma unroll
        for (int c = 0; c < CW; ++c)
          {
            const Number *qc = P + CK.SL * (c % HC) + CK.SH * (c / HC) + (pl ? lane : 0);
            Number in[n], out[n];
#pragma unroll
            for (int j = 0; j < n; ++j) in[j] = qc[CK.SJ * j];
            eo_apply<n, false>(em.NT, in, out);
#pragma unroll
            for (int j = 0; j < n; ++j)
              if (!(jid[c][j] & CONSTRAINED_BIT)) red_add(dst + jid[c][j], out[j]);
          }
        __syncwarp();  // every lane has read the buffer: it may be refilled
      };
      uint32_t kid[CW][n], kidn[CW][n];
      load_ids(item_group(0), kid);
      for (uint32_t q = 0;; ++q)
        {
          if (item_group(q - q % SERVE) >= n_groups)
            {
              if (Cfg::LS)
                {
                  // the last two groups of every served warp are still waiting for their scatter
                  const uint32_t it_end = q / SERVE;  // first iteration without work for anybody
#pragma unroll
                  for (int sw = 0; sw < SERVE; ++sw)
                    {
                      const int      w = l * SERVE + sw;
                      const uint32_t kw = blockIdx.x * Cfg::NCW + w;
                      if (kw >= n_groups) continue;
                      const uint32_t n_it = (n_groups - kw + total_cw - 1) / total_cw;  // iterations of warp w
                      // in the loop the scatter of iteration t ran together with the fill of t + 2 (if that one existed)
                      for (uint32_t t = n_it >= 2 ? n_it - 2 : 0; t < n_it; ++t) scatter_tail(w, kw + t * total_cw, t);
                      (void)it_end;
                    }
                }
              break;  // no served warp has work in this iteration
            }
          const uint32_t g = item_group(q);
          const bool     valid = g < n_groups;
          Number kv[CW][n];
          if (valid)
            {
#pragma unroll
              for (int c = 0; c < CW; ++c)
#pragma unroll
                for (int kk = 0; kk < n; ++kk) kv[c][kk] = (kid[c][kk] & CONSTRAINED_BIT) ? Number(0) : __ldg(src + kid[c][kk]);
            }
          load_ids(item_group(q + 1), kidn);
          if (valid)
            {
              const uint32_t it = q / SERVE;
              const int      w = l * SERVE + q % SERVE, b = it & 1;
              uint64_t      *bars = bars_of(w);
              Number        *P = bufs_of(w) + F * (1 + b);
              if (Cfg::LS)
                {
                  // the buffer still holds the result planes of the group of two iterations ago: scatter them first
                  if (it >= 2) scatter_tail(w, g - 2 * total_cw, it - 2);
                }
              else
                // the buffer must have been released by the contraction warp (its use two iterations ago)
                mbar_wait(bars + 2 + b, ((it >> 1) & 1) ^ 1);
#pragma unroll
              for (int c = 0; c < CW; ++c)
                {
                  Number out[n];
                  eo_apply<n, false>(em.N, kv[c], out);
                  if (pl)
                    {
                      Number *pc = P + KB.SL * (c % HC) + KB.SH * (c / HC) + lane;
#pragma unroll
                      for (int kk = 0; kk < n; ++kk) pc[KB.SK * kk] = out[kk];
                    }
                }
              __syncwarp();
              if (lane == 0) mbar_arrive(bars + b);  // full[b]
            }
#pragma unroll
          for (int c = 0; c < CW; ++c)
#pragma unroll
            for (int kk = 0; kk < n; ++kk) kid[c][kk] = kidn[c][kk];
        }
      return;
    }

  // ------------------------------------------------ contraction warp ------------------------------------------------
  if (Cfg::RC > 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::RC));
  uint64_t *bars = bars_of(warp);
  Number   *W = bufs_of(warp), *Q = W + 3 * F;
  const Slab2Lane lm = slab2_lane<n>(lane);
  const bool active = lm.c >= 0;
  const int  cl = lm.cl, ch = lm.ch, x = lm.x;
  const int  cKB = KB.SL * cl + KB.SH * ch, cBC = BC.SL * cl + BC.SH * ch, cCK = CK.SL * cl + CK.SH * ch;
  const int  bKBr = cKB + KB.SK * x;                          // B: x = k
  const int  bBCw = cBC + BC.SK * x, bBCr = cBC + BC.SJ * x;  // B: x = k ; C: x = j
  const int  bCKw = cCK + CK.SJ * x;                          // C: x = j
  const uint32_t k0 = blockIdx.x * Cfg::NCW + warp;
  if (k0 >= n_groups) return;
  if (lane == 0) bulk_load(W, cwP + (size_t)k0 * F, Cfg::CW_BYTES, bars + 4);
  uint32_t it = 0;
  for (uint32_t k = k0; k < n_groups; k += total_cw, ++it)
    {
      const uint32_t g = k;
      const bool     more = k + total_cw < n_groups;
      const int      b = it & 1;
      Number        *P = W + F * (1 + b);
      Number u[NS], r[NS];
      mbar_wait(bars + b, (it >> 1) & 1);  // full[b]: the loader has stored N_z u of this group
      // ---- B: N_x, N_y -> u at the quadrature points, u[i + n j] ----
#pragma unroll
      for (int j = 0; j < n; ++j)
#pragma unroll
        for (int i = 0; i < n; ++i) u[i + n * j] = P[bKBr + KB.SI * i + KB.SJ * j];
      __syncwarp();  // P consumed: it will carry r
      slab2_apply<n, 1, n, false>(em.N, u);
      slab2_apply<n, n, 1, false>(em.N, u);
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) Q[bBCw + BC.SI * i + BC.SJ * j] = u[i + n * j];
        }
      mbar_wait(bars + 4, it & 1);  // coefficient image of this group has landed
#pragma unroll
      for (int j = 0; j < n; ++j)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int i = 0; i < n; ++i) in[i] = u[i + n * j];
          eo_apply<n, true>(em.D, in, gq);
#pragma unroll
          for (int i = 0; i < n; ++i) gq[i] *= W[bBCw + BC.SI * i + BC.SJ * j];
          eo_apply<n, true>(em.DT, gq, t);
#pragma unroll
          for (int i = 0; i < n; ++i) r[i + n * j] = t[i];
        }
#pragma unroll
      for (int i = 0; i < n; ++i)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int j = 0; j < n; ++j) in[j] = u[i + n * j];
          eo_apply<n, true>(em.D, in, gq);
#pragma unroll
          for (int j = 0; j < n; ++j) gq[j] *= W[bBCw + BC.SI * i + BC.SJ * j];
          eo_apply<n, true>(em.DT, gq, t);
#pragma unroll
          for (int j = 0; j < n; ++j) r[i + n * j] += t[j];
        }
      if (active)
        {
#pragma unroll
          for (int j = 0; j < n; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) P[bBCw + BC.SI * i + BC.SJ * j] = r[i + n * j];
        }
      __syncwarp();
      // ---- C: quadrature phase z on u[i + n k], sum, N_x^T, N_z^T ----
#pragma unroll
      for (int kk = 0; kk < n; ++kk)
#pragma unroll
        for (int i = 0; i < n; ++i) u[i + n * kk] = Q[bBCr + BC.SI * i + BC.SK * kk];
#pragma unroll
      for (int i = 0; i < n; ++i)
        {
          Number in[n], gq[n], t[n];
#pragma unroll
          for (int kk = 0; kk < n; ++kk) in[kk] = u[i + n * kk];
          eo_apply<n, true>(em.D, in, gq);
#pragma unroll
          for (int kk = 0; kk < n; ++kk) gq[kk] *= W[bBCr + BC.SI * i + BC.SK * kk];
          eo_apply<n, true>(em.DT, gq, t);
#pragma unroll
          for (int kk = 0; kk < n; ++kk) u[i + n * kk] = t[kk] + P[bBCr + BC.SI * i + BC.SK * kk];
        }
      __syncwarp();  // P, Q and the coefficient image are consumed
      if (lane == 0)
        {
          if (!Cfg::LS) mbar_arrive(bars + 2 + b);  // empty[b]
          if (more) bulk_load(W, cwP + (size_t)(k + total_cw) * F, Cfg::CW_BYTES, bars + 4);
        }
      if (Cfg::LS)
        {
          // ---- N_x^T, N_z^T, result planes into the group's buffer: the loader warp scatters them (tail[b]) ----
          slab2_apply<n, 1, n, false>(em.NT, u);
          slab2_apply<n, n, 1, false>(em.NT, u);
          if (active)
            {
#pragma unroll
              for (int kk = 0; kk < n; ++kk)
#pragma unroll
                for (int i = 0; i < n; ++i) P[bCKw + CK.SI * i + CK.SK * kk] = u[i + n * kk];
            }
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + 2 + b);  // tail[b]
        }
      else
        {
          // the scatter's index rows: issued here (r is dead), their latency hides behind N_x^T N_z^T and the transpose
          uint32_t jid[CW][n];
#pragma unroll
          for (int c = 0; c < CW; ++c)
            {
              const uint32_t  cell = g * CW + c;
              const bool      ok = pl && cell < n_cells;
              const uint32_t *row = idxJ + (size_t)cell * NPC + lane;
#pragma unroll
              for (int j = 0; j < n; ++j) jid[c][j] = ok ? __ldg(row + NS * j) : CONSTRAINED_BIT;
            }
          slab2_apply<n, 1, n, false>(em.NT, u);
          slab2_apply<n, n, 1, false>(em.NT, u);
          if (active)
            {
#pragma unroll
              for (int kk = 0; kk < n; ++kk)
#pragma unroll
                for (int i = 0; i < n; ++i) Q[bCKw + CK.SI * i + CK.SK * kk] = u[i + n * kk];
            }
          __syncwarp();
          // ---- K': one cell per pass, lane t <-> (i, k), j in registers: N_y^T, red.add ----
          // (loading and contracting all cells first and issuing the 30 reds back to back measured 4 % slower)
#pragma unroll
          for (int c = 0; c < CW; ++c)
            {
              const Number *qc = Q + CK.SL * (c % HC) + CK.SH * (c / HC) + (pl ? lane : 0);
              Number in[n], out[n];
#pragma unroll
              for (int j = 0; j < n; ++j) in[j] = qc[CK.SJ * j];
              eo_apply<n, false>(em.NT, in, out);
#pragma unroll
              for (int j = 0; j < n; ++j)
                if (!(jid[c][j] & CONSTRAINED_BIT)) red_add(dst + jid[c][j], out[j]);
            }
          __syncwarp();  // Q is written again by the next group
        }
    }
}

template <typename Number>
void launch_laplace_slab2_ws(int degree, int shape, const uint32_t *idxLex, const uint32_t *idxJ, const Number *cwP, const Number *src, Number *dst,
                             uint32_t n_groups, uint32_t n_cells, const double *N, const double *D, int sm_count, cudaStream_t stream);

}  // namespace mfg
