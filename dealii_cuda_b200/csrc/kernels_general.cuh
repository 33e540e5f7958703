// kernels_general.cuh -- Laplace cell kernel for non-affine geometry: full inverse Jacobian per quadrature point
// (the reference without -DMATRIX_FREE_UNIFORM_MESH: FEEvaluationGpu::get_gradient / submit_gradient,
// fee_gpu.cuh:219-246, 261-284, with the J^-1 arrays of matrix_free_gpu.cu:326-338; BALL_GRID in poisson_common.h:65-70).
//
// The two metric applications and the coefficient are merged at setup into one symmetric tensor per quadrature point
//     G(q) = a(x_q) JxW_q K_q K_q^T,   K = J^-1 = d xi / d x   (get_gradient: grad_x = K^T grad_xi ; submit: K (a grad_x) JxW),
// stored as dim (dim + 1) / 2 planes per cell (00, 11[, 22], 01[, 02, 12]).  The cell kernel is
//     u_q = (N x N x N) u ;  g_d = D_d u_q ;  t_d = sum_e G_de g_e ;  r = sum_d D_d^T t_d ;  out = (N x N x N)^T r
// in the collocation form of the other kernels.  One thread per tensor entry like the reference's kernel (a CTA holds
// several cells); this path is about coverage, the uniform-mesh kernels are the ones tuned for bandwidth.
#pragma once
#include "kernels_v0.cuh"

namespace mfg {

__host__ __device__ constexpr int gen_cells_per_block(int dim, int n)
{
  const int npc = ipow(n, dim);
  return npc >= 128 ? 1 : 128 / npc;
}

// contraction along direction d of the cell tensor in shared memory: out(.., q, ..) = sum_k M[k*n+q] in(.., k, ..)
// (TR: M[q*n+k]); every thread computes its own entry
template <int dim, int n, typename Number>
__device__ __forceinline__ Number gen_contract(const Number *__restrict__ M, const Number *__restrict__ in, int e, int d, bool tr)
{
  const int stride = d == 0 ? 1 : d == 1 ? n : n * n;
  const int q = (e / stride) % n, base = e - q * stride;
  Number acc = 0;
#pragma unroll
  for (int k = 0; k < n; ++k) acc += (tr ? M[q * n + k] : M[k * n + q]) * in[base + k * stride];
  return acc;
}

template <int dim, int n, typename Number, bool ATOMIC>
__global__ void laplace_cell_general(const uint32_t *__restrict__ idx, const Number *__restrict__ gsym, const Number *__restrict__ src,
                                     Number *__restrict__ dst, const uint32_t cell_begin, const uint32_t cell_end,
                                     const __grid_constant__ ShapeMats<Number, n> sh)
{
  constexpr int NPC = ipow(n, dim), CPB = gen_cells_per_block(dim, n), NC = dim * (dim + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *smem = reinterpret_cast<Number *>(smem_raw);
  const int  lc = threadIdx.x / NPC, e = threadIdx.x % NPC;
  const bool in_blk = lc < CPB;
  const uint32_t cell = cell_begin + blockIdx.x * CPB + lc;
  const bool active = in_blk && cell < cell_end;
  Number *A = smem + (size_t)(in_blk ? lc : 0) * (1 + dim) * NPC;  // values
  Number *T = A + NPC;                                             // dim planes

  // read_dof_values (fee_gpu.cuh:323-338)
  uint32_t id = CONSTRAINED_BIT;
  if (active) id = idx[(size_t)cell * NPC + e];
  Number v = (id & CONSTRAINED_BIT) ? Number(0) : __ldg(src + id);
  // interpolate to the quadrature points, direction by direction (A -> T[0] -> A ...)
  Number *in = A, *out = T;
  if (in_blk) in[e] = v;
  __syncthreads();
#pragma unroll
  for (int d = 0; d < dim; ++d)
    {
      if (in_blk) out[e] = gen_contract<dim, n>(sh.N, in, e, d, false);
      __syncthreads();
      Number *tmp = in; in = out; out = tmp;
    }
  // `in` holds u at the quadrature points; reference-space gradient by the collocation derivative
  Number g[dim];
#pragma unroll
  for (int d = 0; d < dim; ++d) g[d] = in_blk ? gen_contract<dim, n>(sh.D, in, e, d, false) : Number(0);
  __syncthreads();  // all reads of `in` done: A and T are free
  // t = G g (quad_operation fused with both metric terms)
  if (active)
    {
      const Number *G = gsym + (size_t)cell * NC * NPC + e;
      if (dim == 2)
        {
          const Number g00 = G[0], g11 = G[NPC], g01 = G[2 * NPC];
          T[e] = g00 * g[0] + g01 * g[1];
          T[NPC + e] = g01 * g[0] + g11 * g[1];
        }
      else
        {
          const Number g00 = G[0], g11 = G[NPC], g22 = G[2 * NPC], g01 = G[3 * NPC], g02 = G[4 * NPC], g12 = G[5 * NPC];
          T[e] = g00 * g[0] + g01 * g[1] + g02 * g[dim - 1];
          T[NPC + e] = g01 * g[0] + g11 * g[1] + g12 * g[dim - 1];
          T[(dim - 1) * NPC + e] = g02 * g[0] + g12 * g[1] + g22 * g[dim - 1];
        }
    }
  else if (in_blk)
    {
#pragma unroll
      for (int d = 0; d < dim; ++d) T[d * NPC + e] = 0;
    }
  __syncthreads();
  // r = sum_d D_d^T t_d
  Number r = 0;
  if (in_blk)
    {
#pragma unroll
      for (int d = 0; d < dim; ++d) r += gen_contract<dim, n>(sh.D, T + d * NPC, e, d, true);
    }
  __syncthreads();
  // integrate: N^T direction by direction
  in = A; out = T;
  if (in_blk) in[e] = r;
  __syncthreads();
#pragma unroll
  for (int d = 0; d < dim; ++d)
    {
      if (in_blk) out[e] = gen_contract<dim, n>(sh.N, in, e, d, true);
      __syncthreads();
      Number *tmp = in; in = out; out = tmp;
    }
  // distribute_local_to_global (fee_gpu.cuh:346-365)
  if (active && !(id & CONSTRAINED_BIT))
    {
      if (ATOMIC) red_add(dst + id, in[e]);
      else dst[id] += in[e];
    }
}

// diagonal for general geometry: diag_i = sum_q sum_de G_de(q) d_d phi_i(q) d_e phi_i(q)   (thread = local DoF i)
template <int dim, typename Number>
__global__ void diagonal_general(const uint32_t *__restrict__ idx, const Number *__restrict__ gsym, int n, uint32_t n_cells,
                                 const double *__restrict__ val, const double *__restrict__ grad, Number *__restrict__ diag)
{
  const int npc = dim == 2 ? n * n : n * n * n;
  const uint32_t cell = blockIdx.x;
  constexpr int NC = dim * (dim + 1) / 2;
  for (int i = threadIdx.x; i < npc; i += blockDim.x)
    {
      const int i0 = i % n, i1 = (i / n) % n, i2 = i / (n * n);
      double acc = 0;
      for (int q = 0; q < npc; ++q)
        {
          const int q0 = q % n, q1 = (q / n) % n, q2 = q / (n * n);
          const double v0 = val[i0 * n + q0], v1 = val[i1 * n + q1], v2 = dim == 3 ? val[i2 * n + q2] : 1.0;
          const double d0 = grad[i0 * n + q0], d1 = grad[i1 * n + q1], d2 = dim == 3 ? grad[i2 * n + q2] : 0.0;
          double gp[3] = {d0 * v1 * v2, v0 * d1 * v2, v0 * v1 * d2};
          const Number *G = gsym + (size_t)cell * NC * npc + q;
          if (dim == 2)
            acc += (double)G[0] * gp[0] * gp[0] + (double)G[npc] * gp[1] * gp[1] + 2.0 * (double)G[2 * npc] * gp[0] * gp[1];
          else
            acc += (double)G[0] * gp[0] * gp[0] + (double)G[npc] * gp[1] * gp[1] + (double)G[2 * npc] * gp[2] * gp[2] +
                   2.0 * ((double)G[3 * npc] * gp[0] * gp[1] + (double)G[4 * npc] * gp[0] * gp[2] + (double)G[5 * npc] * gp[1] * gp[2]);
        }
      const uint32_t g = idx[(size_t)cell * npc + i];
      if (!(g & CONSTRAINED_BIT)) atomicAdd(diag + g, (Number)acc);
    }
}

template <int dim, typename Number>
void launch_laplace_general_dim(int degree, bool atomic, const uint32_t *idx, const Number *gsym, const Number *src, Number *dst, uint32_t cell_begin,
                                uint32_t cell_end, const double *N, const double *D, cudaStream_t stream);

}  // namespace mfg
