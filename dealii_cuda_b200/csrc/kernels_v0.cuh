// kernels_v0.cuh -- generic sum-factorised Laplace cell kernel ("column"
// kernel) for dim = 2,3, degree 1..8, float/double.
//
// Replaces apply_kernel_shmem<LocalOperator> (matrix_free_gpu.h:318-341) and the
// device code it inlines: FEEvaluationGpu::read_dof_values / evaluate /
// submit_gradient / integrate / distribute_local_to_global (fee_gpu.cuh:197-365),
// TensorOpsShmem::grad_at_quad_pts / quad_int_grad (tensor_ops.cuh:179-261) and
// LocalOperator::quad_operation (laplace_operator_gpu.h:257-260).
//
// Differences in formulation (same bilinear form, fewer flops and no per-FMA
// shared-memory operand):
//  * collocation form: u is first interpolated to the Gauss points (N per
//    direction), gradients are taken with the (p+1)x(p+1) derivative matrix of
//    the Lagrange basis through the Gauss points: 2*dim + 2*dim contractions
//    instead of the reference's 2*dim*dim;
//  * one thread owns a whole 1-D line of the cell tensor in registers, so a
//    contraction is n*n register FMAs whose matrix operand comes from the
//    kernel-parameter constant bank; directions are changed by transposing the
//    cell tensor through shared memory (1 store + 1 load per entry) instead of
//    reading one shared operand per FMA (tensor_ops.cuh:90-103);
//  * inverse Jacobian, JxW and coefficient are merged at setup into one weight
//    per quadrature point, cw = a(x_q) * (J^-1)^2 * JxW_q  (uniform meshes);
//  * constrained DoFs carry bit 31 in the index array: they are read as 0 and
//    never written, which fuses ConstraintHandlerGpu::save_constrained_values /
//    load_and_add_constrained_values (laplace_operator_gpu.h:293,302) into the
//    gather / scatter;
//  * all barriers are executed by every thread of the block (the reference
//    executes them under `if(cell<n_cells)`, matrix_free_gpu.h:338).
#pragma once
#include "common.cuh"

namespace mfg {

template <typename Number, int n> struct ShapeMats
{
  Number N[n * n];  // N[i*n+q] = phi_i(x_q)
  Number D[n * n];  // D[a*n+q] = l_a'(x_q)  (collocation derivative)
};

constexpr uint32_t CONSTRAINED_BIT = 0x80000000u;

// atomicAddWrapper (atomic.cuh:11-32): native RED.ADD.F64 / RED.ADD.F32 on sm_100a
template <typename Number> __device__ __forceinline__ void red_add(Number *addr, Number v) { atomicAdd(addr, v); }

// offset of register entry e of thread t inside the lexicographic cell tensor
// when direction r is the one held in registers
template <int dim, int n, int r> __device__ __forceinline__ int lay(int t, int e)
{
  if (dim == 3)
    {
      if (r == 2) return t + n * n * e;
      if (r == 1) return (t % n) + n * e + n * n * (t / n);
      return e + n * t;
    }
  else
    {
      if (r == 1) return t + n * e;
      return e + n * t;
    }
}

// out[q] = sum_k M[k*n+q] in[k]  (TR=false);   out[k] = sum_q M[k*n+q] in[q]  (TR=true)
template <int n, bool TR, typename Number>
__device__ __forceinline__ void apply1d(const Number *__restrict__ M, const Number (&in)[n], Number (&out)[n])
{
  // outer-product order: the n accumulators advance together, so consecutive FMAs are independent
#pragma unroll
  for (int a = 0; a < n; ++a) out[a] = (TR ? M[a * n + 0] : M[0 * n + a]) * in[0];
#pragma unroll
  for (int b = 1; b < n; ++b)
#pragma unroll
    for (int a = 0; a < n; ++a) out[a] = fma((TR ? M[a * n + b] : M[b * n + a]), in[b], out[a]);
}

template <int dim, int n, int r, typename Number>
__device__ __forceinline__ void st_lay(Number *buf, int t, const Number (&v)[n])
{
#pragma unroll
  for (int e = 0; e < n; ++e) buf[lay<dim, n, r>(t, e)] = v[e];
}
template <int dim, int n, int r, typename Number>
__device__ __forceinline__ void ld_lay(const Number *buf, int t, Number (&v)[n])
{
#pragma unroll
  for (int e = 0; e < n; ++e) v[e] = buf[lay<dim, n, r>(t, e)];
}

constexpr int v0_threads_per_cell(int dim, int n) { return dim == 3 ? n * n : n; }
constexpr int v0_cells_per_block(int dim, int n)
{
  // fill ~128 (or ~256) threads with as few idle lanes as possible
  return dim == 3 ? (n == 2 ? 32 : n == 3 ? 14 : n == 4 ? 8 : n == 5 ? 5 : n == 6 ? 7 : n == 7 ? 5 : n == 8 ? 2 : 3)
                  : (n == 2 ? 64 : n == 3 ? 42 : n == 4 ? 32 : n == 5 ? 25 : n == 6 ? 21 : n == 7 ? 18 : n == 8 ? 16 : 14);
}
constexpr int v0_block_threads(int dim, int n) { return ((v0_threads_per_cell(dim, n) * v0_cells_per_block(dim, n) + 31) / 32) * 32; }
template <typename Number> constexpr size_t v0_smem_bytes(int dim, int n)
{
  return sizeof(Number) * 3 * ipow(n, dim) * v0_cells_per_block(dim, n);
}

// q-point phase for one direction r (register axis): R (+)= D^T ( w .* (D G) )
template <int n, bool FIRST, typename Number>
__device__ __forceinline__ void qphase(const Number *__restrict__ D, const Number (&G)[n], const Number (&w)[n], Number (&R)[n])
{
  Number g[n], t[n];
  apply1d<n, false>(D, G, g);
#pragma unroll
  for (int e = 0; e < n; ++e) g[e] *= w[e];
  apply1d<n, true>(D, g, t);
#pragma unroll
  for (int e = 0; e < n; ++e) R[e] = FIRST ? t[e] : R[e] + t[e];
}

template <typename Number, int n> struct HangingMat
{
  Number W[n * n];  // W[k*n+i] = phi_i(xi_k/2), first child; second child by index reversal (hanging_nodes.cuh:665-682)
};

// resolve_hanging_nodes_shmem<dim,fe_degree,transpose> (hanging_nodes.cuh:617-778) for the column layout:
// the cell tensor is staged in shared memory, one sweep per direction (x, y[, z]) ping-ponging between two
// buffers; every point on a constrained face / edge is replaced by the 1-D interpolation along the sweep
// direction.  `u` enters and leaves in the gather/scatter layout (register axis dim-1).  Executed by every
// thread of the block (uniform barriers); cells with mask 0 just copy.
template <int dim, int n, bool TR, typename Number>
__device__ __forceinline__ void hn_resolve(Number (&u)[n], Number *b0, Number *b1, const int t, const unsigned mask,
                                           const bool in_blk, const Number *__restrict__ W)
{
  constexpr int RL = dim - 1, p = n - 1;
  if (in_blk) st_lay<dim, n, RL>(b0, t, u);
  __syncthreads();
  Number *in = b0, *out = b1;
#pragma unroll
  for (int d = 0; d < dim; ++d)
    {
      if (in_blk)
        {
#pragma unroll
          for (int e = 0; e < n; ++e)
            {
              int c[3];
              if (dim == 3) { c[0] = t % n; c[1] = t / n; c[2] = e; }
              else { c[0] = t; c[1] = e; c[2] = 0; }
              bool flag;
              if (dim == 2)
                {
                  const int a = 1 - d;
                  const bool on = (mask & (1u << a)) ? (c[a] == 0) : (c[a] == p);
                  flag = (mask & (8u << a)) && on;
                }
              else
                {
                  const int f1 = (d + 1) % 3, f2 = (d + 2) % 3;
                  const bool on1 = (mask & (1u << f1)) ? (c[f1] == 0) : (c[f1] == p);
                  const bool on2 = (mask & (1u << f2)) ? (c[f2] == 0) : (c[f2] == p);
                  // edge bit by the direction the edge runs along: x -> EDGE_YZ (1<<7), y -> EDGE_ZX (1<<8), z -> EDGE_XY (1<<6)
                  const unsigned ebit = d == 0 ? (1u << 7) : d == 1 ? (1u << 8) : (1u << 6);
                  flag = ((mask & (8u << f1)) && on1) || ((mask & (8u << f2)) && on2) || ((mask & ebit) && on1 && on2);
                }
              const int stride = d == 0 ? 1 : d == 1 ? n : n * n;
              const int here   = c[0] + n * (c[1] + n * c[2]);
              Number val = in[here];
              if (flag)
                {
                  const int  k     = c[d];
                  const int  base  = here - k * stride;
                  const bool first = mask & (1u << d);
                  Number acc = 0;
                  for (int i = 0; i < n; ++i)
                    {
                      const Number w = first ? (TR ? W[i * n + k] : W[k * n + i]) : (TR ? W[(p - i) * n + p - k] : W[(p - k) * n + p - i]);
                      acc += w * in[base + i * stride];
                    }
                  val = acc;
                }
              out[here] = val;
            }
        }
      __syncthreads();
      Number *tmp = in; in = out; out = tmp;
    }
  if (in_blk) ld_lay<dim, n, RL>(in, t, u);
  __syncthreads();  // the staging buffers are reused by the caller
}

template <int dim, int n, typename Number, bool ATOMIC, bool HN>
__global__ void __launch_bounds__(v0_block_threads(dim, n))
laplace_cell_v0(const uint32_t *__restrict__ idx, const Number *__restrict__ cw, const Number *__restrict__ src,
                Number *__restrict__ dst, const uint32_t cell_begin, const uint32_t cell_end,
                const __grid_constant__ ShapeMats<Number, n> sh, const uint32_t *__restrict__ hn_mask,
                const __grid_constant__ HangingMat<Number, n> hm)
{
  constexpr int TPC = v0_threads_per_cell(dim, n);
  constexpr int CPB = v0_cells_per_block(dim, n);
  constexpr int NPC = ipow(n, dim);
  constexpr int RL  = dim - 1;  // register axis of the (coalesced) gather/scatter layout
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *smem = reinterpret_cast<Number *>(smem_raw);

  const int  lc     = threadIdx.x / TPC;
  const int  t      = threadIdx.x % TPC;
  const bool in_blk = lc < CPB;
  const uint32_t cell = cell_begin + blockIdx.x * CPB + lc;
  const bool active = in_blk && cell < cell_end;
  Number *bufA = smem + (size_t)(in_blk ? lc : 0) * 3 * NPC;
  Number *bufB = bufA + NPC;
  Number *bufW = bufB + NPC;

  uint32_t id[n];
  Number   u[n], w[n], v[n];
  // ---- read_dof_values (fee_gpu.cuh:323-338) + coefficient row -------------
  if (active)
    {
      const uint32_t *row  = idx + (size_t)cell * NPC;
      const Number   *crow = cw + (size_t)cell * NPC;
#pragma unroll
      for (int e = 0; e < n; ++e) id[e] = row[lay<dim, n, RL>(t, e)];
#pragma unroll
      for (int e = 0; e < n; ++e) w[e] = crow[lay<dim, n, RL>(t, e)];
#pragma unroll
      for (int e = 0; e < n; ++e) u[e] = (id[e] & CONSTRAINED_BIT) ? Number(0) : __ldg(src + id[e]);
    }
  else
    {
#pragma unroll
      for (int e = 0; e < n; ++e) { id[e] = CONSTRAINED_BIT; w[e] = 0; u[e] = 0; }
    }
  if (in_blk) st_lay<dim, n, RL>(bufW, t, w);
  unsigned mask = 0;
  if (HN)
    {
      // hanging nodes: interpolate the coarse-face values gathered through the rewritten map (fee_gpu.cuh:333-335)
      if (active) mask = hn_mask[cell];
      hn_resolve<dim, n, false>(u, bufA, bufB, t, mask, in_blk, hm.W);
    }

  // ---- interpolate to Gauss points: N along dim-1, ..., 0 -------------------
  apply1d<n, false>(sh.N, u, v);
  if (dim == 3)
    {
      if (in_blk) st_lay<dim, n, 2>(bufA, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 1>(bufA, t, u);
      apply1d<n, false>(sh.N, u, v);
      if (in_blk) st_lay<dim, n, 1>(bufB, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 0>(bufB, t, u);
    }
  else
    {
      if (in_blk) st_lay<dim, n, 1>(bufB, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 0>(bufB, t, u);
    }
  Number G[n], R[n];
  apply1d<n, false>(sh.N, u, G);  // G: values at quadrature points, x-lines in registers

  // ---- quadrature-point phase: R = sum_d D_d^T ( cw .* D_d G ) --------------
  // (evaluate gradients, quad_operation, first half of integrate)
  if (in_blk) ld_lay<dim, n, 0>(bufW, t, w);  // bufW was written before the first barrier
  qphase<n, true>(sh.D, G, w, R);
  if (in_blk) { st_lay<dim, n, 0>(bufA, t, G); st_lay<dim, n, 0>(bufB, t, R); }
  __syncthreads();
  if (in_blk) { ld_lay<dim, n, 1>(bufA, t, G); ld_lay<dim, n, 1>(bufB, t, R); ld_lay<dim, n, 1>(bufW, t, w); }
  qphase<n, false>(sh.D, G, w, R);
  if (dim == 3)
    {
      if (in_blk) st_lay<dim, n, 1>(bufB, t, R);  // same entries this thread just read
      __syncthreads();
      if (in_blk) { ld_lay<dim, n, 2>(bufA, t, G); ld_lay<dim, n, 2>(bufB, t, R); ld_lay<dim, n, 2>(bufW, t, w); }
      qphase<n, false>(sh.D, G, w, R);
    }

  // ---- integrate: N^T along dim-1, ..., 0 ; back to the gather layout -------
  apply1d<n, true>(sh.N, R, v);
  if (dim == 3)
    {
      if (in_blk) st_lay<dim, n, 2>(bufA, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 1>(bufA, t, u);
      apply1d<n, true>(sh.N, u, v);
      if (in_blk) st_lay<dim, n, 1>(bufB, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 0>(bufB, t, u);
      apply1d<n, true>(sh.N, u, v);
      if (in_blk) st_lay<dim, n, 0>(bufA, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 2>(bufA, t, v);
    }
  else
    {
      if (in_blk) st_lay<dim, n, 1>(bufA, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 0>(bufA, t, u);
      apply1d<n, true>(sh.N, u, v);
      if (in_blk) st_lay<dim, n, 0>(bufB, t, v);
      __syncthreads();
      if (in_blk) ld_lay<dim, n, 1>(bufB, t, v);
    }

  // ---- distribute_local_to_global (fee_gpu.cuh:346-365) ---------------------
  if (HN)
    {
      __syncthreads();  // everybody has read the result before the buffers are recycled
      hn_resolve<dim, n, true>(v, dim == 3 ? bufB : bufA, dim == 3 ? bufA : bufB, t, mask, in_blk, hm.W);
    }
#pragma unroll
  for (int e = 0; e < n; ++e)
    if (!(id[e] & CONSTRAINED_BIT))
      {
        if (ATOMIC) red_add(dst + id[e], v[e]);
        else dst[id[e]] += v[e];
      }
}

// dst[i] = constrained(i) ? src[i] : 0   -- fuses `dst = 0` (vec_init, gpu_vec.cu:281-291)
// with the constrained-row identity of load_and_add_constrained_values
// (constraint_handler_gpu.cu:277-289) for vmult
template <typename Number>
__global__ void vmult_prepare(Number *__restrict__ dst, const Number *__restrict__ src, const uint32_t *__restrict__ cbits, size_t n)
{
  // one 16-byte store per thread and iteration; a 32-bit mask word covers 32 DoFs
  constexpr int V = 16 / (int)sizeof(Number);
  struct alignas(16) Vec { Number v[V]; };
  const size_t nvec   = n / V;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll 4
  for (size_t iv = (size_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += stride)
    {
      const size_t   i    = iv * V;
      const uint32_t bits = (cbits[i >> 5] >> (i & 31)) & ((1u << V) - 1u);
      Vec out;
#pragma unroll
      for (int k = 0; k < V; ++k) out.v[k] = Number(0);
      if (bits)
        {
#pragma unroll
          for (int k = 0; k < V; ++k)
            if ((bits >> k) & 1u) out.v[k] = src[i + k];
        }
      *reinterpret_cast<Vec *>(dst + i) = out;
    }
  if (blockIdx.x == 0 && threadIdx.x < n - nvec * V)
    {
      const size_t i = nvec * V + threadIdx.x;
      dst[i] = ((cbits[i >> 5] >> (i & 31)) & 1u) ? src[i] : Number(0);
    }
}

// dst[c] = src[c] over the constrained list (vmult after a memset of dst)
template <typename Number>
__global__ void constrained_copy(Number *__restrict__ dst, const Number *__restrict__ src, const uint32_t *__restrict__ list, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t c = list[i]; dst[c] = src[c]; }
}

// dst[c] += src[c] over the constrained list (vmult_add)
template <typename Number>
__global__ void constrained_add(Number *__restrict__ dst, const Number *__restrict__ src, const uint32_t *__restrict__ list, size_t n)
{
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const uint32_t c = list[i]; dst[c] += src[c]; }
}

// host-side launcher, explicitly instantiated per (dim, dtype) in kernels_v0_inst.cu
template <int dim, typename Number>
void launch_laplace_v0_dim(int degree, bool atomic, const uint32_t *idx, const Number *cw, const Number *src, Number *dst,
                           uint32_t cell_begin, uint32_t cell_end, const double *N, const double *D, cudaStream_t stream,
                           const uint32_t *hn_mask = nullptr, const double *hn_weights = nullptr);

}  // namespace mfg
