// fee_gpu.cuh -- header-only generic path: FEEvaluationGpu + cell_loop for USER-WRITTEN cell functors, compiled by nvcc
// into the user's translation unit exactly like the reference's templates (matrix_free_gpu.h:318-380, fee_gpu.cuh:89-365).
// The Laplace LocalOperator is special-cased to the precompiled sm_100a kernels of libmfgpu.so (LaplaceOperatorGpu);
// everything else (mass operators, right-hand sides, other PDEs) goes through this file.
//
//   template <int dim, int fe_degree, typename Number> struct MassOp   // the reference's functor protocol
//   {                                                                  // (laplace_operator_gpu.h:247-282)
//     typedef dealii_cuda_b200::FEEvaluationGpu<dim, fe_degree, Number> FEE;
//     __device__ void cell_apply(Number *dst, const Number *src, const typename FEE::data_type *gpu_data, const unsigned int cell,
//                                dealii_cuda_b200::SharedData<dim, Number> *shdata) const
//     {
//       FEE phi(cell, gpu_data, shdata);
//       phi.read_dof_values(src);
//       phi.evaluate(true, false);
//       phi.apply_quad_point_operations(this);   // quad_operation(&phi, q) on every quadrature point
//       phi.integrate(true, false);
//       phi.distribute_local_to_global(dst);
//     }
//     __device__ void quad_operation(FEE *phi, const unsigned int q) const { phi->submit_value(phi->get_value(q), q); }
//   };
//   dealii_cuda_b200::cell_loop<dim, fe_degree>(matrix_free, dst, src, MassOp<dim, fe_degree, Number>());
//
// Thread layout: one thread per local DoF / quadrature point (the reference's "parallel in element" scheme), several
// cells per CTA, values / gradients / two scratch planes of a cell in shared memory.  Every method is called by all
// threads of the CTA (they contain barriers).  No constraint handling here, as in the reference's cell_loop
// (ConstraintHandlerGpu is applied around it); on meshes with hanging nodes read_dof_values / distribute_local_to_global
// interpolate on the cells that carry a constraint mask, like the reference's resolve_hanging_nodes_shmem (fee_gpu.cuh:333-351).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../mfgpu.h"
#include "matrix_free_gpu.h"

namespace dealii_cuda_b200 {

template <int dim, typename Number> struct GpuArray
{
  Number arr[dim];
  __host__ __device__ Number       &operator[](int i) { return arr[i]; }
  __host__ __device__ const Number &operator[](int i) const { return arr[i]; }
};

// MatrixFreeGpu::GpuData (matrix_free_gpu.h:98-121), passed to the kernel by value
template <int dim, typename Number> struct GpuData
{
  const uint32_t *loc2glob;           // [n_cells][n_local_dofs]
  const Number   *JxW;                // [n_cells][n_q_points]
  const Number   *inv_jac;            // uniform: [n_cells]; general: [n_cells][n_q_points][dim][dim]
  const Number   *quadrature_points;  // [n_cells][n_q_points][dim] or nullptr
  const Number   *shape_values;       // [i*n+q] phi_i(x_q)   (set by the kernel: points into its parameter block)
  const Number   *shape_gradients;    // [i*n+q] phi_i'(x_q)
  uint32_t        n_cells;            // cells below this bound are valid in the current launch
  int             general, use_coloring;
  // hanging nodes (-DMATRIX_FREE_HANGING_NODES): non-null only in the launch over the cells that carry a constraint mask
  const uint32_t *constraint_mask;    // [n_cells] 9-bit masks of HangingNodes::setup_constraints (hanging_nodes.cuh:38-50)
  const Number   *hanging_weights;    // [k*n+i] = phi_i(xi_k / 2) (setup_constraint_weights, hanging_nodes.cuh:580-598)
};

template <int dim, typename Number> struct SharedData
{
  Number *values;
  Number *gradients[dim];
  Number *scratch[2];
};

template <typename Number, int n> struct ShapeTables { Number val[n * n], grad[n * n], hang[n * n]; };

template <int dim, int fe_degree, typename Number> class FEEvaluationGpu
{
public:
  typedef Number                number_type;
  typedef GpuData<dim, Number>  data_type;
  typedef GpuArray<dim, Number> gradient_type;
  static const unsigned int n_dofs_1d = fe_degree + 1, n_q_points_1d = fe_degree + 1;
  static const unsigned int n_local_dofs = dim == 2 ? n_dofs_1d * n_dofs_1d : n_dofs_1d * n_dofs_1d * n_dofs_1d;
  static const unsigned int n_q_points = n_local_dofs;

  // cell id, the data cache, the cell's shared scratch (fee_gpu.cuh:172-194).  A cell id beyond the launch (last CTA)
  // is mapped to the last valid cell and never writes, so that all threads can keep calling the collective methods.
  __device__ FEEvaluationGpu(const unsigned int cellid, const data_type *data, SharedData<dim, Number> *shdata)
    : valid(cellid < data->n_cells), cell(cellid < data->n_cells ? cellid : data->n_cells - 1), d(data), values(shdata->values),
      t(threadIdx.x % n_local_dofs)
  {
    for (int k = 0; k < dim; ++k) gradients[k] = shdata->gradients[k];
    scratch[0] = shdata->scratch[0];
    scratch[1] = shdata->scratch[1];
  }

  // values[i] = src[loc2glob[cell][i]]   (:323-338)
  __device__ void read_dof_values(const Number *src)
  {
    values[t] = src[d->loc2glob[(size_t)cell * n_local_dofs + t] & 0x7fffffffu];
    __syncthreads();
    if (d->constraint_mask != nullptr) resolve_hanging_nodes(false);   // resolve_hanging_nodes_shmem (fee_gpu.cuh:333-335)
  }
  // dst[loc2glob[cell][i]] += values[i]: plain if the cells of a launch are conflict free (coloring), else atomic (:346-365)
  __device__ void distribute_local_to_global(Number *dst)
  {
    if (d->constraint_mask != nullptr) resolve_hanging_nodes(true);    // the transposed interpolation (fee_gpu.cuh:349-351)
    if (!valid) return;
    const uint32_t g = d->loc2glob[(size_t)cell * n_local_dofs + t] & 0x7fffffffu;
    if (d->use_coloring) dst[g] += values[t];
    else atomicAdd(dst + g, values[t]);
  }
  // values and / or reference-cell gradients at the quadrature points from the nodal values (:197-210)
  __device__ void evaluate(const bool evaluate_val, const bool evaluate_grad)
  {
    if (evaluate_grad)
      for (int c = 0; c < dim; ++c)
        {
          const Number *in = values;
          for (int dir = 0; dir < dim; ++dir)
            {
              Number *out = dir == dim - 1 ? gradients[c] : scratch[dir % 2];
              out[t] = contract(dir == c ? d->shape_gradients : d->shape_values, in, dir, false);
              __syncthreads();
              in = out;
            }
        }
    if (evaluate_val)
      {
        const Number *in = values;
        for (int dir = 0; dir < dim; ++dir)
          {
            Number *out = scratch[dir % 2];
            out[t] = contract(d->shape_values, in, dir, false);
            __syncthreads();
            in = out;
          }
        values[t] = in[t];
        __syncthreads();
      }
  }

  __device__ Number        get_value(const unsigned int q) const { return values[q]; }
  __device__ gradient_type get_gradient(const unsigned int q) const   // grad_x = K^T grad_xi  (:219-246)
  {
    gradient_type g;
    if (!d->general)
      {
        const Number J = d->inv_jac[cell];
        for (int a = 0; a < dim; ++a) g[a] = J * gradients[a][q];
      }
    else
      {
        const Number *K = d->inv_jac + ((size_t)cell * n_q_points + q) * dim * dim;
        for (int a = 0; a < dim; ++a)
          {
            Number s = 0;
            for (int b = 0; b < dim; ++b) s += K[b * dim + a] * gradients[b][q];
            g[a] = s;
          }
      }
    return g;
  }
  __device__ void submit_dof_value(const Number &val, const unsigned int i) { values[i] = val; }
  __device__ void submit_value(const Number &val, const unsigned int q) { values[q] = val * d->JxW[(size_t)cell * n_q_points + q]; }  // (:255-259)
  __device__ void submit_gradient(const gradient_type &grad, const unsigned int q)                                                 // (:261-284)
  {
    const Number jxw = d->JxW[(size_t)cell * n_q_points + q];
    if (!d->general)
      {
        const Number J = d->inv_jac[cell];
        for (int a = 0; a < dim; ++a) gradients[a][q] = grad[a] * J * jxw;
      }
    else
      {
        const Number *K = d->inv_jac + ((size_t)cell * n_q_points + q) * dim * dim;
        for (int a = 0; a < dim; ++a)
          {
            Number s = 0;
            for (int b = 0; b < dim; ++b) s += K[a * dim + b] * grad[b];
            gradients[a][q] = s * jxw;
          }
      }
  }
  __device__ gradient_type get_quadrature_point(const unsigned int q) const
  {
    gradient_type p;
    for (int a = 0; a < dim; ++a) p[a] = d->quadrature_points[((size_t)cell * n_q_points + q) * dim + a];
    return p;
  }
  // row of per-quadrature-point user arrays [n_cells][n_q_points] (kernel cell order) for local point q
  __device__ unsigned int get_global_q(const unsigned int q) const { return cell * n_q_points + q; }

  // test with all basis functions / their gradients and sum over the quadrature points; result in the nodal values (:286-303)
  __device__ void integrate(const bool integrate_val, const bool integrate_grad)
  {
    __syncthreads();  // the submits of all threads are visible
    Number acc = 0;
    if (integrate_val)
      {
        const Number *in = values;
        for (int dir = 0; dir < dim; ++dir)
          {
            Number *out = scratch[dir % 2];
            out[t] = contract(d->shape_values, in, dir, true);
            __syncthreads();
            in = out;
          }
        acc += in[t];
        __syncthreads();
      }
    if (integrate_grad)
      for (int c = 0; c < dim; ++c)
        {
          const Number *in = gradients[c];
          for (int dir = 0; dir < dim; ++dir)
            {
              Number *out = scratch[dir % 2];
              out[t] = contract(dir == c ? d->shape_gradients : d->shape_values, in, dir, true);
              __syncthreads();
              in = out;
            }
          acc += in[t];
          __syncthreads();
        }
    values[t] = acc;
    __syncthreads();
  }

  // lop->quad_operation(this, q) on every quadrature point, one per thread (:306-316)
  template <typename LocOp> __device__ void apply_quad_point_operations(const LocOp *lop)
  {
    lop->quad_operation(this, t);
    __syncthreads();
  }

private:
  // resolve_hanging_nodes_shmem (hanging_nodes.cuh:617-778) on the cell's nodal values, one thread per local DoF: a sweep along
  // every direction replaces the values on constrained faces / edges by the interpolation of the coarse neighbour's values
  // (which loc2glob put there), W[k][i] = phi_i(xi_k / 2) for a child on the lower side of the direction, the mirrored table for
  // the upper one; transpose = the adjoint, applied before the scatter.  Called by all threads of the CTA (barriers).
  __device__ void resolve_hanging_nodes(const bool transpose)
  {
    constexpr int      n = fe_degree + 1, p = fe_degree;
    const unsigned int mask = d->constraint_mask[cell];
    const Number      *W = d->hanging_weights;
    int                idx[3] = {(int)(t % n), (int)((t / n) % n), dim == 3 ? (int)(t / (n * n)) : 0};
    for (int dir = 0; dir < dim; ++dir)
      {
        bool flag;
        if (dim == 2)
          {
            const int  a = 1 - dir;
            const bool on = (mask & (1u << a)) ? idx[a] == 0 : idx[a] == p;
            flag = (mask & (8u << a)) && on;
          }
        else
          {
            const int      f1 = (dir + 1) % 3, f2 = (dir + 2) % 3;
            const bool     on1 = (mask & (1u << f1)) ? idx[f1] == 0 : idx[f1] == p, on2 = (mask & (1u << f2)) ? idx[f2] == 0 : idx[f2] == p;
            const unsigned edge_bit = dir == 0 ? (1u << 7) : dir == 1 ? (1u << 8) : (1u << 6);   // edge along x: YZ, y: ZX, z: XY (:48-50)
            flag = ((mask & (8u << f1)) && on1) || ((mask & (8u << f2)) && on2) || ((mask & edge_bit) && on1 && on2);
          }
        Number val = values[t];
        if (flag)
          {
            const int  stride = dir == 0 ? 1 : dir == 1 ? n : n * n, k = idx[dir], base = (int)t - k * stride;
            const bool lower = (mask & (1u << dir)) != 0;
            Number     acc = 0;
            for (int i = 0; i < n; ++i)
              {
                const Number w = lower ? (transpose ? W[i * n + k] : W[k * n + i]) : (transpose ? W[(p - i) * n + (p - k)] : W[(p - k) * n + (p - i)]);
                acc += w * values[base + i * stride];
              }
            val = acc;
          }
        __syncthreads();
        values[t] = val;
        __syncthreads();
      }
  }

  // out(.., a, ..) = sum_b M[b*n+a] in(.., b, ..) along direction dir   (tr: M[a*n+b])
  __device__ Number contract(const Number *M, const Number *in, const int dir, const bool tr) const
  {
    constexpr int n = fe_degree + 1;
    const int stride = dir == 0 ? 1 : dir == 1 ? n : n * n;
    const int a = (t / stride) % n, base = t - a * stride;
    Number acc = 0;
    for (int b = 0; b < n; ++b) acc += (tr ? M[a * n + b] : M[b * n + a]) * in[base + b * stride];
    return acc;
  }

  const bool         valid;
  const unsigned int cell;
  const data_type   *d;
  Number            *values;
  Number            *gradients[dim];
  Number            *scratch[2];
  const unsigned int t;
};

// apply_kernel_shmem<LocOp> (matrix_free_gpu.h:318-341): carve the shared memory of the CTA's cells, call loc_op.cell_apply
template <typename LocOp, int dim, int fe_degree, typename Number>
__global__ void apply_kernel_shmem(Number *dst, const Number *src, const LocOp loc_op, const GpuData<dim, Number> gpu_data,
                                   const __grid_constant__ ShapeTables<Number, fe_degree + 1> tab, const unsigned int cell_begin)
{
  constexpr unsigned int npc = FEEvaluationGpu<dim, fe_degree, Number>::n_local_dofs;
  extern __shared__ __align__(16) unsigned char fee_smem_raw[];
  Number *base = reinterpret_cast<Number *>(fee_smem_raw) + (size_t)(threadIdx.x / npc) * (3 + dim) * npc;
  SharedData<dim, Number> sh;
  sh.values = base;
  for (int k = 0; k < dim; ++k) sh.gradients[k] = base + (1 + k) * npc;
  sh.scratch[0] = base + (1 + dim) * npc;
  sh.scratch[1] = base + (2 + dim) * npc;
  GpuData<dim, Number> gd = gpu_data;
  gd.shape_values = tab.val;
  gd.shape_gradients = tab.grad;
  gd.hanging_weights = tab.hang;
  const unsigned int cell = cell_begin + blockIdx.x * (blockDim.x / npc) + threadIdx.x / npc;
  loc_op.cell_apply(dst, src, &gd, cell, &sh);
}

// apply_kernel_shmem<LocOp>(dst, loc_op, gpu_data) (matrix_free_gpu.h:343-365): the variant without a source vector,
// loc_op.cell_apply(dst, gpu_data, cell, shdata) -- the reference uses it for compute_diagonal (laplace_operator_gpu.h:413-414)
template <typename LocOp, int dim, int fe_degree, typename Number>
__global__ void apply_kernel_shmem_dst(Number *dst, const LocOp loc_op, const GpuData<dim, Number> gpu_data,
                                       const __grid_constant__ ShapeTables<Number, fe_degree + 1> tab, const unsigned int cell_begin)
{
  constexpr unsigned int npc = FEEvaluationGpu<dim, fe_degree, Number>::n_local_dofs;
  extern __shared__ __align__(16) unsigned char fee_smem_raw[];
  Number *base = reinterpret_cast<Number *>(fee_smem_raw) + (size_t)(threadIdx.x / npc) * (3 + dim) * npc;
  SharedData<dim, Number> sh;
  sh.values = base;
  for (int k = 0; k < dim; ++k) sh.gradients[k] = base + (1 + k) * npc;
  sh.scratch[0] = base + (1 + dim) * npc;
  sh.scratch[1] = base + (2 + dim) * npc;
  GpuData<dim, Number> gd = gpu_data;
  gd.shape_values = tab.val;
  gd.shape_gradients = tab.grad;
  gd.hanging_weights = tab.hang;
  const unsigned int cell = cell_begin + blockIdx.x * (blockDim.x / npc) + threadIdx.x / npc;
  loc_op.cell_apply(dst, &gd, cell, &sh);
}

// cell_eval_kernel<dim,Number,Op> (matrix_free_gpu.h:397-410): Op::eval(row of vec, quadrature points of the cell).  The
// reference runs one thread per cell, serial over the quadrature points; Op::eval keeps that signature (it loops over
// n_q_points itself), rows are [n_q_points] long here (the reference pads them to `rowlength`).
template <int dim, typename Number, typename Op>
__global__ void cell_eval_kernel(Number *vec, const GpuData<dim, Number> gpu_data, const unsigned int n_q_points)
{
  const unsigned int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < gpu_data.n_cells) Op::eval(vec + (size_t)cell * n_q_points, gpu_data.quadrature_points + (size_t)cell * n_q_points * dim);
}

inline void fee_check(int rc, const char *what)
{
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + mfg_last_error());
}
inline void fee_check_cuda(cudaError_t e, const char *what)
{
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// the kernel's parameter block of 1-D tables: shape values / gradients of the MatrixFreeGpu object, hanging-node weights of its degree
template <typename Number, int n> void fill_shape_tables(const mfg_gpu_data &g, ShapeTables<Number, n> &tab)
{
  double W[n * n];
  fee_check(mfg_hanging_node_weights(n - 1, W), "mfg_hanging_node_weights");
  for (int i = 0; i < n * n; ++i)
    {
      tab.val[i] = (Number)g.shape_values[i];
      tab.grad[i] = (Number)g.shape_gradients[i];
      tab.hang[i] = (Number)W[i];
    }
}

// MatrixFreeGpu::cell_loop(dst, src, loc_op) (matrix_free_gpu.h:369-380) on the raw handle and device pointers: one launch
// per color.  `Number` must be the operator's dtype.
template <int dim, int fe_degree, typename Number, typename LocOp>
void cell_loop(mfg_mf *mf, Number *dst_dev, const Number *src_dev, const LocOp &loc_op)
{
  mfg_gpu_data g;
  fee_check(mfg_mf_get_gpu_data(mf, &g), "mfg_mf_get_gpu_data");
  if (g.dim != dim || g.degree != fe_degree) throw std::runtime_error("cell_loop: dim / fe_degree differ from the MatrixFreeGpu object");
  if ((g.dtype == MFG_F64) != (sizeof(Number) == 8)) throw std::runtime_error("cell_loop: Number differs from the MatrixFreeGpu dtype");
  if (g.constraint_mask != nullptr && g.use_coloring) throw std::runtime_error("cell_loop: hanging nodes need the atomic scatter");
  constexpr unsigned int n = fe_degree + 1, npc = dim == 2 ? n * n : n * n * n;
  static_assert(npc <= 1024, "one thread per local DoF");
  GpuData<dim, Number> gd;
  gd.loc2glob = g.loc2glob;
  gd.JxW = static_cast<const Number *>(g.JxW);
  gd.inv_jac = static_cast<const Number *>(g.inv_jac);
  gd.quadrature_points = static_cast<const Number *>(g.quadrature_points);
  gd.shape_values = gd.shape_gradients = nullptr;
  gd.general = g.general;
  gd.use_coloring = g.use_coloring;
  gd.constraint_mask = nullptr;
  gd.hanging_weights = nullptr;
  ShapeTables<Number, n> tab;
  fill_shape_tables<Number, n>(g, tab);
  const unsigned int cpb = npc >= 128 ? 1 : 128 / npc;
  const size_t       smem = (size_t)cpb * (3 + dim) * npc * sizeof(Number);
  auto kern = apply_kernel_shmem<LocOp, dim, fe_degree, Number>;
  if (smem > 48 * 1024) fee_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
  cudaStream_t st = static_cast<cudaStream_t>(g.cuda_stream);
  for (uint32_t c = 0; c < g.n_colors; ++c)
    {
      const uint32_t c0 = g.color_offsets[c], c1 = g.color_offsets[c + 1];
      if (c1 <= c0) continue;
      gd.n_cells = c1;
      kern<<<(c1 - c0 + cpb - 1) / cpb, cpb * npc, smem, st>>>(dst_dev, src_dev, loc_op, gd, tab, c0);
      fee_check_cuda(cudaGetLastError(), "cell_loop launch");
    }
  // cells with hanging-node constraints (sorted behind the plain ones): read_dof_values / distribute_local_to_global interpolate
  if (g.constraint_mask != nullptr && g.n_plain_cells < g.n_cells)
    {
      gd.constraint_mask = g.constraint_mask;
      gd.n_cells = g.n_cells;
      kern<<<(g.n_cells - g.n_plain_cells + cpb - 1) / cpb, cpb * npc, smem, st>>>(dst_dev, src_dev, loc_op, gd, tab, g.n_plain_cells);
      fee_check_cuda(cudaGetLastError(), "cell_loop launch (hanging-node cells)");
    }
}

// MatrixFreeGpu::cell_loop(dst, loc_op) (matrix_free_gpu.h:382-393): no source vector, loc_op.cell_apply(dst, gpu_data, cell, shdata)
template <int dim, int fe_degree, typename Number, typename LocOp>
void cell_loop(mfg_mf *mf, Number *dst_dev, const LocOp &loc_op)
{
  mfg_gpu_data g;
  fee_check(mfg_mf_get_gpu_data(mf, &g), "mfg_mf_get_gpu_data");
  if (g.dim != dim || g.degree != fe_degree) throw std::runtime_error("cell_loop: dim / fe_degree differ from the MatrixFreeGpu object");
  if ((g.dtype == MFG_F64) != (sizeof(Number) == 8)) throw std::runtime_error("cell_loop: Number differs from the MatrixFreeGpu dtype");
  if (g.constraint_mask != nullptr && g.use_coloring) throw std::runtime_error("cell_loop: hanging nodes need the atomic scatter");
  constexpr unsigned int n = fe_degree + 1, npc = dim == 2 ? n * n : n * n * n;
  GpuData<dim, Number> gd;
  gd.loc2glob = g.loc2glob;
  gd.JxW = static_cast<const Number *>(g.JxW);
  gd.inv_jac = static_cast<const Number *>(g.inv_jac);
  gd.quadrature_points = static_cast<const Number *>(g.quadrature_points);
  gd.shape_values = gd.shape_gradients = nullptr;
  gd.general = g.general;
  gd.use_coloring = g.use_coloring;
  gd.constraint_mask = nullptr;
  gd.hanging_weights = nullptr;
  ShapeTables<Number, n> tab;
  fill_shape_tables<Number, n>(g, tab);
  const unsigned int cpb = npc >= 128 ? 1 : 128 / npc;
  const size_t       smem = (size_t)cpb * (3 + dim) * npc * sizeof(Number);
  auto kern = apply_kernel_shmem_dst<LocOp, dim, fe_degree, Number>;
  if (smem > 48 * 1024) fee_check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute");
  cudaStream_t st = static_cast<cudaStream_t>(g.cuda_stream);
  for (uint32_t c = 0; c < g.n_colors; ++c)
    {
      const uint32_t c0 = g.color_offsets[c], c1 = g.color_offsets[c + 1];
      if (c1 <= c0) continue;
      gd.n_cells = c1;
      kern<<<(c1 - c0 + cpb - 1) / cpb, cpb * npc, smem, st>>>(dst_dev, loc_op, gd, tab, c0);
      fee_check_cuda(cudaGetLastError(), "cell_loop launch");
    }
  if (g.constraint_mask != nullptr && g.n_plain_cells < g.n_cells)
    {
      gd.constraint_mask = g.constraint_mask;
      gd.n_cells = g.n_cells;
      kern<<<(g.n_cells - g.n_plain_cells + cpb - 1) / cpb, cpb * npc, smem, st>>>(dst_dev, loc_op, gd, tab, g.n_plain_cells);
      fee_check_cuda(cudaGetLastError(), "cell_loop launch (hanging-node cells)");
    }
}

// MatrixFreeGpu::evaluate_on_cells<Op>(vec) (matrix_free_gpu.h:415-435): vec[cell][q] = what Op::eval writes for the cell's
// quadrature points; vec_dev holds n_cells * n_q_points entries in kernel cell order (the order of get_gpu_data); the mf
// must carry quadrature points (uniform meshes always do, mfg_mf_desc.quadrature_points otherwise)
template <int dim, int fe_degree, typename Number, typename Op>
void evaluate_on_cells(mfg_mf *mf, Number *vec_dev)
{
  mfg_gpu_data g;
  fee_check(mfg_mf_get_gpu_data(mf, &g), "mfg_mf_get_gpu_data");
  if (g.dim != dim || g.degree != fe_degree) throw std::runtime_error("evaluate_on_cells: dim / fe_degree differ from the MatrixFreeGpu object");
  if ((g.dtype == MFG_F64) != (sizeof(Number) == 8)) throw std::runtime_error("evaluate_on_cells: Number differs from the MatrixFreeGpu dtype");
  if (g.quadrature_points == nullptr) throw std::runtime_error("evaluate_on_cells: the MatrixFreeGpu object has no quadrature points");
  constexpr unsigned int n = fe_degree + 1, nq = dim == 2 ? n * n : n * n * n;
  GpuData<dim, Number> gd;
  gd.loc2glob = g.loc2glob;
  gd.JxW = static_cast<const Number *>(g.JxW);
  gd.inv_jac = static_cast<const Number *>(g.inv_jac);
  gd.quadrature_points = static_cast<const Number *>(g.quadrature_points);
  gd.shape_values = gd.shape_gradients = nullptr;
  gd.general = g.general;
  gd.use_coloring = g.use_coloring;
  gd.constraint_mask = nullptr;
  gd.hanging_weights = nullptr;
  gd.n_cells = g.n_cells;
  if (g.n_cells == 0) return;
  cell_eval_kernel<dim, Number, Op><<<(g.n_cells + 127) / 128, 128, 0, static_cast<cudaStream_t>(g.cuda_stream)>>>(vec_dev, gd, nq);
  fee_check_cuda(cudaGetLastError(), "evaluate_on_cells launch");
}

// the same on the facade classes: data.cell_loop(dst, src, loc_op) of the reference (matrix_free_gpu.h:212-219)
template <int dim, int fe_degree, typename Number, typename LocOp>
void cell_loop(const MatrixFreeGpu<dim, Number> &data, GpuVector<Number> &dst, const GpuVector<Number> &src, const LocOp &loc_op)
{
  cell_loop<dim, fe_degree, Number, LocOp>(data.handle(), dst.getData(), src.getDataRO(), loc_op);
}
template <int dim, int fe_degree, typename Number, typename LocOp>
void cell_loop(const MatrixFreeGpu<dim, Number> &data, GpuVector<Number> &dst, const LocOp &loc_op)
{
  cell_loop<dim, fe_degree, Number, LocOp>(data.handle(), dst.getData(), loc_op);
}
// data.template evaluate_on_cells<Op>(vec) (matrix_free_gpu.h:223-224); vec is resized like the reference does
template <int dim, int fe_degree, typename Number, typename Op>
void evaluate_on_cells(const MatrixFreeGpu<dim, Number> &data, GpuVector<Number> &vec)
{
  constexpr unsigned int n = fe_degree + 1, nq = dim == 2 ? n * n : n * n * n;
  vec.resize((size_t)data.n_cells_tot * nq);
  evaluate_on_cells<dim, fe_degree, Number, Op>(data.handle(), vec.getData());
}

}  // namespace dealii_cuda_b200
