"""dealii_cuda_b200 -- B200-native matrix-free finite-element operator engine.

Python mirror of the reference's operator interface (kalj/dealii-cuda), a thin
layer over the C ABI of libmfgpu.so (include/mfgpu.h).  Class and method names
follow the reference:

  GpuVector                 matrix_free_gpu/gpu_vec.h:22-176
  ConstraintHandlerGpu      matrix_free_gpu/constraint_handler_gpu.h:13-59
  MatrixFreeGpu             matrix_free_gpu/matrix_free_gpu.h:81-229
  LaplaceOperatorGpu        laplace_operator_gpu.h:35-96
  HyperCubeMesh             stands in for Triangulation + DoFHandler +
                            ConstraintMatrix (bmop.cu:111-132)

All numerical work happens in hand-written CUDA kernels inside libmfgpu.so;
this module never computes on the CPU and never imports the oracle.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import F32, F64, SCATTER_ATOMIC, SCATTER_COLOR, MfgError, check, lib

__all__ = ["Context", "GpuVector", "HyperCubeMesh", "AdaptiveMesh", "BallMesh", "MatrixFreeGpu", "ConstraintHandlerGpu", "LaplaceOperatorGpu",
           "shape_info", "solver_cg", "hanging_node_weights", "F32", "F64", "SCATTER_ATOMIC", "SCATTER_COLOR", "MfgError"]

_NP = {F32: np.float32, F64: np.float64}


def _dtype_code(dtype):
    if dtype in (F32, F64):
        return dtype
    dt = np.dtype(dtype)
    if dt == np.float64:
        return F64
    if dt == np.float32:
        return F32
    raise ValueError("dtype must be float32 or float64")


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """Device + stream the library enqueues on (the reference uses device 0 / default stream)."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        check(lib.mfg_ctx_create(int(device), C.c_void_p(stream or 0), C.byref(h)))
        self.h = h
        self.device = device

    def set_stream(self, stream):
        check(lib.mfg_ctx_set_stream(self.h, C.c_void_p(stream or 0)))

    def synchronize(self):
        check(lib.mfg_ctx_synchronize(self.h))

    def device_info(self):
        sm, ma, mi, l2 = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(lib.mfg_ctx_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(l2)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "l2_bytes": l2.value}

    def __del__(self):
        try:
            if self.h:
                lib.mfg_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass


class GpuVector:
    """GpuVector<Number> (gpu_vec.h:22-176)."""

    def __init__(self, ctx, n=0, dtype=np.float64, _handle=None, _borrowed=False, _keepalive=None):
        self.ctx = ctx
        self._borrowed = _borrowed
        self._keepalive = _keepalive
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            check(lib.mfg_vec_create(ctx.h, _dtype_code(dtype), int(n), C.byref(h)))
            self.h = h

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a)
        v = cls(ctx, a.size, a.dtype)
        v.fromHost(a)
        return v

    @classmethod
    def wrap(cls, ctx, tensor):
        """Non-owning view of a contiguous 1-D torch CUDA tensor (float32/float64)."""
        import torch
        assert tensor.is_cuda and tensor.is_contiguous() and tensor.dim() == 1
        code = F64 if tensor.dtype == torch.float64 else F32
        h = C.c_void_p()
        check(lib.mfg_vec_wrap(ctx.h, code, tensor.numel(), C.c_void_p(tensor.data_ptr()), C.byref(h)))
        return cls(ctx, _handle=h, _keepalive=tensor)

    def __del__(self):
        try:
            if self.h and not self._borrowed:
                lib.mfg_vec_destroy(self.h)
            self.h = None
        except Exception:
            pass

    def size(self):
        return lib.mfg_vec_size(self.h)

    @property
    def dtype(self):
        return _NP[lib.mfg_vec_dtype(self.h)]

    def getData(self):
        return lib.mfg_vec_data(self.h)

    getDataRO = getData

    def resize(self, n):
        check(lib.mfg_vec_resize(self.h, int(n)))

    def reinit(self, n_or_vec, leave_elements_uninitialized=False):
        n = n_or_vec.size() if isinstance(n_or_vec, GpuVector) else int(n_or_vec)
        self.resize(n)
        if not leave_elements_uninitialized:
            self.fill(0.0)

    def fromHost(self, a):
        a = np.ascontiguousarray(a, dtype=self.dtype)
        check(lib.mfg_vec_from_host(self.h, a.ctypes.data_as(C.c_void_p), a.size))

    def toVector(self):
        out = np.empty(self.size(), dtype=self.dtype)
        check(lib.mfg_vec_to_host(self.h, out.ctypes.data_as(C.c_void_p), out.size))
        return out

    copyToHost = toVector

    def assign(self, other):
        """operator=(GpuVector) / operator=(Number)."""
        if isinstance(other, GpuVector):
            check(lib.mfg_vec_copy(self.h, other.h))
        else:
            self.fill(float(other))
        return self

    def fill(self, a):
        check(lib.mfg_vec_fill(self.h, float(a)))

    def swap(self, other):
        check(lib.mfg_vec_swap(self.h, other.h))
        self._keepalive, other._keepalive = other._keepalive, self._keepalive

    def sadd(self, s, a_or_v, v=None):
        if v is None:  # sadd(s, V): this = s*this + V
            check(lib.mfg_vec_sadd(self.h, float(s), 1.0, a_or_v.h))
        else:
            check(lib.mfg_vec_sadd(self.h, float(s), float(a_or_v), v.h))

    def add(self, a_or_v, v=None):
        if v is None:
            check(lib.mfg_vec_sadd(self.h, 1.0, 1.0, a_or_v.h))
        else:
            check(lib.mfg_vec_sadd(self.h, 1.0, float(a_or_v), v.h))

    def equ(self, a, x):
        check(lib.mfg_vec_equ(self.h, float(a), x.h))

    def scale(self, x):
        check(lib.mfg_vec_scale(self.h, x.h))

    def __itruediv__(self, x):
        check(lib.mfg_vec_divide(self.h, x.h))
        return self

    def __imul__(self, a):
        check(lib.mfg_vec_scal(self.h, float(a)))
        return self

    def __iadd__(self, x):
        self.add(x)
        return self

    def __isub__(self, x):
        self.add(-1.0, x)
        return self

    def invert(self):
        check(lib.mfg_vec_invert(self.h))
        return self

    def dot(self, other):
        out = C.c_double()
        check(lib.mfg_vec_dot(self.h, other.h, C.byref(out)))
        return out.value

    __mul__ = dot

    def add_and_dot(self, a, x, v):
        out = C.c_double()
        check(lib.mfg_vec_add_and_dot(self.h, float(a), x.h, v.h, C.byref(out)))
        return out.value

    def l2_norm(self):
        out = C.c_double()
        check(lib.mfg_vec_l2_norm(self.h, C.byref(out)))
        return out.value

    def all_zero(self):
        out = C.c_int()
        check(lib.mfg_vec_all_zero(self.h, C.byref(out)))
        return bool(out.value)

    def memory_consumption(self):
        return self.size() * np.dtype(self.dtype).itemsize


class HyperCubeMesh:
    """hyper_cube(left,right)^dim + refine_global(r) + FE_Q(degree) + Dirichlet boundary,
    or a general box of 2^k cells per direction (partitions of the cube)."""

    def __init__(self, ctx, dim, degree, n_refine=None, left=-1.0, right=1.0, box=None):
        self.ctx = ctx
        h = C.c_void_p()
        if box is None:
            check(lib.mfg_mesh_hyper_cube(ctx.h, dim, degree, int(n_refine), float(left), float(right), C.byref(h)))
        else:
            d = _capi.BoxDesc()
            d.dim, d.degree = dim, degree
            for k in range(3):
                d.log2_cells[k] = int(box["log2_cells"][k]) if k < dim else 0
                d.origin[k] = float(box["origin"][k]) if k < dim else 0.0
            d.h = float(box["h"])
            d.dirichlet_faces = int(box.get("dirichlet_faces", 0x3f))
            check(lib.mfg_mesh_create_box(ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self.dim, self.degree = dim, degree
        self.n_cells = lib.mfg_mesh_n_cells(h)
        self.n_dofs = lib.mfg_mesh_n_dofs(h)
        self.dofs_per_cell = lib.mfg_mesh_dofs_per_cell(h)
        self.n_constrained = lib.mfg_mesh_n_constrained(h)

    def __del__(self):
        try:
            if self.h:
                lib.mfg_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def loc2glob(self):
        out = np.empty((self.n_cells, self.dofs_per_cell), dtype=np.uint32)
        check(lib.mfg_mesh_get_loc2glob(self.h, _u32p(out)))
        return out

    def constrained_dofs(self):
        out = np.empty(self.n_constrained, dtype=np.uint32)
        check(lib.mfg_mesh_get_constrained(self.h, _u32p(out)))
        return out

    def cell_coords(self):
        out = np.empty((self.n_cells, 3), dtype=np.uint32)
        check(lib.mfg_mesh_get_cell_coords(self.h, _u32p(out)))
        return out

    def support_points(self):
        """[n_dofs][dim] support point of every DoF (DoFTools::map_dofs_to_support_points)."""
        out = np.empty((self.n_dofs, self.dim), dtype=np.float64)
        check(lib.mfg_mesh_get_support_points(self.h, _dp(out)))
        return out

    def lattice_to_dof(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.uint32).reshape(-1, 3)
        out = np.empty(xyz.shape[0], dtype=np.uint32)
        check(lib.mfg_mesh_lattice_to_dof(self.h, xyz.shape[0], _u32p(xyz), _u32p(out)))
        return out

    def color_cells(self):
        out = np.empty(self.n_cells, dtype=np.uint32)
        nc = C.c_uint32()
        check(lib.mfg_mesh_color_cells(self.h, _u32p(out), C.byref(nc)))
        return out, nc.value


def graph_coloring(loc2glob, n_indices):
    """GraphColoringWrapper::make_graph_coloring (coloring.cc:8-33) restated on the host: (color_of_cell, n_colors)"""
    l2g = np.ascontiguousarray(loc2glob, dtype=np.uint32)
    col = np.zeros(l2g.shape[0], dtype=np.uint32)
    nc = C.c_uint32()
    check(lib.mfg_graph_coloring(l2g.shape[0], l2g.shape[1], _u32p(l2g), int(n_indices), _u32p(col), C.byref(nc)))
    return col, nc.value


def assemble_laplace_csr(dim, degree, loc2glob, n_dofs, inv_jac, coefficient, constrained):
    """host assembly of the operator as CSR (bmop_spm.cu:150-201 restated in the library, mfg_csr_assemble_laplace):
    returns (row_ptr, col, val) numpy arrays"""
    l2g = np.ascontiguousarray(loc2glob, dtype=np.uint32)
    ij = np.ascontiguousarray(inv_jac, dtype=np.float64)
    cf = np.ascontiguousarray(coefficient, dtype=np.float64)
    con = np.ascontiguousarray(constrained, dtype=np.uint32)
    h = C.c_void_p()
    check(lib.mfg_csr_assemble_laplace(int(dim), int(degree), l2g.shape[0], int(n_dofs), _u32p(l2g), _dp(ij), _dp(cf), _u32p(con), con.size, C.byref(h)))
    try:
        n, nnz = C.c_uint32(), C.c_size_t()
        check(lib.mfg_csr_sizes(h, C.byref(n), C.byref(nnz)))
        rp, col, val = np.zeros(n.value + 1, np.uint32), np.zeros(nnz.value, np.uint32), np.zeros(nnz.value)
        check(lib.mfg_csr_get(h, _u32p(rp), _u32p(col), _dp(val)))
    finally:
        lib.mfg_csr_destroy(h)
    return rp, col, val


class SparseMatrixGpu:
    """CUDAWrappers::SparseMatrix<Number> (matrix_free_gpu/cuda_sparse_matrix.h): the assembled competitor of the matrix-free
    operator (bmop_spm.cu), here for the Laplace operator of a uniform mesh: reinit(mesh), vmult, m, n_nonzero_elements"""

    def __init__(self, ctx, dtype=np.float64):
        self.ctx, self.code, self.h = ctx, _dtype_code(dtype), None

    def reinit(self, mesh):
        self.clear()
        h = C.c_void_p()
        check(lib.mfg_spm_create_from_mesh(self.ctx.h, mesh.h, self.code, C.byref(h)))
        self.h = h

    def clear(self):
        if self.h:
            lib.mfg_spm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

    def m(self):
        return lib.mfg_spm_m(self.h)

    n = m

    def n_nonzero_elements(self):
        return lib.mfg_spm_n_nonzero_elements(self.h)

    def memory_consumption(self):
        return lib.mfg_spm_memory_consumption(self.h)

    def vmult(self, dst, src):
        check(lib.mfg_spm_vmult(self.h, dst.h, src.h))


def hanging_node_weights(degree):
    """W[k][i] = phi_i(xi_k/2) (setup_constraint_weights, hanging_nodes.cuh:580-598)."""
    n = degree + 1
    w = np.empty((n, n))
    check(lib.mfg_hanging_node_weights(degree, _dp(w)))
    return w


def shape_info(degree):
    """ShapeInfo::shape_values / shape_gradients [i*n+q], Gauss points and weights on [0,1]."""
    n = degree + 1
    val, grad, xq, wq = np.empty((n, n)), np.empty((n, n)), np.empty(n), np.empty(n)
    check(lib.mfg_shape_info(degree, _dp(val), _dp(grad), _dp(xq), _dp(wq)))
    return val, grad, xq, wq


def solver_cg(op, x, b, abs_tol, max_iter=10000, use_jacobi=True, history=False):
    """SolverCG<GpuVector>::solve(op, x, b, preconditioner) as the reference runs it (poisson.cu:247-260).
    Returns (iterations, last residual[, residual history])."""
    it, res = C.c_int(), C.c_double()
    hist = np.zeros(max_iter + 1) if history else None
    check(lib.mfg_solver_cg(op.h, x.h, b.h, float(abs_tol), int(max_iter), 1 if use_jacobi else 0, C.byref(it), C.byref(res),
                            _dp(hist) if history else None))
    if history:
        return it.value, res.value, hist[:it.value + 1]
    return it.value, res.value


class ConstraintHandlerGpu:
    """ConstraintHandlerGpu<Number> (constraint_handler_gpu.h:13-59)."""

    def __init__(self, ctx, dtype=np.float64):
        self.ctx, self.code, self.h = ctx, _dtype_code(dtype), None

    def reinit(self, constraints, n_dofs=None, edge_indices=None):
        """constraints: HyperCubeMesh (its boundary ConstraintMatrix) or an ascending index array."""
        self._free()
        h = C.c_void_p()
        if isinstance(constraints, HyperCubeMesh):
            check(lib.mfg_ch_create_from_mesh(self.ctx.h, self.code, constraints.h, C.byref(h)))
        else:
            c = np.ascontiguousarray(constraints, dtype=np.uint32)
            e = np.ascontiguousarray(edge_indices if edge_indices is not None else [], dtype=np.uint32)
            check(lib.mfg_ch_create(self.ctx.h, self.code, _u32p(c), c.size, _u32p(e), e.size, C.byref(h)))
        self.h = h

    def _free(self):
        if self.h:
            lib.mfg_ch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self._free()
        except Exception:
            pass

    def n_constrained(self):
        return lib.mfg_ch_n_constrained(self.h)

    def set_constrained_values(self, v, val):
        check(lib.mfg_ch_set_constrained_values(self.h, v.h, float(val)))

    def save_constrained_values(self, v1, v2=None):
        if v2 is None:
            check(lib.mfg_ch_save_constrained_values(self.h, v1.h))
        else:
            check(lib.mfg_ch_save_constrained_values2(self.h, v1.h, v2.h))

    def load_constrained_values(self, v):
        check(lib.mfg_ch_load_constrained_values(self.h, v.h))

    def load_and_add_constrained_values(self, v1, v2):
        check(lib.mfg_ch_load_and_add_constrained_values(self.h, v1.h, v2.h))

    def copy_edge_values(self, dst, src):
        check(lib.mfg_ch_copy_edge_values(self.h, dst.h, src.h))


class AdaptiveMesh:
    """Triangulation + DoFHandler + HangingNodes on hyper_cube(left, right) with local refinement (mfg_amesh_*, host code of
    the library): what the reference takes from deal.II for its adaptive-grid runs (bmop_common.h:9-120,
    matrix_free_gpu/hanging_nodes.cuh:209-454).  Needs no device; LaplaceOperatorGpu.reinit(adaptive_mesh) builds the
    device objects."""

    def __init__(self, dim, degree, left=-1.0, right=1.0, limit_level_difference_at_vertices=False):
        h = C.c_void_p()
        check(lib.mfg_amesh_create(int(dim), int(degree), float(left), float(right), C.byref(h)))
        self.h, self.dim, self.degree, self.left, self.right = h, dim, degree, left, right
        if limit_level_difference_at_vertices:   # Triangulation::MeshSmoothing of the reference's MG drivers (poisson_mg.cu:132)
            check(lib.mfg_amesh_set_limit_level_difference_at_vertices(h, 1))
        self.dofs_per_cell = (degree + 1) ** dim

    def __del__(self):
        try:
            if self.h:
                lib.mfg_amesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def refine_global(self, times=1):
        check(lib.mfg_amesh_refine_global(self.h, int(times)))
        return self

    def set_refine_flags(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.uint8)
        check(lib.mfg_amesh_set_refine_flags(self.h, f.ctypes.data_as(C.POINTER(C.c_uint8)), f.size))

    def _center(self, center):
        if center is None:
            return None
        c = np.zeros(3)
        c[:self.dim] = np.asarray(center, dtype=np.float64)[:self.dim]
        return c

    def mark_cells_in_annulus(self, R, r=0.0, center=None):
        c = self._center(center)
        check(lib.mfg_amesh_mark_cells_in_annulus(self.h, float(R), float(r), _dp(c) if c is not None else None))

    def mark_cells_on_shell(self, R, center=None):
        c = self._center(center)
        check(lib.mfg_amesh_mark_cells_on_shell(self.h, float(R), _dp(c) if c is not None else None))

    def mark_octant(self):
        check(lib.mfg_amesh_mark_octant(self.h))

    def execute_coarsening_and_refinement(self):
        check(lib.mfg_amesh_execute_refinement(self.h))
        return self

    def pseudo_adaptive_refinement(self, n_ref):
        """bmop_common.h:49-105 (domain CUBE) on the unrefined hyper_cube"""
        check(lib.mfg_amesh_pseudo_adaptive_refinement(self.h, int(n_ref)))
        return self

    @property
    def n_cells(self):
        return lib.mfg_amesh_n_active_cells(self.h)

    @property
    def n_levels(self):
        return lib.mfg_amesh_n_levels(self.h)

    def active_cells(self):
        """[n_cells][1 + dim]: level and integer coordinates of the active cells in iteration order"""
        out = np.zeros((self.n_cells, 4), dtype=np.uint32)
        check(lib.mfg_amesh_get_active_cells(self.h, _u32p(out)))
        return out[:, :1 + self.dim]

    def level_cells(self, level):
        """[n][dim + 1]: integer coordinates of ALL cells of a level in storage order and 1 where the cell has children"""
        out = np.zeros((lib.mfg_amesh_n_level_cells(self.h, int(level)), 4), dtype=np.uint32)
        if out.shape[0]:
            check(lib.mfg_amesh_get_level_cells(self.h, int(level), _u32p(out)))
        return np.concatenate([out[:, :self.dim], out[:, 3:]], axis=1)

    def distribute_dofs(self):
        check(lib.mfg_amesh_distribute_dofs(self.h))
        self.n_dofs = lib.mfg_amesh_n_dofs(self.h)
        return self

    def boundary_dofs(self):
        out = np.zeros(lib.mfg_amesh_n_boundary(self.h), np.uint32)
        check(lib.mfg_amesh_get_boundary(self.h, _u32p(out)))
        return out

    def support_points(self):
        """DoFTools::map_dofs_to_support_points: [n_dofs][dim]"""
        out = np.zeros((self.n_dofs, self.dim))
        check(lib.mfg_amesh_get_support_points(self.h, _dp(out)))
        return out

    def build_mg(self, min_level=0):
        """level meshes, level DoFs, MGConstrainedDoFs sets, transfer blocks and copy indices of the multigrid hierarchy"""
        check(lib.mfg_amesh_build_mg(self.h, int(min_level)))
        self.mg_min_level = int(min_level)
        return self

    def mg_level(self, level):
        """dict of the host arrays of one level of the hierarchy (mfg_amesh_mg_level_get)"""
        sz = np.zeros(6, np.uint32)
        check(lib.mfg_amesh_mg_level_sizes(self.h, int(level), _u32p(sz)))
        nc, nd, nb, ne, npar, ncp = (int(v) for v in sz)
        npc, nF, n3 = self.dofs_per_cell, (2 * self.degree + 1) ** self.dim, 3 ** self.dim
        a = dict(n_dofs=nd, loc2glob=np.zeros((nc, npc), np.uint32), boundary=np.zeros(nb, np.uint32), edge=np.zeros(ne, np.uint32),
                 coefficient=np.zeros((nc, npc)), copy_global=np.zeros(ncp, np.uint32), copy_level=np.zeros(ncp, np.uint32),
                 coarse_idx=np.zeros((npar, npc), np.uint32), fine_idx=np.zeros((npar, nF), np.uint32), weights=np.zeros((npar, n3)))
        check(lib.mfg_amesh_mg_level_get(self.h, int(level), _u32p(a["loc2glob"]), _u32p(a["boundary"]), _u32p(a["edge"]), _dp(a["coefficient"]),
                                         _u32p(a["copy_global"]), _u32p(a["copy_level"]), _u32p(a["coarse_idx"]), _u32p(a["fine_idx"]), _dp(a["weights"])))
        return a

    def arrays(self, quadrature_points=False):
        """dict: loc2glob (rewritten), loc2glob_unconstrained, constraint_mask, constrained, hanging, inv_jac, coefficient"""
        nc, npc = self.n_cells, self.dofs_per_cell
        a = dict(loc2glob=np.zeros((nc, npc), np.uint32), loc2glob_unconstrained=np.zeros((nc, npc), np.uint32),
                 constraint_mask=np.zeros(nc, np.uint32), constrained=np.zeros(lib.mfg_amesh_n_constrained(self.h), np.uint32),
                 hanging=np.zeros(lib.mfg_amesh_n_hanging(self.h), np.uint32), inv_jac=np.zeros(nc), coefficient=np.zeros((nc, npc)))
        qp = np.zeros((nc, npc, self.dim)) if quadrature_points else None
        check(lib.mfg_amesh_get_arrays(self.h, _u32p(a["loc2glob"]), _u32p(a["loc2glob_unconstrained"]), _u32p(a["constraint_mask"]),
                                       _u32p(a["constrained"]), _u32p(a["hanging"]), _dp(a["inv_jac"]), _dp(a["coefficient"]),
                                       _dp(qp) if qp is not None else None))
        if qp is not None:
            a["quadrature_points"] = qp
        return a


class BallMesh:
    """GridGenerator::hyper_ball + SphericalManifold on the boundary + refine_global (the reference's BALL_GRID,
    poisson_common.h:59-72) with FE_Q DoFs and MappingQ1 geometry: host substrate of the library (mfg_umesh_*).  Needs no device;
    LaplaceOperatorGpu.reinit(ball_mesh) builds the operator on the general-geometry path."""

    def __init__(self, dim, degree, n_refine=0, radius=1.0):
        h = C.c_void_p()
        check(lib.mfg_umesh_hyper_ball(int(dim), int(degree), float(radius), C.byref(h)))
        self.h, self.dim, self.degree, self.radius = h, dim, degree, radius
        self.dofs_per_cell = (degree + 1) ** dim
        if n_refine:
            self.refine_global(n_refine)

    def __del__(self):
        try:
            if self.h:
                lib.mfg_umesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def refine_global(self, times=1):
        check(lib.mfg_umesh_refine_global(self.h, int(times)))
        return self

    def distribute_dofs(self):
        check(lib.mfg_umesh_distribute_dofs(self.h))
        self.n_dofs = lib.mfg_umesh_n_dofs(self.h)
        return self

    @property
    def n_cells(self):
        return lib.mfg_umesh_n_cells(self.h)

    def mesh(self):
        """(vertices [n_vertices][dim], cell_vertices [n_cells][2^dim] in local lexicographic order)"""
        v = np.zeros((lib.mfg_umesh_n_vertices(self.h), self.dim))
        c = np.zeros((self.n_cells, 1 << self.dim), np.uint32)
        check(lib.mfg_umesh_get_mesh(self.h, _dp(v), _u32p(c)))
        return v, c

    def support_points(self):
        """DoFTools::map_dofs_to_support_points (MappingQ1): [n_dofs][dim]"""
        out = np.zeros((self.n_dofs, self.dim))
        check(lib.mfg_umesh_get_support_points(self.h, _dp(out)))
        return out

    def arrays(self):
        """dict: loc2glob, boundary (Dirichlet DoFs), inv_jac [cell][q][d1][d2], JxW, quadrature_points, coefficient"""
        nc, npc, dim = self.n_cells, self.dofs_per_cell, self.dim
        a = dict(loc2glob=np.zeros((nc, npc), np.uint32), boundary=np.zeros(lib.mfg_umesh_n_boundary(self.h), np.uint32),
                 inv_jac=np.zeros((nc, npc, dim, dim)), JxW=np.zeros((nc, npc)), quadrature_points=np.zeros((nc, npc, dim)),
                 coefficient=np.zeros((nc, npc)))
        check(lib.mfg_umesh_get_arrays(self.h, _u32p(a["loc2glob"]), _u32p(a["boundary"]), _dp(a["inv_jac"]), _dp(a["JxW"]),
                                       _dp(a["quadrature_points"]), _dp(a["coefficient"])))
        return a


class MatrixFreeGpu:
    """MatrixFreeGpu<dim,Number> (matrix_free_gpu.h:81-229): reinit / counters / free."""

    def __init__(self, ctx, dtype=np.float64):
        self.ctx, self.code, self.h = ctx, _dtype_code(dtype), None
        self._keep = None

    def reinit(self, mesh_or_arrays, use_coloring=False):
        self.free()
        h = C.c_void_p()
        scatter = SCATTER_COLOR if use_coloring else SCATTER_ATOMIC
        if isinstance(mesh_or_arrays, HyperCubeMesh):
            check(lib.mfg_mf_reinit_from_mesh(self.ctx.h, mesh_or_arrays.h, self.code, scatter, C.byref(h)))
            self._keep = mesh_or_arrays
        elif isinstance(mesh_or_arrays, AdaptiveMesh):
            assert not use_coloring, "hanging nodes need the atomic scatter"
            check(lib.mfg_mf_reinit_from_amesh(self.ctx.h, mesh_or_arrays.h, self.code, C.byref(h)))
            self._keep = mesh_or_arrays
        elif isinstance(mesh_or_arrays, BallMesh):
            assert not use_coloring
            check(lib.mfg_mf_reinit_from_umesh(self.ctx.h, mesh_or_arrays.h, self.code, C.byref(h)))
            self._keep = mesh_or_arrays
        else:
            a = mesh_or_arrays
            d = _capi.MfDesc()
            l2g = np.ascontiguousarray(a["loc2glob"], dtype=np.uint32)
            invj = np.ascontiguousarray(a["inv_jac"], dtype=np.float64)
            d.dim, d.degree, d.dtype = int(a["dim"]), int(a["degree"]), self.code
            d.n_cells, d.n_dofs = l2g.shape[0], int(a["n_dofs"])
            # inv_jac [n_cells]: uniform mesh; [n_cells][npc][dim][dim]: full inverse Jacobian per quadrature point
            d.loc2glob, d.geometry, d.inv_jac = _u32p(l2g), (1 if invj.ndim == 4 else 0), _dp(invj)
            keep = [l2g, invj]
            if a.get("JxW") is not None:
                jxw = np.ascontiguousarray(a["JxW"], dtype=np.float64)
                d.JxW = _dp(jxw)
                keep.append(jxw)
            d.scatter = scatter
            if use_coloring:
                off = np.ascontiguousarray(a["color_offsets"], dtype=np.uint32)
                d.n_colors, d.color_offsets = off.size - 1, _u32p(off)
                keep.append(off)
            if a.get("constraint_mask") is not None:
                cm = np.ascontiguousarray(a["constraint_mask"], dtype=np.uint32)
                d.constraint_mask = _u32p(cm)
                keep.append(cm)
            check(lib.mfg_mf_reinit(self.ctx.h, C.byref(d), C.byref(h)))
        self.h = h
        self.n_dofs = lib.mfg_mf_n_dofs(h)
        self.n_cells_tot = lib.mfg_mf_n_cells(h)
        self.num_colors = lib.mfg_mf_n_colors(h)
        self.use_coloring = bool(use_coloring)

    def get_gpu_data(self):
        """MatrixFreeGpu::get_gpu_data (matrix_free_gpu.h:261-278): raw device pointers and counters for user kernels
        (torch / cupy / the header-only FEEvaluationGpu path); valid until free()."""
        g = _capi.GpuData()
        check(lib.mfg_mf_get_gpu_data(self.h, C.byref(g)))
        return g

    def free(self):
        if self.h:
            lib.mfg_mf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def memory_consumption(self):
        return lib.mfg_mf_memory_consumption(self.h)


class LaplaceOperatorGpu:
    """LaplaceOperatorGpu<dim,fe_degree,Number> (laplace_operator_gpu.h:35-96)."""

    def __init__(self, ctx, dtype=np.float64, use_coloring=False):
        self.ctx, self.code, self.h = ctx, _dtype_code(dtype), None
        self.use_coloring = use_coloring
        self._keep = None

    def reinit(self, mesh, constraints=None, coefficient=None):
        """reinit(dof_handler, constraints): `mesh` is a HyperCubeMesh (carrying its boundary
        constraints), or a MatrixFreeGpu built from explicit arrays together with a
        ConstraintHandlerGpu and the coefficient values at the quadrature points."""
        self.clear()
        h = C.c_void_p()
        if isinstance(mesh, HyperCubeMesh):
            scatter = SCATTER_COLOR if self.use_coloring else SCATTER_ATOMIC
            check(lib.mfg_laplace_create(self.ctx.h, mesh.h, self.code, scatter, C.byref(h)))
            self._keep = mesh
        elif isinstance(mesh, AdaptiveMesh):
            assert not self.use_coloring, "hanging nodes need the atomic scatter"
            check(lib.mfg_laplace_create_from_amesh(self.ctx.h, mesh.h, self.code, C.byref(h)))
            self._keep = mesh
        elif isinstance(mesh, BallMesh):
            assert not self.use_coloring
            check(lib.mfg_laplace_create_from_umesh(self.ctx.h, mesh.h, self.code, C.byref(h)))
            self._keep = mesh
        else:
            coef = np.ascontiguousarray(coefficient, dtype=np.float64)
            check(lib.mfg_laplace_create_from_arrays(self.ctx.h, mesh.h, constraints.h, _dp(coef), C.byref(h)))
            self._keep = (mesh, constraints)
        self.h = h

    def clear(self):
        if self.h:
            lib.mfg_laplace_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

    def m(self):
        return lib.mfg_laplace_m(self.h)

    n = m

    def set_coefficient(self, coef):
        coef = np.ascontiguousarray(coef, dtype=np.float64)
        check(lib.mfg_laplace_set_coefficient(self.h, _dp(coef)))

    def set_variant(self, variant):
        check(lib.mfg_laplace_set_variant(self.h, int(variant)))

    def vmult(self, dst, src):
        check(lib.mfg_laplace_vmult(self.h, dst.h, src.h))

    Tvmult = vmult

    def vmult_add(self, dst, src):
        check(lib.mfg_laplace_vmult_add(self.h, dst.h, src.h))

    Tvmult_add = vmult_add

    def vmult_ptr(self, dst_ptr, src_ptr):
        check(lib.mfg_laplace_vmult_ptr(self.h, C.c_void_p(dst_ptr), C.c_void_p(src_ptr)))

    def set_interface_dofs(self, dofs):
        """Multi-GPU: DoFs whose partial sums are exchanged after the cell loop; returns the number of cell groups that
        contribute to them (0: the active kernel has no work list, part 0 then does the whole apply)."""
        d = np.ascontiguousarray(dofs, dtype=np.uint32)
        k = C.c_uint32()
        check(lib.mfg_laplace_set_interface_dofs(self.h, d.ctypes.data_as(C.POINTER(C.c_uint32)), d.size, C.byref(k)))
        return k.value

    def vmult_part_ptr(self, dst_ptr, src_ptr, part, stream=None):
        """part 0: zero/constraint pass, 1: interface cell groups, 2: the other groups, -1: everything; stream: raw
        CUDA stream to enqueue on (None: the context's)."""
        check(lib.mfg_laplace_vmult_part_ptr(self.h, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(part), C.c_void_p(stream or 0)))

    def vmult_host(self, dst, src):
        """dst, src: contiguous numpy arrays (or pinned torch CPU tensors via .numpy())."""
        check(lib.mfg_laplace_vmult_host(self.h, dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p)))

    def vmult_host_async(self, dst, src, slot):
        """pipelined variant: returns immediately; alternate slot 0/1 with separate pinned buffers, then host_sync()"""
        check(lib.mfg_laplace_vmult_host_async(self.h, dst.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), int(slot)))

    def host_sync(self):
        check(lib.mfg_laplace_host_sync(self.h))

    def compute_diagonal(self):
        check(lib.mfg_laplace_compute_diagonal(self.h))

    def get_diagonal_inverse(self):
        h = C.c_void_p()
        check(lib.mfg_laplace_get_diagonal_inverse(self.h, C.byref(h)))
        return GpuVector(self.ctx, _handle=h, _borrowed=True, _keepalive=self)

    def memory_consumption(self):
        return lib.mfg_laplace_memory_consumption(self.h)

    def launches_per_vmult(self):
        return lib.mfg_laplace_launches_per_vmult(self.h)

    def cell_launches_per_vmult(self):
        return lib.mfg_laplace_cell_launches_per_vmult(self.h)

    def active_variant(self):
        v = lib.mfg_laplace_active_variant(self.h)
        if v < 0:   # a requested variant this operator cannot run
            raise MfgError(-4, lib.mfg_last_error().decode(errors="replace"))
        return v

    def set_option(self, name, value):
        check(lib.mfg_laplace_set_option(self.h, name.encode(), int(value)))

    def stage_stats(self):
        """plan of the staged kernel (variant 40) after the first apply; per-group figures are averages over the staged groups"""
        st = (C.c_uint32 * 8)()
        check(lib.mfg_laplace_stage_stats(self.h, st))
        g = max(1, st[1])
        return dict(groups=st[0], staged=st[1], fallback=st[0] - st[1], patterns=st[2], own_per_group=st[3] / 16.0, halo_per_group=st[4] / 16.0,
                    plain_per_group=st[5] / 16.0, red_per_group=st[6] / 16.0, smem_wavefronts_per_group=st[7] / 16.0)

    def enable_kernel_timing(self, on):
        """on = True / k: bracket every (k-th) cell-kernel launch with CUDA events; False: off."""
        check(lib.mfg_laplace_enable_kernel_timing(self.h, int(on)))

    def kernel_time_ms(self):
        ms, nl = C.c_double(), C.c_int()
        check(lib.mfg_laplace_kernel_time_ms(self.h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value

    def bmop(self, dst, src, k, init=0.1):
        """bmop.cu:135-153 loop; returns device milliseconds for the k applications."""
        ms = C.c_float()
        check(lib.mfg_laplace_bmop(self.h, dst.h, src.h, int(k), float(init), C.byref(ms)))
        return ms.value
