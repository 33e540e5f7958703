"""ctypes binding of the CPU oracle (oracle/mf_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs -- never from the product
package (dealii_cuda_b200), which must fail loudly when its CUDA library is
missing instead of falling back to this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.orc_create_box.restype = C.c_void_p
        L.orc_create_box.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_double, C.c_uint32]
        L.orc_cell_loop_range.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_uint32, C.c_uint32]
        L.orc_destroy.argtypes = [C.c_void_p]
        for f in ("orc_n_cells", "orc_n_dofs", "orc_dofs_per_cell", "orc_n_constrained"):
            getattr(L, f).restype = C.c_uint32
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_loc2glob", "orc_constrained", "orc_dof_lattice", "orc_cell_coords", "orc_coefficient",
                  "orc_shape_values", "orc_shape_gradients", "orc_lex2hier"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        L.orc_vmult.argtypes = [C.c_void_p, dp, dp]
        L.orc_vmult_add.argtypes = [C.c_void_p, dp, dp]
        L.orc_vmult_omp.argtypes = [C.c_void_p, dp, dp]
        L.orc_vmult_fast.argtypes = [C.c_void_p, dp, dp]
        L.orc_bmop.argtypes = [C.c_void_p, C.c_int, C.c_double, dp]
        L.orc_inverse_diagonal.argtypes = [C.c_void_p, dp]
        L.orc_assemble_dense.argtypes = [C.c_void_p, dp]
        L.orc_assemble_dense.restype = C.c_int
        L.orc_dense_vmult.argtypes = [C.c_uint32, dp, dp, dp]
        L.orc_fill_sm64.argtypes = [C.c_uint64, C.c_uint32, dp]
        L.orc_set_constant_coefficient.argtypes = [C.c_void_p, C.c_double]
        L.orc_clear_constraints.argtypes = [C.c_void_p]
        L.orc_shape_1d.argtypes = [C.c_int, dp, dp, dp, dp, dp]
        L.orc_hier_to_lex.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint32)]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.restype = None
        L.orc_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    ct = {np.uint32: C.c_uint32, np.float64: C.c_double}[dtype]
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,))
    return arr.reshape(shape)


def sm64(seed, n):
    u = np.empty(n, dtype=np.float64)
    lib().orc_fill_sm64(seed, n, _dp(u))
    return u


def shape_1d(p):
    n = p + 1
    val = np.empty(n * n); grad = np.empty(n * n); xn = np.empty(n); xq = np.empty(n); wq = np.empty(n)
    lib().orc_shape_1d(p, _dp(val), _dp(grad), _dp(xn), _dp(xq), _dp(wq))
    return val.reshape(n, n), grad.reshape(n, n), xn, xq, wq


def hier_to_lex(dim, p):
    out = np.empty((p + 1) ** dim, dtype=np.uint32)
    lib().orc_hier_to_lex(dim, p, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


class OracleMesh:
    """hyper_cube(left,right)^dim + refine_global(r) + FE_Q(p) + Dirichlet boundary."""

    def __init__(self, dim, p, r=None, left=-1.0, right=1.0, box=None):
        self.L = lib()
        if box is None:
            self.h = self.L.orc_create(dim, p, r, left, right)
        else:  # box = dict(log2_cells, origin, h, dirichlet_faces) like mfg_box_desc
            lg = (C.c_int * 3)(*[int(x) for x in (list(box["log2_cells"]) + [0, 0, 0])[:3]])
            org = (C.c_double * 3)(*[float(x) for x in (list(box["origin"]) + [0.0, 0.0, 0.0])[:3]])
            self.h = self.L.orc_create_box(dim, p, lg, org, float(box["h"]), int(box.get("dirichlet_faces", 0x3f)))
            r = int(box["log2_cells"][0])
        if not self.h:
            raise ValueError("orc_create failed")
        self.dim, self.p, self.r, self.left, self.right = dim, p, r, left, right
        self.n_cells = self.L.orc_n_cells(self.h)
        self.n_dofs = self.L.orc_n_dofs(self.h)
        self.dofs_per_cell = self.L.orc_dofs_per_cell(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n_constrained(self):
        return self.L.orc_n_constrained(self.h)

    @property
    def loc2glob(self):
        return _view(self.L.orc_loc2glob(self.h), (self.n_cells, self.dofs_per_cell), np.uint32)

    @property
    def constrained(self):
        return _view(self.L.orc_constrained(self.h), (self.n_constrained,), np.uint32)

    @property
    def dof_lattice(self):
        return _view(self.L.orc_dof_lattice(self.h), (self.n_dofs, 3), np.uint32)

    @property
    def cell_coords(self):
        return _view(self.L.orc_cell_coords(self.h), (self.n_cells, 3), np.uint32)

    @property
    def coefficient(self):
        return _view(self.L.orc_coefficient(self.h), (self.n_cells, self.dofs_per_cell), np.float64)

    @property
    def shape_values(self):
        n = self.p + 1
        return _view(self.L.orc_shape_values(self.h), (n, n), np.float64)

    @property
    def shape_gradients(self):
        n = self.p + 1
        return _view(self.L.orc_shape_gradients(self.h), (n, n), np.float64)

    def set_constant_coefficient(self, a):
        self.L.orc_set_constant_coefficient(self.h, a)

    def clear_constraints(self):
        self.L.orc_clear_constraints(self.h)

    def vmult(self, src, threaded=False, fast=False):
        """scalar reference restatement; threaded=True: same arithmetic, OpenMP over colors; fast=True: the timed
        CPU baseline (collocation form, 8-cell SIMD batches, OpenMP) -- same operator, different summation order"""
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.empty_like(src)
        f = self.L.orc_vmult_fast if fast else (self.L.orc_vmult_omp if threaded else self.L.orc_vmult)
        f(self.h, _dp(dst), _dp(src))
        return dst

    def cell_loop_range(self, src, cell_begin, cell_end):
        """partial sums of the cells [cell_begin, cell_end) (no constrained-row identity)"""
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.zeros_like(src)
        self.L.orc_cell_loop_range(self.h, _dp(dst), _dp(src), int(cell_begin), int(cell_end))
        return dst

    def vmult_add(self, dst, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.array(dst, dtype=np.float64, copy=True)
        self.L.orc_vmult_add(self.h, _dp(dst), _dp(src))
        return dst

    def bmop(self, k, init=0.1):
        out = np.empty(self.n_dofs)
        self.L.orc_bmop(self.h, k, init, _dp(out))
        return out

    def inverse_diagonal(self):
        out = np.empty(self.n_dofs)
        self.L.orc_inverse_diagonal(self.h, _dp(out))
        return out

    def assemble_dense(self):
        K = np.empty((self.n_dofs, self.n_dofs))
        if self.L.orc_assemble_dense(self.h, _dp(K)) != 0:
            raise MemoryError("mesh too large for dense assembly")
        return K
