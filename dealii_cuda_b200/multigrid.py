"""Geometric multigrid on globally refined meshes: MGTransferMatrixFreeGpu, level operators, Chebyshev smoother,
V-cycle and a preconditioned CG driver.  In the reference this orchestration lives in deal.II templates
(Multigrid, PreconditionMG, PreconditionChebyshev, SolverCG instantiated on GpuVector, poisson_mg.cu:430-552,
bmop_mg.cu:300-340); here it is host-side Python over the C ABI, every vector operation a kernel of libmfgpu.so."""
import ctypes as C

import numpy as np

from . import GpuVector, HyperCubeMesh, LaplaceOperatorGpu, _capi, check, lib, solver_cg


class MGTransferMatrixFreeGpu:
    """MGTransferMatrixFreeGpu<dim,Number> (mg_transfer_matrix_free_gpu.h:64-307) for a globally refined hierarchy.
    Levels are numbered like deal.II's: level l has 2^l cells per direction."""

    def __init__(self, ctx, dtype=np.float64):
        self.ctx, self.code = ctx, _capi.F64 if np.dtype(dtype) == np.float64 else _capi.F32
        self.h = {}
        self.meshes = None

    def build(self, level_meshes):
        """level_meshes: dict level -> HyperCubeMesh (consecutive levels)."""
        self.clear()
        self.meshes = level_meshes
        levels = sorted(level_meshes)
        for lc, lf in zip(levels[:-1], levels[1:]):
            assert lf == lc + 1
            h = C.c_void_p()
            check(lib.mfg_mgt_build(self.ctx.h, level_meshes[lc].h, level_meshes[lf].h, self.code, C.byref(h)))
            self.h[lf] = h

    def clear(self):
        for h in self.h.values():
            lib.mfg_mgt_destroy(h)
        self.h = {}

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass

    def prolongate(self, to_level, dst, src):
        check(lib.mfg_mgt_prolongate(self.h[to_level], dst.h, src.h))

    def restrict_and_add(self, from_level, dst, src):
        check(lib.mfg_mgt_restrict_and_add(self.h[from_level], dst.h, src.h))

    def copy_to_mg(self, dst_levels, src):
        """on a globally refined mesh the finest level is the active mesh: plain copy (mg_transfer...cu:688-727)"""
        dst_levels[max(dst_levels)].assign(src)

    def copy_from_mg(self, dst, src_levels):
        dst.assign(src_levels[max(src_levels)])


class ChebyshevSmoother:
    """PreconditionChebyshev on Dinv A (poisson_mg.cu:461-470: degree 5, smoothing range 15, 15 CG/Lanczos iterations for the
    eigenvalue estimate, deal.II's procedure): binding of mfg_chebyshev_* (csrc/multigrid.cu)."""

    def __init__(self, ctx, op, degree=5, smoothing_range=15.0, eig_iterations=15, dtype=np.float64):
        self.ctx, self.op, self.degree = ctx, op, degree
        h = C.c_void_p()
        check(lib.mfg_chebyshev_create(op.h, int(degree), float(smoothing_range), int(eig_iterations), C.byref(h)))
        self.h = h
        lmax, lmin, theta, delta, its = C.c_double(), C.c_double(), C.c_double(), C.c_double(), C.c_int()
        check(lib.mfg_chebyshev_info(h, C.byref(lmax), C.byref(lmin), C.byref(theta), C.byref(delta), C.byref(its)))
        self.lambda_max, self.lambda_min, self.theta, self.delta, self.eig_iterations = lmax.value, lmin.value, theta.value, delta.value, its.value

    def __del__(self):
        try:
            if self.h:
                lib.mfg_chebyshev_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def vmult(self, dst, src):
        check(lib.mfg_chebyshev_vmult(self.h, dst.h, src.h))

    def step(self, x, b, zero_guess=False):
        """one Chebyshev sweep of the given degree on A x = b (zero_guess: PreconditionChebyshev::vmult, else ::step)"""
        check((lib.mfg_chebyshev_vmult if zero_guess else lib.mfg_chebyshev_step)(self.h, x.h, b.h))


class _LevelOperator:
    """borrowed level operator of a GeometricMultigrid (same interface as LaplaceOperatorGpu for vmult / m)"""

    def __init__(self, ctx, handle, n, dtype):
        self.ctx, self.h, self._n, self.dtype = ctx, handle, n, dtype

    def m(self):
        return self._n

    n = m

    def vmult(self, dst, src):
        check(lib.mfg_laplace_vmult(self.h, dst.h, src.h))


class GeometricMultigrid:
    """V-cycle preconditioner (Multigrid + PreconditionMG, poisson_mg.cu:456-518) on hyper_cube meshes
    refine_global(min_level..max_level): binding of mfg_mg_* (csrc/multigrid.cu)."""

    def __init__(self, ctx, dim, degree, min_level, max_level, dtype=np.float64, left=-1.0, right=1.0, smoother_degree=5,
                 smoothing_range=15.0, eig_iterations=15):
        self.ctx, self.dtype = ctx, dtype
        self.levels = list(range(min_level, max_level + 1))
        code = _capi.F64 if np.dtype(dtype) == np.float64 else _capi.F32
        h = C.c_void_p()
        check(lib.mfg_mg_create(ctx.h, dim, degree, min_level, max_level, code, float(left), float(right), int(smoother_degree),
                                float(smoothing_range), int(eig_iterations), C.byref(h)))
        self.h = h
        self.ops, self.lambda_max = {}, {}
        for l in self.levels:
            oh, lm, ci, nd = C.c_void_p(), C.c_double(), C.c_long(), C.c_size_t()
            check(lib.mfg_mg_level_operator(h, l, C.byref(oh)))
            check(lib.mfg_mg_info(h, l, C.byref(lm), C.byref(ci), C.byref(nd)))
            self.ops[l] = _LevelOperator(ctx, oh, nd.value, dtype)
            self.lambda_max[l] = lm.value

    def __del__(self):
        try:
            if self.h:
                lib.mfg_mg_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def coarse_iterations(self):
        ci = C.c_long()
        check(lib.mfg_mg_info(self.h, self.levels[0], None, C.byref(ci), None))
        return ci.value

    def vmult(self, dst, src):
        """PreconditionMG::vmult: copy_to_mg, one V-cycle, copy_from_mg"""
        check(lib.mfg_mg_vcycle(self.h, dst.h, src.h))

    def solve_cg(self, x, b, abs_tol, max_iter=1000, history=False):
        """SolverCG on the finest level preconditioned by the V-cycle, in the library (poisson_mg.cu:504-518)"""
        its, res = C.c_int(), C.c_double()
        hist = (C.c_double * (max_iter + 1))() if history else None
        check(lib.mfg_mg_solve_cg(self.h, x.h, b.h, float(abs_tol), int(max_iter), C.byref(its), C.byref(res), hist))
        return (its.value, res.value, list(hist[:its.value + 1])) if history else (its.value, res.value)


class AdaptiveMultigrid:
    """Multigrid with local smoothing on an adaptively refined mesh (poisson_mg.cu / bmop_mg.cu with an adaptive grid): binding of
    mfg_amg_* (csrc/multigrid.cu) over the host hierarchy of AdaptiveMesh.build_mg (csrc/adaptive_mesh.cu).  The mesh must have
    been created with limit_level_difference_at_vertices=True like the reference's Triangulation (poisson_mg.cu:132)."""

    def __init__(self, ctx, amesh, min_level=0, dtype=np.float64, smoother_degree=5, smoothing_range=15.0, eig_iterations=15):
        self.ctx, self.dtype, self.amesh = ctx, dtype, amesh
        code = _capi.F64 if np.dtype(dtype) == np.float64 else _capi.F32
        h = C.c_void_p()
        check(lib.mfg_amg_create(ctx.h, amesh.h, int(min_level), code, int(smoother_degree), float(smoothing_range), int(eig_iterations), C.byref(h)))
        self.h = h
        self.levels = list(range(min_level, amesh.n_levels))
        oh = C.c_void_p()
        check(lib.mfg_amg_active_operator(h, C.byref(oh)))
        self.op = _LevelOperator(ctx, oh, amesh.n_dofs, dtype)          # the operator on the active mesh (hanging nodes resolved)
        self.ops, self.lambda_max, self.n_dofs, self.n_edge = {}, {}, {}, {}
        for l in self.levels:
            oh, lm, ci, nd, ne = C.c_void_p(), C.c_double(), C.c_long(), C.c_size_t(), C.c_size_t()
            check(lib.mfg_amg_level_operator(h, l, C.byref(oh)))
            check(lib.mfg_amg_info(h, l, C.byref(lm), C.byref(ci), C.byref(nd), C.byref(ne)))
            self.ops[l] = _LevelOperator(ctx, oh, nd.value, dtype)
            self.lambda_max[l], self.n_dofs[l], self.n_edge[l] = lm.value, nd.value, ne.value

    def __del__(self):
        try:
            if self.h:
                lib.mfg_amg_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def coarse_iterations(self):
        ci = C.c_long()
        check(lib.mfg_amg_info(self.h, self.levels[0], None, C.byref(ci), None, None))
        return ci.value

    def vmult(self, dst, src):
        """PreconditionMG::vmult on vectors of the active mesh: copy_to_mg, one V-cycle with the edge matrices, copy_from_mg"""
        check(lib.mfg_amg_vcycle(self.h, dst.h, src.h))

    def vmult_interface_down(self, level, dst, src):
        check(lib.mfg_amg_vmult_interface_down(self.h, int(level), dst.h, src.h))

    def vmult_interface_up(self, level, dst, src):
        check(lib.mfg_amg_vmult_interface_up(self.h, int(level), dst.h, src.h))

    def prolongate(self, to_level, dst, src):
        check(lib.mfg_amg_prolongate(self.h, int(to_level), dst.h, src.h))

    def restrict_and_add(self, from_level, dst, src):
        check(lib.mfg_amg_restrict_and_add(self.h, int(from_level), dst.h, src.h))

    def copy_to_level(self, level, dst, src):
        check(lib.mfg_amg_copy_to_level(self.h, int(level), dst.h, src.h))

    def copy_from_level(self, level, dst, src):
        check(lib.mfg_amg_copy_from_level(self.h, int(level), dst.h, src.h))

    def solve_cg(self, x, b, abs_tol, max_iter=1000, history=False):
        """SolverCG on the active operator preconditioned by the V-cycle (poisson_mg.cu:504-518)"""
        its, res = C.c_int(), C.c_double()
        hist = (C.c_double * (max_iter + 1))() if history else None
        check(lib.mfg_amg_solve_cg(self.h, x.h, b.h, float(abs_tol), int(max_iter), C.byref(its), C.byref(res), hist))
        return (its.value, res.value, list(hist[:its.value + 1])) if history else (its.value, res.value)


def solver_cg_preconditioned(ctx, op, x, b, precond, abs_tol, max_iter=1000):
    """SolverCG control flow with an arbitrary preconditioner object (vmult(dst, src)); host-side
    variant for preconditioners written in Python (the library's own is GeometricMultigrid.solve_cg)."""
    n, dtype = op.m(), x.dtype
    g, h, d = (GpuVector(ctx, n, dtype) for _ in range(3))
    if x.all_zero():
        g.equ(-1.0, b)
    else:
        op.vmult(g, x); g.sadd(1.0, -1.0, b)
    res = g.l2_norm()
    hist = [res]
    if res <= abs_tol:
        return 0, hist
    precond.vmult(h, g)
    d.equ(-1.0, h)
    gh = g.dot(h)
    for it in range(1, max_iter + 1):
        op.vmult(h, d)
        alpha = gh / d.dot(h)
        x.add(alpha, d)
        res = np.sqrt(g.add_and_dot(alpha, h, g))
        hist.append(res)
        if res <= abs_tol:
            return it, hist
        precond.vmult(h, g)
        beta = gh
        gh = g.dot(h)
        beta = gh / beta
        d.sadd(beta, -1.0, h)
    return max_iter, hist
