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
    """PreconditionChebyshev on D^-1 A (poisson_mg.cu:461-470: degree 5, smoothing range 15, 15 iterations for the
    eigenvalue estimate).  lambda_max is estimated by power iteration on D^-1 A with a 1.2 safety factor."""

    def __init__(self, ctx, op, degree=5, smoothing_range=15.0, eig_iterations=15, dtype=np.float64):
        self.ctx, self.op, self.degree = ctx, op, degree
        n = op.m()
        op.compute_diagonal()
        self.dinv = op.get_diagonal_inverse()
        self.r, self.d, self.t = (GpuVector(ctx, n, dtype) for _ in range(3))
        v = GpuVector.from_numpy(ctx, (1.0 + (np.arange(n) % 11) / 11.0).astype(dtype))
        lam = 1.0
        for _ in range(eig_iterations):
            op.vmult(self.t, v)
            self.t.scale(self.dinv)
            lam = self.t.l2_norm() / v.l2_norm()
            v.equ(1.0 / self.t.l2_norm(), self.t)
        self.lambda_max = lam
        beta, alpha = 1.2 * lam, 1.2 * lam / smoothing_range
        self.theta, self.delta = 0.5 * (beta + alpha), 0.5 * (beta - alpha)

    def step(self, x, b, zero_guess):
        """one Chebyshev sweep of the given degree on A x = b"""
        op, r, d, t = self.op, self.r, self.d, self.t
        theta, delta = self.theta, self.delta
        sigma = theta / delta
        rho = 1.0 / sigma
        if zero_guess:
            r.assign(b)
        else:
            op.vmult(r, x)
            r.sadd(-1.0, 1.0, b)                    # r = b - A x
        d.assign(r); d.scale(self.dinv); d *= 1.0 / theta
        if zero_guess:
            x.assign(d)
        else:
            x.add(d)
        for _ in range(self.degree):
            op.vmult(r, x)
            r.sadd(-1.0, 1.0, b)
            rho_new = 1.0 / (2.0 * sigma - rho)
            t.assign(r); t.scale(self.dinv)
            d.sadd(rho_new * rho, 2.0 * rho_new / delta, t)
            x.add(d)
            rho = rho_new


class GeometricMultigrid:
    """V-cycle preconditioner (Multigrid + PreconditionMG, poisson_mg.cu:456-518) on hyper_cube meshes
    refine_global(min_level..max_level)."""

    def __init__(self, ctx, dim, degree, min_level, max_level, dtype=np.float64, left=-1.0, right=1.0, smoother_degree=5):
        self.ctx, self.dtype = ctx, dtype
        self.levels = list(range(min_level, max_level + 1))
        self.meshes = {l: HyperCubeMesh(ctx, dim, degree, l, left, right) for l in self.levels}
        self.ops = {}
        for l in self.levels:
            self.ops[l] = LaplaceOperatorGpu(ctx, dtype)
            self.ops[l].reinit(self.meshes[l])
        self.transfer = MGTransferMatrixFreeGpu(ctx, dtype)
        self.transfer.build(self.meshes)
        self.smoothers = {l: ChebyshevSmoother(ctx, self.ops[l], smoother_degree, dtype=dtype) for l in self.levels[1:]}
        self.x = {l: GpuVector(ctx, self.meshes[l].n_dofs, dtype) for l in self.levels}
        self.b = {l: GpuVector(ctx, self.meshes[l].n_dofs, dtype) for l in self.levels}
        self.t = {l: GpuVector(ctx, self.meshes[l].n_dofs, dtype) for l in self.levels}
        self.ops[min_level].compute_diagonal()
        self.coarse_iterations = 0

    def _cycle(self, l):
        x, b, t, op = self.x[l], self.b[l], self.t[l], self.ops[l]
        if l == self.levels[0]:
            x.fill(0.0)                                                       # coarse CG to 1e-10 (poisson_mg.cu:73-80)
            it, _ = solver_cg(op, x, b, 1e-10 * max(b.l2_norm(), 1e-300), 2000, use_jacobi=True)
            self.coarse_iterations += it
            return
        self.smoothers[l].step(x, b, zero_guess=True)                         # pre-smoothing
        op.vmult(t, x)
        t.sadd(-1.0, 1.0, b)                                                  # residual
        self.b[l - 1].fill(0.0)
        self.transfer.restrict_and_add(l, self.b[l - 1], t)
        self._cycle(l - 1)
        self.transfer.prolongate(l, t, self.x[l - 1])
        x.add(t)
        self.smoothers[l].step(x, b, zero_guess=False)                        # post-smoothing

    def vmult(self, dst, src):
        """PreconditionMG::vmult: copy_to_mg, one V-cycle, copy_from_mg"""
        top = self.levels[-1]
        self.transfer.copy_to_mg({top: self.b[top]}, src)
        self._cycle(top)
        self.transfer.copy_from_mg(dst, {top: self.x[top]})


def solver_cg_preconditioned(ctx, op, x, b, precond, abs_tol, max_iter=1000):
    """SolverCG control flow with an arbitrary preconditioner object (vmult(dst, src))."""
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
