"""CG solve: the native fused CG against a numpy restatement of deal.II's SolverCG control flow
 running on the CPU oracle operator."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import OracleMesh, sm64  # noqa: E402


def numpy_cg(o, b, tol, max_iter, jacobi):
    """SolverCG::solve as in deal.II 8.4 with x0 = 0 (poisson.cu:247-260)"""
    minv = o.inverse_diagonal() if jacobi else np.ones(o.n_dofs)
    x = np.zeros(o.n_dofs)
    g = -b.copy()
    h = minv * g
    d = -h
    gh = g @ h
    res = np.sqrt(g @ g)
    hist = [res]
    it = 0
    if res <= tol:
        return x, 0, hist
    for it in range(1, max_iter + 1):
        h = o.vmult(d)
        alpha = gh / (d @ h)
        x += alpha * d
        g += alpha * h
        res = np.sqrt(g @ g)
        hist.append(res)
        if res <= tol:
            break
        h = minv * g
        beta = gh
        gh = g @ h
        beta = gh / beta
        d = beta * d - h
    return x, it, hist


@pytest.mark.parametrize("dim,p,r,jacobi", [(2, 4, 3, True), (3, 2, 2, True), (3, 4, 2, True), (3, 4, 2, False), (3, 3, 3, True)])
def test_cg_matches_numpy_restatement(ctx, dim, p, r, jacobi):
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    u_exact = sm64(31, o.n_dofs)
    b = o.vmult(u_exact)                        # consistent right-hand side (constrained rows: b = u)
    tol = 1e-10 * np.linalg.norm(b)
    xr, itr, hr = numpy_cg(o, b, tol, 2000, jacobi)
    m = mf.HyperCubeMesh(ctx, dim, p, r)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    x, vb = mf.GpuVector(ctx, o.n_dofs), mf.GpuVector.from_numpy(ctx, b)
    it, res, hist = mf.solver_cg(op, x, vb, tol, 2000, use_jacobi=jacobi, history=True)
    assert abs(it - itr) <= 1, (it, itr)
    # CG iterates are roundoff-sensitive once orthogonality degrades: the first iterations are compared tightly, the
    # whole run through the iteration count, the final residual and the solution
    k = min(15, it, itr)
    assert np.allclose(hist[:k], hr[:k], rtol=1e-9)
    got = x.toVector()
    assert np.linalg.norm(got - u_exact) <= 1e-8 * np.linalg.norm(u_exact)
    assert np.linalg.norm(got - xr) <= 1e-8 * np.linalg.norm(xr)
    assert res <= tol


def test_cg_fp32_and_nonzero_start(ctx):
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, 3, 2)
    u_exact = sm64(5, o.n_dofs)
    b = o.vmult(u_exact)
    m = mf.HyperCubeMesh(ctx, 3, 3, 2)
    op = mf.LaplaceOperatorGpu(ctx, np.float32)
    op.reinit(m)
    x = mf.GpuVector.from_numpy(ctx, (0.5 * u_exact).astype(np.float32))   # non-zero start: g = A x - b
    vb = mf.GpuVector.from_numpy(ctx, b.astype(np.float32))
    it, res = mf.solver_cg(op, x, vb, 1e-4 * np.linalg.norm(b), 500)
    assert 0 < it < 500
    assert np.linalg.norm(x.toVector() - u_exact) <= 1e-3 * np.linalg.norm(u_exact)


@pytest.mark.parametrize("p,r", [(4, 2), (3, 3), (2, 3), (5, 1)])
def test_cg_fused_loop_equals_unfused_loop(ctx, p, r):
    """the fused loop (d . A d emitted by the cell kernel, the operator's zero pass done by cg_advance; solver.cu) against the
    loop with vmult + cg_dot: same iteration count, residual histories equal to rounding, and the first iterates within
    1e-12 of the numpy restatement of SolverCG on the oracle operator"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, p, r)
    u_exact = sm64(13, o.n_dofs)
    b = o.vmult(u_exact)
    tol = 1e-10 * np.linalg.norm(b)
    m = mf.HyperCubeMesh(ctx, 3, p, r)
    runs = []
    for unfused in (False, True):
        op = mf.LaplaceOperatorGpu(ctx, np.float64)
        op.reinit(m)
        op.set_option("cg_fused", 0 if unfused else 1)
        assert op.active_variant() == 50
        x, vb = mf.GpuVector(ctx, o.n_dofs), mf.GpuVector.from_numpy(ctx, b)
        it, res, hist = mf.solver_cg(op, x, vb, tol, 2000, history=True)
        runs.append((it, np.asarray(hist), x.toVector()))
    (it_f, h_f, x_f), (it_u, h_u, x_u) = runs
    assert abs(it_f - it_u) <= 1
    k = min(12, it_f, it_u)
    assert np.allclose(h_f[:k], h_u[:k], rtol=1e-11)
    assert np.linalg.norm(x_f - x_u) <= 1e-9 * np.linalg.norm(x_u)
    xr, itr, hr = numpy_cg(o, b, tol, 2000, True)
    assert np.allclose(h_f[:6], hr[:6], rtol=1e-12)
    assert np.linalg.norm(x_f - u_exact) <= 1e-8 * np.linalg.norm(u_exact)
