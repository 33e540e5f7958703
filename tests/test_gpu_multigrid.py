"""MGTransferMatrixFreeGpu and the V-cycle.  The reference compares its transfer with deal.II's CPU
MGTransferMatrixFree (test_mg_transfer.cc:88-165), unavailable here; the restatable invariants are used instead:
prolongation = interpolation of the coarse FE function (checked against a geometric numpy matrix and through exact
reproduction of polynomials of degree <= p), restriction = prolongation^T, Dirichlet rows/columns zero."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import OracleMesh, shape_1d, sm64  # noqa: E402


def numpy_prolongation(oc, of):
    """P (n_fine x n_coarse): coarse FE function evaluated at every fine support point; coarse Dirichlet columns = 0"""
    dim, p, n = oc.dim, oc.p, oc.p + 1
    _, _, xn, _, _ = shape_1d(p)
    P = np.zeros((of.n_dofs, oc.n_dofs))
    lat = of.dof_lattice            # fine lattice coordinates 0..p*Nf
    Nc = 1 << oc.r
    # physical coordinate of a fine dof in units of coarse cells
    def coord(X):
        c, i = divmod(int(X), p)
        if c == 2 * Nc:
            c, i = c - 1, p
        return (c + xn[i]) / 2.0
    cc = {tuple(int(v) for v in oc.cell_coords[c][:dim]): c for c in range(oc.n_cells)}
    for g in range(of.n_dofs):
        x = [coord(lat[g, d]) for d in range(dim)]
        cell = tuple(min(int(np.floor(v)), Nc - 1) for v in x)
        xi = [x[d] - cell[d] for d in range(dim)]
        ci = cc[cell]
        for li in range(n ** dim):
            idx = [(li // n ** d) % n for d in range(dim)]
            val = 1.0
            for d in range(dim):
                val *= np.prod([(xi[d] - xn[m]) / (xn[idx[d]] - xn[m]) for m in range(n) if m != idx[d]])
            if abs(val) > 1e-15:
                P[g, oc.loc2glob[ci, li]] = val
    P[:, oc.constrained] = 0.0
    return P


@pytest.mark.parametrize("dim,p,lc", [(2, 1, 2), (2, 3, 1), (2, 4, 2), (3, 1, 1), (3, 2, 1), (3, 4, 1), (3, 3, 2)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_transfer_matches_geometric_matrix(ctx, dim, p, lc, dtype):
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import MGTransferMatrixFreeGpu
    tol = 1e-13 if dtype == np.float64 else 1e-5
    oc, of = OracleMesh(dim, p, lc), OracleMesh(dim, p, lc + 1)
    P = numpy_prolongation(oc, of)
    meshes = {lc: mf.HyperCubeMesh(ctx, dim, p, lc), lc + 1: mf.HyperCubeMesh(ctx, dim, p, lc + 1)}
    tr = MGTransferMatrixFreeGpu(ctx, dtype)
    tr.build(meshes)
    uc = sm64(1, oc.n_dofs).astype(dtype)
    src, dst = mf.GpuVector.from_numpy(ctx, uc), mf.GpuVector(ctx, of.n_dofs, dtype)
    dst.fill(7.0)                                         # prolongate overwrites
    tr.prolongate(lc + 1, dst, src)
    want = P @ uc.astype(np.float64)
    assert np.linalg.norm(dst.toVector() - want) <= tol * np.linalg.norm(want)
    # restrict_and_add into a pre-filled vector (test_mg_transfer.cc:156-165)
    rf = sm64(2, of.n_dofs).astype(dtype); d0 = sm64(3, oc.n_dofs).astype(dtype)
    vsrc, vdst = mf.GpuVector.from_numpy(ctx, rf), mf.GpuVector.from_numpy(ctx, d0)
    tr.restrict_and_add(lc + 1, vdst, vsrc)
    want = d0.astype(np.float64) + P.T @ rf.astype(np.float64)
    assert np.linalg.norm(vdst.toVector() - want) <= tol * np.linalg.norm(want)
    assert np.array_equal(vdst.toVector()[oc.constrained], d0[oc.constrained])   # Dirichlet rows untouched


def test_prolongation_reproduces_polynomials(ctx):
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import MGTransferMatrixFreeGpu
    dim, p, lc = 3, 3, 1
    oc, of = OracleMesh(dim, p, lc), OracleMesh(dim, p, lc + 1)
    _, _, xn, _, _ = shape_1d(p)

    def coords(o):
        N = 1 << o.r
        lat = o.dof_lattice
        c = np.minimum(lat // p, N - 1); i = lat - c * p
        return (c + xn[i]) / N                          # in [0,1]^3
    f = lambda x: (x[:, 0] ** 3 - 2 * x[:, 1] ** 2 * x[:, 2] + x[:, 0] * x[:, 1] * x[:, 2] + 1.0)   # degree <= 3 per variable
    uc, uf = f(coords(oc)), f(coords(of))
    meshes = {lc: mf.HyperCubeMesh(ctx, dim, p, lc), lc + 1: mf.HyperCubeMesh(ctx, dim, p, lc + 1)}
    tr = MGTransferMatrixFreeGpu(ctx); tr.build(meshes)
    src, dst = mf.GpuVector.from_numpy(ctx, uc), mf.GpuVector(ctx, of.n_dofs)
    tr.prolongate(lc + 1, dst, src)
    got = dst.toVector()
    # exact away from the Dirichlet boundary of the COARSE level (those coarse DoFs are read as 0)
    lat = of.dof_lattice
    M = p * (1 << of.r)
    interior = np.all((lat[:, :dim] >= 2 * p) & (lat[:, :dim] <= M - 2 * p), axis=1)
    assert interior.sum() > 0
    assert np.abs(got[interior] - uf[interior]).max() <= 1e-13


@pytest.mark.parametrize("dim,p", [(2, 2), (3, 2), (3, 4)])
def test_multigrid_preconditioned_cg(ctx, dim, p):
    """iteration counts of MG-preconditioned CG stay bounded under refinement (the property multigrid is for)"""
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import GeometricMultigrid, solver_cg_preconditioned
    its = []
    for top in ((3, 4, 5) if dim == 2 else (2, 3)):
        mg = GeometricMultigrid(ctx, dim, p, 1, top)
        op = mg.ops[top]
        n = op.m()
        ue = mf.GpuVector.from_numpy(ctx, sm64(9, n))
        b, x = mf.GpuVector(ctx, n), mf.GpuVector(ctx, n)
        op.vmult(b, ue)
        it, hist = solver_cg_preconditioned(ctx, op, x, b, mg, 1e-10 * b.l2_norm(), 100)
        x.add(-1.0, ue)
        assert x.l2_norm() <= 1e-7 * ue.l2_norm()
        its.append(it)
    assert max(its) <= 25 and max(its) - min(its) <= 6, its


def numpy_chebyshev(o, degree, smoothing_range, n_eig):
    """deal.II 8.5 PreconditionChebyshev restated on the oracle operator: eigenvalue estimate by n_eig
    SolverCG steps on Dinv A from x = 0 with rhs = 1/sqrt(n), entry 0 zeroed; Lanczos matrix from the CG coefficients;
    beta = 1.2 lmax, alpha = lmax / range; returns (lmax, theta, delta, apply(b, x0=None))"""
    n = o.n_dofs
    dinv = o.inverse_diagonal()
    rhs = np.full(n, 1.0 / np.sqrt(n)); rhs[0] = 0.0
    g = -rhs; h = dinv * g; d = -h; gh = g @ h
    al, be = [], []
    for _ in range(n_eig):
        if np.sqrt(g @ g) <= 1e-2:
            break
        h = o.vmult(d)
        alpha = gh / (d @ h); al.append(alpha)
        g = g + alpha * h
        h = dinv * g
        ghn = g @ h; beta = ghn / gh; gh = ghn; be.append(beta)
        d = beta * d - h
    k = len(al)
    T = np.zeros((k, k))
    for j in range(k):
        T[j, j] = 1.0 / al[j] + (be[j - 1] / al[j - 1] if j else 0.0)
        if j + 1 < k:
            T[j, j + 1] = T[j + 1, j] = np.sqrt(be[j]) / al[j]
    lmax = np.linalg.eigvalsh(T)[-1]
    b_, a_ = 1.2 * lmax, lmax / smoothing_range
    theta, delta = 0.5 * (b_ + a_), 0.5 * (b_ - a_)

    def apply(b, x0=None):
        rhok, sigma = delta / theta, theta / delta
        if x0 is None:
            dvec = dinv * b / theta
            x = dvec.copy()
        else:
            dvec = dinv * (b - o.vmult(x0)) / theta
            x = x0 + dvec
        for _ in range(degree):
            r = b - o.vmult(x)
            rhokp = 1.0 / (2.0 * sigma - rhok)
            dvec = rhokp * rhok * dvec + 2.0 * rhokp / delta * (dinv * r)
            rhok = rhokp
            x = x + dvec
        return x
    return lmax, theta, delta, apply


@pytest.mark.parametrize("dim,p,r", [(2, 3, 3), (3, 2, 2), (3, 4, 2)])
def test_chebyshev_matches_restatement_of_dealii(ctx, dim, p, r):
    """mfg_chebyshev_* (csrc/multigrid.cu) against the numpy restatement on the oracle: the eigenvalue estimate, the
    polynomial from a zero guess (vmult) and from a given guess (step)"""
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import ChebyshevSmoother
    o = OracleMesh(dim, p, r)
    lmax, theta, delta, apply = numpy_chebyshev(o, 5, 15.0, 15)
    m = mf.HyperCubeMesh(ctx, dim, p, r)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    sm = ChebyshevSmoother(ctx, op, 5, 15.0, 15)
    assert abs(sm.lambda_max - lmax) <= 1e-9 * lmax
    assert abs(sm.theta - theta) <= 1e-9 * theta and abs(sm.delta - delta) <= 1e-9 * delta
    b = sm64(3, o.n_dofs)
    vb, vx = mf.GpuVector.from_numpy(ctx, b), mf.GpuVector(ctx, o.n_dofs)
    vx.fill(99.0)
    sm.vmult(vx, vb)
    want = apply(b)
    assert np.linalg.norm(vx.toVector() - want) <= 1e-11 * np.linalg.norm(want)
    x0 = sm64(4, o.n_dofs)
    vx.fromHost(x0)
    sm.step(vx, vb)
    want = apply(b, x0)
    assert np.linalg.norm(vx.toVector() - want) <= 1e-11 * np.linalg.norm(want)


def test_library_mg_cg_equals_host_orchestrated_loop(ctx):
    """mfg_mg_solve_cg (C++ loop) against solver_cg_preconditioned (Python loop over the same V-cycle): same iterations"""
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import GeometricMultigrid, solver_cg_preconditioned
    mg = GeometricMultigrid(ctx, 3, 4, 1, 3)
    op = mg.ops[3]
    n = op.m()
    ue = mf.GpuVector.from_numpy(ctx, sm64(9, n))
    b, x1, x2 = mf.GpuVector(ctx, n), mf.GpuVector(ctx, n), mf.GpuVector(ctx, n)
    op.vmult(b, ue)
    tol = 1e-10 * b.l2_norm()
    it1, hist1 = solver_cg_preconditioned(ctx, op, x1, b, mg, tol, 100)
    it2, res2, hist2 = mg.solve_cg(x2, b, tol, 100, history=True)
    assert it1 == it2 and it2 <= 12
    assert np.allclose(hist1, hist2, rtol=1e-6)
    x2.add(-1.0, ue)
    assert x2.l2_norm() <= 1e-7 * ue.l2_norm()
    assert mg.coarse_iterations > 0 and all(mg.lambda_max[l] > 1.0 for l in mg.levels[1:])
