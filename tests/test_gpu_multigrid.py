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
