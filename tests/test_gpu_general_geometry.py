"""Non-affine geometry (MFG_GEOM_GENERAL): full inverse Jacobian per quadrature point.

Checker: an independent numpy restatement of the reference's general path -- FEEvaluationGpu::get_gradient applies
K^T = J^-T, submit_gradient applies K and JxW (fee_gpu.cuh:219-246, 261-284), K in FEValues::get_inverse_jacobians
order (matrix_free_gpu.cu:326-338) -- as dense cell matrices B_q^T (a JxW K K^T) B_q on a smoothly deformed cube
(the reference's own non-affine case is BALL_GRID, poisson_common.h:65-70).  Topology, DoF numbering and constraints
come from the uniform-mesh oracle; the deformation only changes the metric.
"""
import numpy as np
import pytest

from oracle.oracle import OracleMesh, sm64

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dealii_cuda_b200 as mf
    return mf.Context(0, None)


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300)


def deformation(dim, eps):
    """x = X + eps * s(X): smooth, nonlinear, invertible for small eps; returns (map, Jacobian dx/dX)."""
    def phi(X):
        x = X.copy()
        x[..., 0] += eps * np.sin(1.3 * X[..., 1] + 0.4) * np.cos(0.7 * X[..., dim - 1])
        x[..., 1] += eps * np.sin(0.9 * X[..., 0] - 0.2) * (1.0 + 0.5 * X[..., dim - 1])
        if dim == 3:
            x[..., 2] += eps * np.cos(1.1 * X[..., 0]) * np.sin(0.8 * X[..., 1] + 0.3)
        return x

    def jac(X):
        # central differences of the analytic map are exact enough for building test DATA (the same J feeds both sides)
        J = np.zeros(X.shape + (dim,))
        if eps == 0.0:
            J[..., np.arange(dim), np.arange(dim)] = 1.0
            return J
        hfd = 1e-6
        for e in range(dim):
            dX = np.zeros(dim); dX[e] = hfd
            J[..., :, e] = (phi(X + dX) - phi(X - dX)) / (2 * hfd)
        return J
    return phi, jac


def geometry(o, dim, p, eps, left=-1.0, right=1.0):
    """per cell and quadrature point: K = (d x / d xi)^-1 [d1][d2], JxW, quadrature point x_q"""
    import dealii_cuda_b200 as mf
    n = p + 1
    _, _, xq, wq = mf.shape_info(p)
    ncell_1d = round(o.n_cells ** (1.0 / dim))
    h = (right - left) / ncell_1d
    cc = np.asarray(o.cell_coords)[:, :dim].astype(np.float64)
    q = np.arange(n ** dim)
    q_idx = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)   # lexicographic: q -> (q0, q1[, q2]), q0 fastest
    X = left + h * (cc[:, None, :] + xq[q_idx][None, :, :])        # undeformed quadrature points [cell][q][dim]
    W = np.prod(wq[q_idx], axis=1)                                  # reference weights [q]
    phi, jac = deformation(dim, eps)
    J = jac(X) * h                                                  # d x / d xi  (xi in [0,1]^dim)
    K = np.linalg.inv(J)
    JxW = np.linalg.det(J) * W[None, :]
    return K, JxW, phi(X), q_idx


def reference_apply(o, dim, p, K, JxW, coef, q_idx, u):
    """dense restatement: dst = sum_cells P^T B^T diag(a JxW K K^T) B P u, constrained rows = identity"""
    n = p + 1
    N, Dn = np.asarray(o.shape_values), np.asarray(o.shape_gradients)   # [i][q]: phi_i(x_q), phi_i'(x_q)
    npc = n ** dim
    # reference-space gradient of basis function i at quadrature point q: B[d][q][i]
    B = np.zeros((dim, npc, npc))
    for d in range(dim):
        f = np.ones((npc, npc))
        for e in range(dim):
            M = Dn if e == d else N
            f *= M[q_idx[None, :, e], q_idx[:, None, e]]   # [q][i] -> M[i_e][q_e]
        B[d] = f
    l2g = np.asarray(o.loc2glob).astype(np.int64)
    con = np.zeros(o.n_dofs, bool)
    con[np.asarray(o.constrained)] = True
    uu = np.where(con, 0.0, u)
    dst = np.zeros(o.n_dofs)
    for c in range(o.n_cells):
        ul = uu[l2g[c]]
        g = np.einsum("dqi,i->qd", B, ul)                      # grad_xi at q
        gx = np.einsum("qed,qe->qd", K[c], g)                  # K^T g  (get_gradient)
        fl = gx * (coef[c] * JxW[c])[:, None]                  # a * JxW
        t = np.einsum("qde,qe->qd", K[c], fl)                  # K flux  (submit_gradient)
        rl = np.einsum("dqi,qd->i", B, t)
        np.add.at(dst, l2g[c], rl)
    dst[con] = u[con]
    return dst


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("coloring", [False, True])
@pytest.mark.parametrize("dim,p,r", [(2, 1, 2), (2, 2, 2), (2, 4, 2), (3, 1, 1), (3, 2, 1), (3, 3, 1), (3, 4, 1), (2, 7, 1), (3, 5, 1)])
def test_general_geometry_matches_dense_restatement(ctx, dim, p, r, coloring, dtype):
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    K, JxW, xq, q_idx = geometry(o, dim, p, eps=0.08)
    coef = 1.0 / (0.05 + 2.0 * (xq ** 2).sum(-1))               # Coefficient::value (poisson_common.h:155-157) at the deformed points
    l2g = np.asarray(o.loc2glob)
    data = dict(dim=dim, degree=p, n_dofs=o.n_dofs, loc2glob=l2g, inv_jac=K, JxW=JxW)
    if coloring:
        # parity coloring of the structured mesh: cells sorted by color
        cc = np.asarray(o.cell_coords)[:, :dim]
        color = (cc % 2 * (1 << np.arange(dim))).sum(1)
        order = np.argsort(color, kind="stable")
        l2g, K, JxW, coef = l2g[order], K[order], JxW[order], coef[order]
        data.update(loc2glob=l2g, inv_jac=K, JxW=JxW, color_offsets=np.concatenate([[0], np.cumsum(np.bincount(color, minlength=1 << dim))]))
    mfree = mf.MatrixFreeGpu(ctx, dtype)
    mfree.reinit(data, use_coloring=coloring)
    ch = mf.ConstraintHandlerGpu(ctx, dtype)
    ch.reinit(np.asarray(o.constrained), o.n_dofs)
    op = mf.LaplaceOperatorGpu(ctx, dtype, use_coloring=coloring)
    op.reinit(mfree, ch, coefficient=coef)
    assert op.active_variant() == 1
    u = sm64(5, o.n_dofs)
    # the dense restatement works on the (possibly permuted) cell arrays directly
    class O:  # view of the oracle with permuted cells
        pass
    ov = O()
    ov.n_cells, ov.n_dofs, ov.constrained, ov.loc2glob = o.n_cells, o.n_dofs, o.constrained, l2g
    ov.shape_values, ov.shape_gradients = o.shape_values, o.shape_gradients
    want = reference_apply(ov, dim, p, K, JxW, coef, q_idx, u)
    src, dst = mf.GpuVector.from_numpy(ctx, u.astype(dtype)), mf.GpuVector(ctx, o.n_dofs, dtype)
    dst.fill(3.0)
    op.vmult(dst, src)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    assert rel_err(dst.toVector(), want) <= tol
    # vmult_add and the diagonal
    d0 = sm64(6, o.n_dofs)
    dst.fromHost(d0.astype(dtype))
    op.vmult_add(dst, src)
    want_add = d0 + want
    con = np.asarray(o.constrained)
    assert rel_err(dst.toVector(), want_add) <= tol
    if dtype == np.float64 and o.n_dofs <= 400:
        A = np.stack([reference_apply(ov, dim, p, K, JxW, coef, q_idx, e) for e in np.eye(o.n_dofs)], axis=1)
        assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()   # symmetric
        op.compute_diagonal()
        inv_diag = op.get_diagonal_inverse().toVector()
        dd = np.diag(A).copy()
        dd[con] = 1.0
        assert rel_err(inv_diag, 1.0 / dd) <= 1e-12


def test_general_geometry_reduces_to_uniform(ctx):
    """eps = 0: the general path with K = I/h must reproduce the uniform-mesh operator (and the C oracle)"""
    import dealii_cuda_b200 as mf
    dim, p, r = 3, 3, 2
    o = OracleMesh(dim, p, r)
    K, JxW, xq, q_idx = geometry(o, dim, p, eps=0.0)
    data = dict(dim=dim, degree=p, n_dofs=o.n_dofs, loc2glob=np.asarray(o.loc2glob), inv_jac=K, JxW=JxW)
    mfree = mf.MatrixFreeGpu(ctx, np.float64)
    mfree.reinit(data)
    ch = mf.ConstraintHandlerGpu(ctx, np.float64)
    ch.reinit(np.asarray(o.constrained), o.n_dofs)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(mfree, ch, coefficient=np.asarray(o.coefficient))
    u = sm64(2, o.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs)
    op.vmult(dst, src)
    assert rel_err(dst.toVector(), o.vmult(u)) <= 1e-12
