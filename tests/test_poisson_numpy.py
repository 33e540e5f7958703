"""The reference's Poisson problem on LOCALLY REFINED meshes (poisson.cu with grid_refinement = NONUNIFORM, poisson_common.h:76-92),
restated in numpy on the oracle with the meshes of the library's host substrate: right-hand side with Dirichlet lifting assembled
through the hanging-node interpolation (rewritten loc2glob + masks), Jacobi-CG on the oracle operator, L2 error against the analytic
solution of poisson_common.cc.  Pins the whole hanging-node scheme to mathematics: the error falls like h^(p+1).  The device path
(examples/poisson.cu ... nonuniform) asserts the same in tests/late_gpu/."""
import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.adaptive import AdaptiveMesh, resolve_hanging_nodes

CENTERS = {2: np.array([[-0.5, 0.5], [-0.5, -0.5], [0.5, -0.5]]),
           3: np.array([[-0.5, 0.5, 0.25], [-0.6, -0.5, -0.125], [0.5, -0.5, 0.5]])}


def solution(x):
    """u, grad u, laplace u of Solution<dim> (poisson_common.cc:31-174): three Gaussians of width 1/3"""
    dim = x.shape[-1]
    w2 = 1.0 / 9.0
    norm = (np.sqrt(2 * np.pi) / 3) ** dim
    u, gu, lu = np.zeros(x.shape[:-1]), np.zeros(x.shape), np.zeros(x.shape[:-1])
    for c in CENTERS[dim]:
        t = x - c
        r2 = (t * t).sum(-1)
        e = np.exp(-r2 / w2) / norm
        u += e
        gu += -2 * t / w2 * e[..., None]
        lu += (-2 * dim / w2 + 4 * r2 / w2 ** 2) * e
    return u, gu, lu


def solve_poisson(dim, p, r):
    """(n_dofs, CG iterations, L2 error) on refine_global(r) + two refinements of the octant x_d > 0.2"""
    am = mf.AdaptiveMesh(dim, p).refine_global(r)
    for _ in range(2):
        am.mark_octant()
        am.execute_coarsening_and_refinement()
    am.distribute_dofs()
    a = am.arrays(quadrature_points=True)
    o = AdaptiveMesh(dim, p, 0, [], cells=am.active_cells().tolist())
    n, npc = p + 1, (p + 1) ** dim
    shape = (n,) * dim
    val, grad, _, wq = (np.asarray(t) for t in mf.shape_info(p))
    q = np.arange(npc)
    qi = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    Nq = np.ones((npc, npc))                                     # [q][i] = phi_i(x_q)
    for e in range(dim):
        Nq *= val[qi[None, :, e], qi[:, None, e]]
    G = []                                                       # [d][q][i] = reference-cell derivative d of phi_i at x_q
    for d in range(dim):
        t = np.ones((npc, npc))
        for e in range(dim):
            t *= (grad if e == d else val)[qi[None, :, e], qi[:, None, e]]
        G.append(t)
    wref = np.prod(wq[qi], axis=1)
    # interpolate_boundary_values(Solution) (poisson.cu:155-158)
    bnd = am.boundary_dofs()
    lift = np.zeros(am.n_dofs)
    lift[bnd] = solution(am.support_points()[bnd])[0]
    # assemble_system (poisson.cu:153-229): rhs_i = sum_q (phi_i f - a grad phi_i . grad g~) JxW through the interpolation
    rhs = np.zeros(am.n_dofs)
    for ci in range(am.n_cells):
        row, mask, h = a["loc2glob"][ci].astype(np.int64), int(a["constraint_mask"][ci]), o.h[ci]
        X = a["quadrature_points"][ci]
        _, gu, lu = solution(X)
        coef = 1.0 / (0.05 + 2.0 * (X * X).sum(-1))
        ga = -4.0 * X * (coef ** 2)[:, None]
        f = -coef * lu - (ga * gu).sum(-1)
        gl = resolve_hanging_nodes(lift[row].reshape(shape), mask, p, dim).ravel()
        jxw = h ** dim * wref
        v = Nq.T @ (f * jxw)
        for d in range(dim):
            v += G[d].T @ (-coef * (G[d] @ gl) / h * jxw) / h
        v = resolve_hanging_nodes(v.reshape(shape), mask, p, dim, transpose=True).ravel()
        np.add.at(rhs, row, v)
    rhs[o.constrained] = 0.0
    rhs += lift
    # SolverCG with the inverse diagonal (poisson.cu:233-260) on the oracle operator (identity on constrained rows)
    dinv = o.inverse_diagonal()
    x = np.zeros(am.n_dofs)
    g = -rhs.copy()
    hh = dinv * g
    d = -hh
    gh = g @ hh
    its, tol = 0, 1e-12 * np.linalg.norm(rhs)
    while np.linalg.norm(g) > tol and its < 20000:
        its += 1
        Ad = o.vmult(d)
        alpha = gh / (d @ Ad)
        x += alpha * d
        g += alpha * Ad
        hh = dinv * g
        gh_new = g @ hh
        d = gh_new / gh * d - hh
        gh = gh_new
    err2 = 0.0
    for ci in range(am.n_cells):
        row, mask, h = a["loc2glob"][ci].astype(np.int64), int(a["constraint_mask"][ci]), o.h[ci]
        xl = resolve_hanging_nodes(x[row].reshape(shape), mask, p, dim).ravel()
        e = Nq @ xl - solution(a["quadrature_points"][ci])[0]
        err2 += (e * e * h ** dim * wref).sum()
    return am.n_dofs, its, np.sqrt(err2)


@pytest.mark.parametrize("dim,p,rs", [(2, 2, (3, 4)), (2, 4, (2, 3))])
def test_l2_error_on_locally_refined_meshes_falls_at_the_optimal_rate(dim, p, rs):
    errs = []
    for r in rs:
        _, its, err = solve_poisson(dim, p, r)
        assert its > 0
        errs.append(err)
    ratio = errs[0] / errs[1]
    assert 0.6 * 2 ** (p + 1) <= ratio <= 1.6 * 2 ** (p + 1), (errs, ratio)


def solve_poisson_on_ball(dim, p, r):
    """(n_dofs, L2 error) on the library's BALL_GRID substrate (hyper_ball + SphericalManifold + refine_global(r), MappingQ1): the
    same problem with the non-affine geometry arrays (K, JxW per quadrature point), Dirichlet data on the polygonal boundary,
    sparse direct solve"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    bm = mf.BallMesh(dim, p, r).distribute_dofs()
    a = bm.arrays()
    n, npc = p + 1, (p + 1) ** dim
    val, grad, _, _ = (np.asarray(t) for t in mf.shape_info(p))
    q = np.arange(npc)
    qi = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    Nq = np.ones((npc, npc))
    for e in range(dim):
        Nq *= val[qi[None, :, e], qi[:, None, e]]
    B = np.zeros((dim, npc, npc))
    for d in range(dim):
        t = np.ones((npc, npc))
        for e in range(dim):
            t *= (grad if e == d else val)[qi[None, :, e], qi[:, None, e]]
        B[d] = t
    nd, bnd = bm.n_dofs, a["boundary"].astype(np.int64)
    lift = np.zeros(nd)
    lift[bnd] = solution(bm.support_points()[bnd])[0]
    rows, cols, vals, rhs = [], [], [], np.zeros(nd)
    for c in range(bm.n_cells):
        row, X, coef, jxw = a["loc2glob"][c].astype(np.int64), a["quadrature_points"][c], a["coefficient"][c], a["JxW"][c]
        gx = np.einsum("qed,eqi->qdi", a["inv_jac"][c], B)                 # grad_x phi_i = K^T grad_xi phi_i
        Ac = np.einsum("qdi,q,qdj->ij", gx, coef * jxw, gx)
        rows.append(np.repeat(row, npc)); cols.append(np.tile(row, npc)); vals.append(Ac.ravel())
        _, gu, lu = solution(X)
        f = -coef * lu - ((-4.0 * X * (coef ** 2)[:, None]) * gu).sum(-1)
        rhs[row] += Nq.T @ (f * jxw) - Ac @ lift[row]
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nd, nd))
    free = np.ones(nd, bool)
    free[bnd] = False
    x = lift.copy()
    x[free] = spl.spsolve(A[free][:, free].tocsc(), rhs[free])
    err2 = 0.0
    for c in range(bm.n_cells):
        e = Nq @ x[a["loc2glob"][c].astype(np.int64)] - solution(a["quadrature_points"][c])[0]
        err2 += (e * e * a["JxW"][c]).sum()
    return nd, np.sqrt(err2)


@pytest.mark.parametrize("dim,p,rs", [(2, 2, (2, 3)), (2, 4, (2, 3)), (3, 2, (1, 2))])
def test_l2_error_on_the_ball_mesh_falls_at_the_optimal_rate(dim, p, rs):
    errs = [solve_poisson_on_ball(dim, p, r)[1] for r in rs]
    ratio = errs[0] / errs[1]
    assert 0.6 * 2 ** (p + 1) <= ratio <= 1.6 * 2 ** (p + 1), (errs, ratio)
