"""Host substrate of the reference's BALL_GRID (csrc/ball_mesh.cu, mfg_umesh_*): hyper_ball + SphericalManifold on the boundary +
refine_global as an unstructured mesh with FE_Q DoFs and MappingQ1 geometry.  deal.II's numbering of this mesh cannot be pinned
here; what is checked (numpy, CPU): every DoF has ONE support point whichever cell evaluates it (i.e. edges and faces shared
by differently oriented cells are matched correctly), the DoF count follows from the entity counts, boundary DoFs lie on
boundary faces, the measure converges to the ball's with h^2, and the operator built from K / JxW is symmetric, positive
semi-definite and annihilates constants."""
import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.oracle import shape_1d


def support_points_by_cell(bm, p):
    """[n_cells][npc][dim]: the tri-linear image of the local support points"""
    dim, n = bm.dim, p + 1
    _, _, xn, _, _ = shape_1d(p)
    V, C = bm.mesh()
    li = np.arange(n ** dim)
    xi = np.stack([xn[(li // n ** d) % n] for d in range(dim)], axis=1)            # [npc][dim]
    N = np.ones((n ** dim, 1 << dim))
    for v in range(1 << dim):
        for d in range(dim):
            N[:, v] *= xi[:, d] if (v >> d) & 1 else 1 - xi[:, d]
    return np.einsum("iv,cvd->cid", N, V[C.astype(np.int64)])


def entity_counts(C, dim):
    edges, faces = set(), set()
    for c in C.tolist():
        for a in range(1 << dim):
            for d in range(dim):
                if not (a >> d) & 1:
                    edges.add(tuple(sorted((c[a], c[a | (1 << d)]))))
        if dim == 3:
            for d in range(3):
                for side in (0, 1):
                    faces.add(tuple(sorted(c[v] for v in range(8) if ((v >> d) & 1) == side)))
    return len(edges), len(faces)


@pytest.mark.parametrize("dim,p,r", [(2, 1, 2), (2, 3, 2), (2, 4, 1), (3, 1, 1), (3, 2, 1), (3, 3, 1), (3, 4, 0), (3, 4, 1)])
def test_dofs_of_shared_entities_are_matched(dim, p, r):
    bm = mf.BallMesh(dim, p, r).distribute_dofs()
    a = bm.arrays()
    X = support_points_by_cell(bm, p)
    l2g = a["loc2glob"].astype(np.int64)
    # one support point per global DoF, whichever cell computes it
    first = np.full((bm.n_dofs, dim), np.nan)
    for c in range(bm.n_cells):
        seen = ~np.isnan(first[l2g[c], 0])
        assert np.abs(first[l2g[c]][seen] - X[c][seen]).max(initial=0.0) <= 1e-13
        first[l2g[c]] = np.where(seen[:, None], first[l2g[c]], X[c])
    assert not np.isnan(first).any()
    # and different DoFs have different support points
    key = np.round(first * 2 ** 30).astype(np.int64)
    assert len({tuple(k) for k in key.tolist()}) == bm.n_dofs
    # count: vertices + (p-1) per edge + (p-1)^2 per face + (p-1)^dim per cell
    V, C = bm.mesh()
    ne, nf = entity_counts(C, dim)
    used_vertices = len(np.unique(C))
    want = used_vertices + ne * (p - 1) + (nf * (p - 1) ** 2 if dim == 3 else 0) + bm.n_cells * (p - 1) ** dim
    assert bm.n_dofs == want
    # first touch: the first cell numbers its DoFs 0 .. npc-1 in hierarchic order
    from oracle.oracle import hier_to_lex
    assert np.array_equal(l2g[0][hier_to_lex(dim, p)], np.arange((p + 1) ** dim))


@pytest.mark.parametrize("dim", [2, 3])
def test_boundary_and_measure(dim):
    p = 2
    vol = []
    for r in (1, 2, 3):
        bm = mf.BallMesh(dim, p, r).distribute_dofs()
        a = bm.arrays()
        assert bm.n_cells == (5 if dim == 2 else 7) * (1 << dim) ** r
        assert (a["JxW"] > 0).all()
        vol.append(a["JxW"].sum())
        # boundary DoFs: exactly those whose support point lies on a boundary face; its vertices are on the sphere
        X = support_points_by_cell(bm, p)
        first = np.zeros((bm.n_dofs, dim))
        first[a["loc2glob"].astype(np.int64)] = X
        V, C = bm.mesh()
        rad = np.linalg.norm(V, axis=1)
        on_sphere = np.abs(rad - 1.0) < 1e-12
        b = np.zeros(bm.n_dofs, bool); b[a["boundary"]] = True
        assert np.linalg.norm(first[b], axis=1).max() <= 1.0 + 1e-12
        assert np.linalg.norm(first[~b], axis=1).max() < 1.0 - 1e-6
        assert on_sphere.sum() > 0 and rad.max() <= 1 + 1e-12
    exact = np.pi if dim == 2 else 4.0 / 3.0 * np.pi
    err = [exact - v for v in vol]
    assert all(e > 0 for e in err)                     # the polygon / polyhedron is inscribed
    assert 3.0 <= err[0] / err[1] <= 5.0 and 3.3 <= err[1] / err[2] <= 4.7, err     # O(h^2)


@pytest.mark.parametrize("dim,p,r", [(2, 2, 1), (2, 4, 1), (3, 1, 1), (3, 2, 0), (3, 3, 0)])
def test_operator_from_ball_geometry(dim, p, r):
    """dense numpy operator from loc2glob, K, JxW, a: symmetric, constants in its kernel, positive on the rest"""
    bm = mf.BallMesh(dim, p, r).distribute_dofs()
    a = bm.arrays()
    n, npc = p + 1, (p + 1) ** dim
    val, grad, _, _ = (np.asarray(t) for t in mf.shape_info(p))
    q = np.arange(npc)
    qi = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    B = np.zeros((dim, npc, npc))                        # [d][q][i]
    for d in range(dim):
        t = np.ones((npc, npc))
        for e in range(dim):
            t *= (grad if e == d else val)[qi[None, :, e], qi[:, None, e]]
        B[d] = t
    A = np.zeros((bm.n_dofs, bm.n_dofs))
    for c in range(bm.n_cells):
        K, w = a["inv_jac"][c], a["coefficient"][c] * a["JxW"][c]
        gx = np.einsum("qed,eqi->qdi", K, B)             # grad_x phi_i at q = K^T grad_xi
        Ac = np.einsum("qdi,q,qdj->ij", gx, w, gx)
        row = a["loc2glob"][c].astype(np.int64)
        A[np.ix_(row, row)] += Ac
    assert np.abs(A - A.T).max() <= 1e-12 * np.abs(A).max()
    assert np.abs(A @ np.ones(bm.n_dofs)).max() <= 1e-11 * np.abs(A).max()
    ev = np.linalg.eigvalsh(A)
    assert ev[0] >= -1e-10 * ev[-1] and ev[1] > 1e-8 * ev[-1]          # one zero eigenvalue: the constants
    # u = x_0 lies in the iso-parametric space (p >= 1): grad u = e_0 exactly, so u^T A u = sum_q a(x_q) JxW_q
    X = support_points_by_cell(bm, p)
    u = np.zeros(bm.n_dofs); u[a["loc2glob"].astype(np.int64)] = X[:, :, 0]
    assert abs(u @ A @ u - (a["coefficient"] * a["JxW"]).sum()) <= 1e-11 * (a["coefficient"] * a["JxW"]).sum()
