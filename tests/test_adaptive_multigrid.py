"""Multigrid hierarchy on adaptively refined meshes (local smoothing; poisson_mg.cu / bmop_mg.cu with an adaptive grid).
CPU part: the library's host substrate (csrc/adaptive_mesh.cu: mfg_amesh_build_mg) against the numpy oracle
(oracle/adaptive_mg.py) -- level DoF maps, MGConstrainedDoFs sets and copy indices bit for bit, the transfer blocks + weights
(what the transfer kernel consumes) against the GEOMETRIC prolongation matrix; and the oracle itself: its V-cycle is a symmetric
positive definite preconditioner with bounded CG iteration counts."""

import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.adaptive import AdaptiveMesh as OracleAdaptive
from oracle.adaptive_mg import AdaptiveMultigridOracle, cg_preconditioned
from oracle.oracle import shape_1d, sm64

CASES = [
    (2, 1, 2, [(0.6, 0.0, None), (0.4, 0.1, (-0.1, -0.2))]),
    (2, 2, 2, [(0.6, 0.0, None), (0.4, 0.1, (-0.1, -0.2))]),
    (2, 3, 1, [(0.9, 0.0, None), (0.5, 0.0, None), (0.3, 0.0, (-0.1, -0.2))]),
    (3, 1, 1, [(0.9, 0.0, None), (0.5, 0.0, (-0.1, -0.2, -0.3))]),
    (3, 2, 1, [(0.9, 0.0, None), (0.5, 0.0, (-0.1, -0.2, -0.3))]),
]


def build(dim, p, base, steps, min_level=0):
    am = mf.AdaptiveMesh(dim, p, limit_level_difference_at_vertices=True).refine_global(base)
    for R, r, c in steps:
        am.mark_cells_in_annulus(R, r, c)
        am.execute_coarsening_and_refinement()
    am.distribute_dofs().build_mg(min_level)
    o = OracleAdaptive(dim, p, 0, [], cells=am.active_cells().tolist())
    lc = {l: [tuple(int(v) for v in row) for row in am.level_cells(l)] for l in range(am.n_levels)}
    return am, o, AdaptiveMultigridOracle(dim, p, lc, o, min_level=min_level)


def block_prolongation(dim, p, lv, n_coarse):
    """the matrix the transfer kernel applies from the library's blocks: per refined coarse cell gather (boundary DoFs read as 0),
    tensor-product interpolation to the (2p+1)^dim lattice, weights, scatter-add"""
    n, nf = p + 1, 2 * p + 1
    _, _, xn, _, _ = shape_1d(p)
    P1 = np.zeros((nf, n))
    for f in range(nf):
        x = xn[f] / 2 if f <= p else 0.5 + xn[f - p] / 2
        for i in range(n):
            P1[f, i] = np.prod([(x - xn[m]) / (xn[i] - xn[m]) for m in range(n) if m != i])
    Pd = P1
    for _ in range(dim - 1):
        Pd = np.kron(P1, Pd)                  # lexicographic, x fastest
    P = np.zeros((lv["n_dofs"], n_coarse))
    for q in range(lv["coarse_idx"].shape[0]):
        w = np.zeros(nf ** dim)
        for f in range(nf ** dim):
            a = [(f // nf ** d) % nf for d in range(dim)]
            r = sum((0 if a[d] == 0 else 2 if a[d] == 2 * p else 1) * 3 ** d for d in range(dim))
            w[f] = lv["weights"][q, r]
        ci = lv["coarse_idx"][q]
        keep = (ci & 0x80000000) == 0
        blk = (w[:, None] * Pd)[:, keep]
        np.add.at(P, (lv["fine_idx"][q][:, None], (ci[keep] & 0x7fffffff)[None, :]), blk)
    return P


@pytest.mark.parametrize("dim,p,base,steps", CASES)
def test_hierarchy_matches_oracle(dim, p, base, steps):
    am, o, mg = build(dim, p, base, steps)
    assert o.mask.max() > 0
    for l in range(am.n_levels):
        lv, lm = am.mg_level(l), mg.levels[l]
        assert lv["n_dofs"] == lm.n_dofs
        assert np.array_equal(lv["loc2glob"], lm.l2g)
        assert np.array_equal(lv["boundary"], lm.boundary)
        assert np.array_equal(lv["edge"], lm.edge)
        assert np.allclose(lv["coefficient"], lm.am.coef, rtol=1e-14, atol=0)
        g, lvl = mg.copy[l]
        assert np.array_equal(lv["copy_global"], g) and np.array_equal(lv["copy_level"], lvl)
        if l > 0:
            Pb = block_prolongation(dim, p, lv, mg.levels[l - 1].n_dofs)
            assert np.abs(Pb - mg.P[l]).max() <= 1e-13
    # some level has a refinement edge, and every non-hanging active DoF is reached by exactly one copy pair
    assert any(am.mg_level(l)["edge"].size for l in range(am.n_levels))
    hits = np.zeros(o.n_dofs, dtype=int)
    for l in range(am.n_levels):
        hits[am.mg_level(l)["copy_global"]] += 1
    hanging = np.zeros(o.n_dofs, dtype=bool); hanging[o.hanging] = True
    assert np.all(hits[hanging] == 0) and np.all(hits[~hanging] >= 1)


@pytest.mark.parametrize("dim,p,base,steps", CASES[:2] + CASES[3:4])
def test_oracle_vcycle_is_a_symmetric_positive_definite_preconditioner(dim, p, base, steps):
    am, o, mg = build(dim, p, base, steps)
    M = mg.matrix()
    free = np.setdiff1d(np.arange(o.n_dofs), o.constrained)
    Mf = M[np.ix_(free, free)]
    assert np.abs(Mf - Mf.T).max() <= 1e-12 * np.abs(Mf).max()
    assert np.linalg.eigvalsh(0.5 * (Mf + Mf.T)).min() > 0
    assert np.abs(M[o.hanging]).max() == 0.0          # copy_from_mg leaves hanging DoFs at zero


@pytest.mark.parametrize("dim,p,base,steps", CASES)
def test_oracle_mg_cg_iteration_counts(dim, p, base, steps):
    """the property multigrid is for: few iterations, far fewer than Jacobi-preconditioned CG, and the right solution"""
    am, o, mg = build(dim, p, base, steps)
    ue = sm64(5, o.n_dofs); ue[o.constrained] = 0.0
    b = o.vmult(ue)
    tol = 1e-10 * np.linalg.norm(b)
    x, it, hist = cg_preconditioned(o.vmult, mg.vmult, b, tol)
    assert it <= 12, it
    assert np.linalg.norm(x - ue) <= 1e-8 * np.linalg.norm(ue)
    dinv = o.inverse_diagonal()
    _, it_jacobi, _ = cg_preconditioned(o.vmult, lambda g: dinv * g, b, tol, 5000)
    assert it_jacobi >= 2 * it


def test_min_level_above_zero_and_its_limit():
    am, o, mg = build(2, 2, 2, [(0.6, 0.0, None)], min_level=2)
    lv = am.mg_level(2)
    assert lv["coarse_idx"].shape[0] == 0 and lv["edge"].size == 0      # the coarsest level of the hierarchy covers the domain
    ue = sm64(5, o.n_dofs); ue[o.constrained] = 0.0
    b = o.vmult(ue)
    x, it, _ = cg_preconditioned(o.vmult, mg.vmult, b, 1e-10 * np.linalg.norm(b))
    assert it <= 12
    with pytest.raises(mf.MfgError):
        am.build_mg(3)       # level-2 cells are active: the hierarchy cannot start above them


@pytest.mark.parametrize("dim,p,seed", [(2, 2, 1), (3, 1, 2)])
def test_hierarchy_of_randomly_refined_meshes(dim, p, seed):
    """the same comparison on meshes refined by pseudo-random flags (vertex-balanced): level maps, index sets, copy pairs, transfer
    blocks -- and the V-cycle on them is symmetric"""
    def flagged(center, h):
        key = np.floor((np.asarray(center) + 1.0) * 4096).astype(np.int64)
        return int((key * np.array([73856093, 19349663, 83492791])[:dim]).sum() * (seed * 2 + 1) + int(h * 65536)) % 100 < 30
    am = mf.AdaptiveMesh(dim, p, limit_level_difference_at_vertices=True).refine_global(1)
    for _ in range(3):
        cells = am.active_cells()
        hs = 2.0 / (1 << cells[:, 0].astype(np.int64))
        centers = -1.0 + hs[:, None] * (cells[:, 1:] + 0.5)
        am.set_refine_flags([flagged(c, h / 2) for c, h in zip(centers, hs)])
        am.execute_coarsening_and_refinement()
    am.distribute_dofs().build_mg(0)
    o = OracleAdaptive(dim, p, 0, [], cells=am.active_cells().tolist())
    lc = {l: [tuple(int(v) for v in row) for row in am.level_cells(l)] for l in range(am.n_levels)}
    mg = AdaptiveMultigridOracle(dim, p, lc, o)
    for l in range(am.n_levels):
        lv, lm = am.mg_level(l), mg.levels[l]
        assert np.array_equal(lv["loc2glob"], lm.l2g) and np.array_equal(lv["boundary"], lm.boundary) and np.array_equal(lv["edge"], lm.edge)
        g, lvl = mg.copy[l]
        assert np.array_equal(lv["copy_global"], g) and np.array_equal(lv["copy_level"], lvl)
        if l > 0:
            assert np.abs(block_prolongation(dim, p, lv, mg.levels[l - 1].n_dofs) - mg.P[l]).max() <= 1e-13
    M = mg.matrix()
    free = np.setdiff1d(np.arange(o.n_dofs), o.constrained)
    Mf = M[np.ix_(free, free)]
    assert np.abs(Mf - Mf.T).max() <= 1e-12 * np.abs(Mf).max() and np.linalg.eigvalsh(0.5 * (Mf + Mf.T)).min() > 0
