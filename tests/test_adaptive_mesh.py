"""Host substrate for adaptively refined meshes (csrc/adaptive_mesh.cu, mfg_amesh_*): Triangulation with local refinement,
DoFHandler::distribute_dofs, HangingNodes::setup_constraints (matrix_free_gpu/hanging_nodes.cuh:209-454), the reference's
flagging helpers (bmop_common.h:9-105).  Pure host code: checked here on the CPU, bit for bit, against the numpy / Python
oracle (oracle/adaptive.py), which is given the SAME active cells in the same order; the cell SET is checked against the
oracle's own refinement; the hanging-node algebra is checked through the oracle's mask-free constraint matrix."""
import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.adaptive import AdaptiveMesh as OracleAdaptive


def build(dim, p, base, steps):
    """steps: list of ("annulus", R, r, center) | ("shell", R, center) | ("octant",) | ("flags", callable(cells) -> flags)"""
    am = mf.AdaptiveMesh(dim, p)
    am.refine_global(base)
    for s in steps:
        if s[0] == "annulus":
            am.mark_cells_in_annulus(s[1], s[2], s[3])
        elif s[0] == "shell":
            am.mark_cells_on_shell(s[1], s[2])
        elif s[0] == "octant":
            am.mark_octant()
        else:
            am.set_refine_flags(s[1](am.active_cells()))
        am.execute_coarsening_and_refinement()
    return am.distribute_dofs()


def oracle_step(s, dim):
    """the same criterion as a callable(center, half_size) for the oracle's own refinement"""
    if s[0] == "annulus":
        c0 = np.zeros(dim) if s[3] is None else np.asarray(s[3], float)[:dim]
        return lambda c, h: s[2] < np.sqrt(((c - c0) ** 2).sum()) < s[1]
    if s[0] == "shell":
        c0 = np.zeros(dim) if s[2] is None else np.asarray(s[2], float)[:dim]

        def crit(c, h):
            k = sum(np.sqrt(((c + h * (2 * np.array([(v >> d) & 1 for d in range(dim)]) - 1) - c0) ** 2).sum()) < s[1] for v in range(1 << dim))
            return 0 < k < (1 << dim)
        return crit
    if s[0] == "octant":
        return lambda c, h: bool(np.all(c > 0.2))
    raise ValueError


CASES = [
    (2, 1, 2, [("annulus", 0.55, 0.0, None)]),
    (2, 2, 2, [("annulus", 0.5, 0.0, None), ("annulus", 0.42, 0.3, None)]),
    (2, 3, 3, [("shell", 0.6, (-0.1, -0.2)), ("shell", 0.6, (-0.1, -0.2))]),
    (2, 4, 2, [("octant",), ("octant",)]),
    (3, 1, 2, [("annulus", 0.55, 0.0, None)]),
    (3, 2, 1, [("octant",), ("octant",)]),
    (3, 2, 2, [("annulus", 0.6, 0.0, None), ("annulus", 0.45, 0.2, (-0.1, -0.2, -0.3))]),
    (3, 3, 2, [("shell", 0.7, None)]),
    (3, 4, 1, [("octant",)]),
]


@pytest.mark.parametrize("dim,p,base,steps", CASES)
def test_amesh_matches_oracle_bit_for_bit(dim, p, base, steps):
    am = build(dim, p, base, steps)
    cells = am.active_cells()
    a = am.arrays()
    assert a["constraint_mask"].max() > 0, "the case must have hanging nodes"
    # 1. the same active cells as the oracle's own refinement under deal.II's one-level rule (faces, 3D: + edges)
    o_own = OracleAdaptive(dim, p, base, [oracle_step(s, dim) for s in steps], balance="dealii")
    assert set(map(tuple, cells.tolist())) == set(o_own.cells)
    # 2. deal.II's iteration order: level by level, and inside the level of the base mesh along the Morton curve
    assert np.all(np.diff(cells[:, 0].astype(np.int64)) >= 0)
    # 3. every array against the oracle on the same cells in the same order
    o = OracleAdaptive(dim, p, 0, [], cells=cells.tolist())
    assert am.n_dofs == o.n_dofs
    assert np.array_equal(a["loc2glob_unconstrained"], o.l2g_own)
    assert np.array_equal(a["loc2glob"], o.l2g)
    assert np.array_equal(a["constraint_mask"], o.mask)
    assert np.array_equal(a["hanging"], o.hanging)
    assert np.array_equal(a["constrained"], o.constrained)
    assert np.array_equal(a["inv_jac"], o.inv_jac)
    assert np.allclose(a["coefficient"], o.coef, rtol=1e-14, atol=0)


@pytest.mark.parametrize("dim,p,base,steps", [CASES[1], CASES[5]])
def test_amesh_operator_data_reproduce_the_mask_free_assembled_operator(dim, p, base, steps):
    """the arrays of the C++ builder, pushed through the oracle's matrix-free hanging-node apply, equal C^T A C built from
    geometry alone (no masks): the masks / rewritten maps mean what HangingNodes::setup_constraints means"""
    am = build(dim, p, base, steps)
    a = am.arrays()
    o = OracleAdaptive(dim, p, 0, [], cells=am.active_cells().tolist())
    o.l2g, o.mask = a["loc2glob"].copy(), a["constraint_mask"].copy()   # the product's arrays drive the oracle's apply
    o.constrained = a["constrained"].copy()
    o.is_constrained = np.zeros(o.n_dofs, dtype=bool)
    o.is_constrained[o.constrained] = True
    u = np.random.default_rng(3).standard_normal(o.n_dofs)
    want = o.assembled_vmult(u)
    got = o.vmult(u)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)


def test_refine_global_is_the_uniform_mesh():
    """refine_global(r) + distribute_dofs reproduces the uniform substrate (cells along the Morton curve, deal.II numbering)"""
    from oracle.oracle import OracleMesh
    for dim, p, r in [(2, 3, 3), (3, 2, 2), (3, 4, 1)]:
        am = mf.AdaptiveMesh(dim, p).refine_global(r).distribute_dofs()
        o = OracleMesh(dim, p, r)
        a = am.arrays()
        assert am.n_dofs == o.n_dofs and am.n_cells == o.n_cells
        assert np.array_equal(a["loc2glob"], np.asarray(o.loc2glob))
        assert np.array_equal(a["constrained"], np.asarray(o.constrained))
        assert a["constraint_mask"].max() == 0 and a["hanging"].size == 0
        assert np.array_equal(am.active_cells()[:, 1:], np.asarray(o.cell_coords)[:, :dim])
        assert np.allclose(a["coefficient"], np.asarray(o.coefficient), rtol=1e-14, atol=0)


def test_creation_order_of_children():
    """children are appended to their level in the order their parents are visited (deal.II keeps cells per level in
    creation order): refining the LAST cell of a level first puts its children in front of those of cells refined later"""
    am = mf.AdaptiveMesh(2, 1).refine_global(1)
    am.set_refine_flags([0, 0, 0, 1]); am.execute_coarsening_and_refinement()
    am.set_refine_flags([1, 0, 0, 0, 0, 0, 0]); am.execute_coarsening_and_refinement()
    cells = am.active_cells()
    assert cells[:, 0].tolist() == [1, 1] + [2] * 8
    assert cells[2:6, 1:].tolist() == [[2, 2], [3, 2], [2, 3], [3, 3]]      # children of cell (1,1), created first
    assert cells[6:, 1:].tolist() == [[0, 0], [1, 0], [0, 1], [1, 1]]       # children of cell (0,0), created later


def test_one_level_rule_closes_the_flags():
    """a flagged cell next to a coarser cell forces that cell to refine as well (faces; in 3D also edges, not vertices)"""
    am = mf.AdaptiveMesh(2, 1).refine_global(1)
    am.set_refine_flags([1, 0, 0, 0]); am.execute_coarsening_and_refinement()       # level 2 in the lower left quadrant
    flags = np.zeros(am.n_cells, np.uint8)
    cells = am.active_cells()
    target = np.nonzero((cells[:, 0] == 2) & (cells[:, 1] == 1) & (cells[:, 2] == 1))[0][0]   # touches (1,0), (0,1) by faces, (1,1) by a vertex
    flags[target] = 1
    am.set_refine_flags(flags); am.execute_coarsening_and_refinement()
    cells = set(map(tuple, am.active_cells().tolist()))
    assert (3, 2, 2) in cells and (3, 3, 3) in cells                 # the flagged cell was refined
    assert (1, 1, 0) not in cells and (1, 0, 1) not in cells         # its coarser face neighbours had to follow ...
    assert (2, 2, 1) in cells and (2, 1, 2) in cells
    assert (1, 1, 1) in cells                                        # ... the cell it only touches in a vertex did not (2D)
    assert len(cells) == 1 + 3 + 8 + 4


def test_pseudo_adaptive_refinement_counts():
    """bmop_common.h:49-105: sizes of the meshes it produces (regression numbers of this implementation) and basic sanity"""
    am = mf.AdaptiveMesh(2, 2).pseudo_adaptive_refinement(4).distribute_dofs()
    a = am.arrays()
    assert am.n_levels >= 6 and a["constraint_mask"].max() > 0
    # every hanging DoF is constrained, no DoF index is left unused
    used = np.zeros(am.n_dofs, bool); used[a["loc2glob_unconstrained"].ravel()] = True
    assert used.all()
    assert np.isin(a["hanging"], a["constrained"]).all()
    am3 = mf.AdaptiveMesh(3, 1).pseudo_adaptive_refinement(5).distribute_dofs()
    assert am3.n_cells > 512 and am3.arrays()["constraint_mask"].max() > 0


@pytest.mark.parametrize("dim,p,n_ref", [(3, 2, 4), (2, 3, 5)])
def test_cxx_facade_on_the_host_substrates(dim, p, n_ref):
    """include/dealii_cuda_b200/matrix_free_gpu.h (AdaptiveMesh<dim>, BallMesh<dim>) through a host-only example: the same meshes,
    DoF counts and multigrid hierarchy as the Python binding"""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "_build", "host_substrates")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(root, "examples"), "-s", "_build/host_substrates"])
    out = subprocess.run([exe, str(dim), str(p), str(n_ref)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    num = lambda line: {k: int(v) for k, v in re.findall(r"([a-z]+) (\d+)", line)}
    am = mf.AdaptiveMesh(dim, p, limit_level_difference_at_vertices=True).pseudo_adaptive_refinement(n_ref).distribute_dofs().build_mg(0)
    a = am.arrays()
    f = num(lines[0])
    assert (f["cells"], f["levels"], f["dofs"], f["constrained"], f["boundary"]) == (am.n_cells, am.n_levels, am.n_dofs, a["constrained"].size,
                                                                                  am.boundary_dofs().size)
    for l in range(am.n_levels):
        f, lv = num(lines[1 + l]), am.mg_level(l)
        assert (f["cells"], f["dofs"], f["boundary"], f["edge"], f["blocks"], f["copy"]) == (
            lv["loc2glob"].shape[0], lv["n_dofs"], lv["boundary"].size, lv["edge"].size, lv["coarse_idx"].shape[0], lv["copy_global"].size)
    bm = mf.BallMesh(dim, p, max(n_ref - 2, 0)).distribute_dofs()
    f = num(lines[-1])
    assert (f["cells"], f["dofs"], f["boundary"]) == (bm.n_cells, bm.n_dofs, bm.arrays()["boundary"].size)


@pytest.mark.parametrize("dim,p,seed,smooth", [(2, 2, 1, False), (2, 3, 2, True), (2, 4, 6, False), (3, 1, 3, False), (3, 1, 7, True), (3, 2, 4, True)])
def test_random_refinement_matches_oracle(dim, p, seed, smooth):
    """three passes of pseudo-random flags (the reference's RANDOM grid case, poisson_common.h:37-40, with a reproducible
    generator): the same cell set as the oracle's refinement under the same one-level rule, and the same arrays bit for bit"""
    def flagged(center, h):       # a deterministic function of the cell: both sides flag the same cells
        key = np.floor((np.asarray(center) + 1.0) * 4096).astype(np.int64)
        return int((key * np.array([73856093, 19349663, 83492791])[:dim]).sum() * (seed * 2 + 1) + int(h * 65536)) % 100 < 35
    am = mf.AdaptiveMesh(dim, p, limit_level_difference_at_vertices=smooth).refine_global(1)
    for _ in range(3):
        cells = am.active_cells()
        hs = 2.0 / (1 << cells[:, 0].astype(np.int64))
        centers = -1.0 + hs[:, None] * (cells[:, 1:] + 0.5)
        am.set_refine_flags([flagged(c, h / 2) for c, h in zip(centers, hs)])
        am.execute_coarsening_and_refinement()
    am.distribute_dofs()
    o_own = OracleAdaptive(dim, p, 1, [flagged] * 3, balance="vertex" if smooth else "dealii")
    cells = am.active_cells()
    assert set(map(tuple, cells.tolist())) == set(o_own.cells)
    o = OracleAdaptive(dim, p, 0, [], cells=cells.tolist())
    a = am.arrays()
    assert np.array_equal(a["loc2glob_unconstrained"], o.l2g_own) and np.array_equal(a["loc2glob"], o.l2g)
    assert np.array_equal(a["constraint_mask"], o.mask) and np.array_equal(a["constrained"], o.constrained)
    assert np.array_equal(a["hanging"], o.hanging)
    # and the masks / maps mean the right thing: matrix-free apply == C^T A C from geometry alone
    u = np.random.default_rng(seed).standard_normal(o.n_dofs)
    want = o.assembled_vmult(u)
    assert np.linalg.norm(o.vmult(u) - want) <= 1e-12 * np.linalg.norm(want)
