"""The oracle pinned against exact mathematics instead of against itself.

deal.II is not available here, so oracle/ is a restatement of the reference's CPU operator (laplace_operator_cpu.cc:125-143).
For a constant coefficient on an affine cell Gauss(p+1) quadrature is exact, so the cell matrix of the restatement must be
K = h^(dim-2) (S x M x M + M x S x M + M x M x S) with the exact 1-D mass / stiffness matrices of the Gauss-Lobatto
Lagrange basis -- tests/golden/exact_cell_matrices.json, computed symbolically by tests/golden/make_exact_cell_matrices.py
(rational for p <= 2, 40-digit algebraic numbers for p = 3, 4).  This anchors shape functions, support points, quadrature,
the metric terms and the lexicographic <-> hierarchic numbering of the oracle (and, through the GPU parity tests, of the
CUDA path) to values that depend on neither deal.II nor this repository."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle.oracle import OracleMesh, shape_1d

GOLD = os.path.join(os.path.dirname(__file__), "golden", "exact_cell_matrices.json")


def exact_1d(p):
    d = json.load(open(GOLD))[str(p)]
    M = np.array([[np.longdouble(v) for v in row] for row in d["mass"]])
    S = np.array([[np.longdouble(v) for v in row] for row in d["stiffness"]])
    return d, M, S


def kron_cell_matrix(dim, M, S, h):
    terms = []
    for d in range(dim):
        mats = [S if e == d else M for e in range(dim)]
        K = mats[0]
        for e in range(1, dim):
            K = np.kron(mats[e], K)  # x fastest: the first factor is the innermost index
        terms.append(K)
    return sum(terms) * np.longdouble(h) ** (dim - 2)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_single_cell_matrix_equals_exact_kronecker_form(dim, p):
    o = OracleMesh(dim, p, 0)  # one cell [-1,1]^dim, h = 2
    o.set_constant_coefficient(1.0)
    o.clear_constraints()
    A = o.assemble_dense()
    l2g = o.loc2glob[0].astype(int)
    K = A[np.ix_(l2g, l2g)]
    _, M, S = exact_1d(p)
    want = kron_cell_matrix(dim, M, S, 2.0).astype(np.float64)
    assert np.abs(K - want).max() <= 2e-14 * np.abs(want).max()
    # and the matrix-free apply of the restatement reproduces the same matrix column by column
    for j in (0, len(l2g) // 2, len(l2g) - 1):
        e = np.zeros(o.n_dofs)
        e[l2g[j]] = 1.0
        col = o.vmult(e)[l2g]
        assert np.abs(col - want[:, j]).max() <= 2e-14 * np.abs(want).max()


def test_exact_fixture_is_what_the_textbooks_say():
    """Q1: M = [[1/3,1/6],[1/6,1/3]], S = [[1,-1],[-1,1]]; Q2: S = 1/3 [[7,-8,1],[-8,16,-8],[1,-8,7]], M = 1/30 [[4,2,-1],[2,16,2],[-1,2,4]]"""
    d1, d2 = json.load(open(GOLD))["1"], json.load(open(GOLD))["2"]
    F = lambda rows: [[Fraction(v) for v in r] for r in rows]
    assert F(d1["mass_exact"]) == [[Fraction(1, 3), Fraction(1, 6)], [Fraction(1, 6), Fraction(1, 3)]]
    assert F(d1["stiffness_exact"]) == [[1, -1], [-1, 1]]
    assert F(d2["stiffness_exact"]) == [[Fraction(7, 3), Fraction(-8, 3), Fraction(1, 3)], [Fraction(-8, 3), Fraction(16, 3), Fraction(-8, 3)],
                                        [Fraction(1, 3), Fraction(-8, 3), Fraction(7, 3)]]
    assert F(d2["mass_exact"]) == [[Fraction(2, 15), Fraction(1, 15), Fraction(-1, 30)], [Fraction(1, 15), Fraction(8, 15), Fraction(1, 15)],
                                   [Fraction(-1, 30), Fraction(1, 15), Fraction(2, 15)]]


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_oracle_support_points_and_quadrature_against_exact_values(p):
    """Gauss-Lobatto nodes from the symbolic fixture; Gauss(p+1) integrates the exact mass matrix through the oracle's own
    shape values (sum_q w_q phi_i(x_q) phi_j(x_q) = M_ij)"""
    d, M, _ = exact_1d(p)
    val, grad, nodes, xq, wq = shape_1d(p)
    assert np.abs(np.asarray(nodes, float) - np.array([float(v) for v in d["nodes"]])).max() <= 1e-15
    val = np.asarray(val).reshape(p + 1, p + 1)  # [i][q]
    Mq = np.einsum("iq,jq,q->ij", val, val, np.asarray(wq))
    assert np.abs(Mq - M.astype(float)).max() <= 1e-15
