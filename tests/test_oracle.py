"""CPU tests: the oracle against the fixtures of SURVEY.md Appendix A.3 / B, an
independent numpy restatement, and analytic identities.  (parity unpinned at the
deal.II boundary: these fixtures are restatements, not deal.II output.)"""
import json
import os

import numpy as np
import pytest

from oracle.oracle import OracleMesh, hier_to_lex, shape_1d, sm64

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_sm64_known_values():
    u = sm64(1, 4)
    # splitmix64 with seed 1, SURVEY.md Appendix B definition, evaluated independently in Python ints
    def ref(seed, i):
        M = (1 << 64) - 1
        z = (seed + (i + 1) * 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        return (z >> 11) * 2.0 ** -53
    assert [ref(1, i) for i in range(4)] == u.tolist()


def test_dof_map_fixtures():
    # SURVEY.md Appendix A.3 (hyper_cube + refine_global(1), lexicographic loc2glob, Morton cell order)
    m = OracleMesh(2, 1, 1)
    assert m.n_dofs == 9
    assert m.loc2glob.tolist() == [[0, 1, 2, 3], [1, 4, 3, 5], [2, 3, 6, 7], [3, 5, 7, 8]]
    m = OracleMesh(2, 2, 1)
    assert m.n_dofs == 25
    assert m.loc2glob.tolist() == [[0, 6, 1, 4, 8, 5, 2, 7, 3], [1, 12, 9, 5, 14, 11, 3, 13, 10],
                                   [2, 7, 3, 17, 20, 18, 15, 19, 16], [3, 13, 10, 18, 24, 22, 16, 23, 21]]
    assert m.constrained.tolist() == [0, 1, 2, 4, 6, 9, 10, 11, 12, 15, 16, 17, 19, 21, 22, 23]
    m = OracleMesh(3, 1, 1)
    assert m.n_dofs == 27
    assert m.loc2glob.tolist() == [[0, 1, 2, 3, 4, 5, 6, 7], [1, 8, 3, 9, 5, 10, 7, 11], [2, 3, 12, 13, 6, 7, 14, 15],
                                   [3, 9, 13, 16, 7, 11, 15, 17], [4, 5, 6, 7, 18, 19, 20, 21], [5, 10, 7, 11, 19, 22, 21, 23],
                                   [6, 7, 14, 15, 20, 21, 24, 25], [7, 11, 15, 17, 21, 23, 25, 26]]
    assert OracleMesh(3, 4, 1).n_dofs == 729
    assert OracleMesh(3, 4, 2).n_dofs == 4913
    assert OracleMesh(3, 2, 2).n_dofs == 729


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 4), (3, 1), (3, 2), (3, 4), (3, 7)])
def test_hier_to_lex_is_permutation(dim, p):
    h2l = hier_to_lex(dim, p)
    assert sorted(h2l.tolist()) == list(range((p + 1) ** dim))
    n = p + 1
    # vertices first, in lexicographic vertex order
    verts = [sum(((v >> d) & 1) * p * n ** d for d in range(dim)) for v in range(2 ** dim)]
    assert h2l[: 2 ** dim].tolist() == verts


# (dim, p, r, left, right, n_dofs, n_constrained, |A u_sm64|, |A 0.1|, |A^3 0.1|, max|A 0.1|)
APPENDIX_B = [
    (2, 4, 2, 0.0, 1.0, 289, 64, 9.2452053712745e+01, 9.0181032174043e+00, 6.2926436808840e+04, 7.0955757041464e+00),
    (3, 4, 1, -1.0, 1.0, 729, 386, 1.8605806893187e+01, 2.0225656506840e+00, 1.0951331868094e+01, 1.0e-01),
    (3, 4, 2, -1.0, 1.0, 4913, 1538, 3.1828688562091e+01, 3.9421206661754e+00, 3.9246902827962e+00, 1.0e-01),
    (3, 2, 2, -1.0, 1.0, 729, 386, 1.4173263165029e+01, 1.9925922749530e+00, 2.0875001459513e+00, 1.0e-01),
    (2, 3, 3, -1.0, 1.0, 625, 96, 8.3511543612832e+01, 1.2206924282420e+00, 7.6402394121926e+00, 1.0677317000395e-01),
]


@pytest.mark.parametrize("case", APPENDIX_B)
def test_appendix_b_norms(case):
    dim, p, r, lo, hi, nd, nc, n1, n2, n3, mx = case
    m = OracleMesh(dim, p, r, lo, hi)
    assert (m.n_dofs, m.n_constrained) == (nd, nc)
    rel = 1e-11  # fixtures carry 14 significant digits
    assert abs(np.linalg.norm(m.vmult(sm64(1, nd))) - n1) <= rel * n1
    b = m.bmop(1)
    assert abs(np.linalg.norm(b) - n2) <= rel * n2
    assert abs(np.abs(b).max() - mx) <= rel * mx
    assert abs(np.linalg.norm(m.bmop(3)) - n3) <= rel * n3


def numpy_assembled_apply(m, u):
    """Independent restatement (numpy, dense cell matrices, no sum factorisation) of the
    procedure in test_laplace_op.cu:50-120: assemble, identity on constrained rows/cols."""
    dim, p, n = m.dim, m.p, m.p + 1
    from numpy.polynomial.legendre import leggauss
    xg, wg = leggauss(n)
    xq, wq = 0.5 * (xg + 1), 0.5 * wg
    # GLL nodes: roots of (1-x^2) P'_{n-1}
    from numpy.polynomial import legendre as L
    c = np.zeros(n); c[-1] = 1
    xn = np.sort(np.concatenate(([-1.0], L.legroots(L.legder(c)) if n > 2 else [], [1.0])))
    xn = 0.5 * (xn + 1)
    V = np.zeros((n, n)); G = np.zeros((n, n))
    for i in range(n):  # product form of the Lagrange polynomials (monomial coefficients are ill-conditioned)
        others = [j for j in range(n) if j != i]
        den = np.prod([xn[i] - xn[j] for j in others])
        V[i] = np.prod([xq - xn[j] for j in others], axis=0) / den
        G[i] = sum(np.prod([xq - xn[j] for j in others if j != l], axis=0) for l in others) / den
    h = (m.right - m.left) / (1 << m.r)
    # per-cell gradient tables B[d][i][q]
    idx = np.array(np.unravel_index(np.arange(n ** dim), (n,) * dim, order="F")).T  # x fastest
    B = np.zeros((dim, n ** dim, n ** dim)); W = np.ones(n ** dim)
    for d in range(dim):
        t = np.ones((n ** dim, n ** dim))
        for e in range(dim):
            M = G if e == d else V
            t *= M[idx[:, e][:, None], idx[:, e][None, :]]
        B[d] = t / h
    for e in range(dim):
        W *= wq[idx[:, e]] * h
    nd = m.n_dofs
    K = np.zeros((nd, nd))
    coords = m.cell_coords
    for c in range(m.n_cells):
        xq_phys = [m.left + h * (coords[c, e] + xq[idx[:, e]]) for e in range(dim)]
        a = 1.0 / (0.05 + 2.0 * sum(x * x for x in xq_phys))
        Kc = sum((B[d] * (a * W)[None, :]) @ B[d].T for d in range(dim))
        rows = m.loc2glob[c]
        K[np.ix_(rows, rows)] += Kc
    con = m.constrained
    K[con, :] = 0; K[:, con] = 0; K[con, con] = 1.0
    return K @ u


@pytest.mark.parametrize("dim,p,r,lo,hi", [(2, 4, 2, 0.0, 1.0), (3, 1, 2, -1.0, 1.0), (3, 2, 1, -1.0, 1.0),
                                           (3, 3, 1, -1.0, 1.0), (3, 4, 1, -1.0, 1.0), (2, 7, 1, -1.0, 1.0)])
def test_oracle_vs_numpy_assembled(dim, p, r, lo, hi):
    m = OracleMesh(dim, p, r, lo, hi)
    u = sm64(1, m.n_dofs)
    a = m.vmult(u)
    b = numpy_assembled_apply(m, u)
    assert np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(a)


@pytest.mark.parametrize("dim,p,r", [(2, 4, 2), (3, 2, 2), (3, 4, 1)])
def test_oracle_vs_c_dense_assembly(dim, p, r):
    m = OracleMesh(dim, p, r)
    u = sm64(3, m.n_dofs)
    K = m.assemble_dense()
    assert np.abs(K - K.T).max() == 0.0
    a = m.vmult(u)
    assert np.linalg.norm(K @ u - a) <= 1e-13 * np.linalg.norm(a)
    # inverse diagonal == 1/diag(K)
    assert np.allclose(m.inverse_diagonal(), 1.0 / np.diag(K), rtol=1e-13, atol=0)


def test_analytic_identities():
    # a == 1, no constraints, 3D Q4 r=1: A*1 = 0, A*(linear) = 0 at interior rows, u=x -> u^T A u = |Omega| = 8
    m = OracleMesh(3, 4, 1)
    m.set_constant_coefficient(1.0)
    m.clear_constraints()
    assert np.abs(m.vmult(np.ones(m.n_dofs))).max() < 1e-14
    _, _, xn, _, _ = shape_1d(4)
    lat = m.dof_lattice
    x = -1.0 + 1.0 * (lat[:, 0] // 4 + xn[lat[:, 0] % 4])  # h = 1
    x[lat[:, 0] == 8] = 1.0
    Ax = m.vmult(x)
    interior = np.all((lat > 0) & (lat < 8), axis=1)
    assert np.abs(Ax[interior]).max() < 1e-13
    assert abs(x @ Ax - 8.0) < 1e-12
    # symmetry
    u, v = sm64(1, m.n_dofs), sm64(2, m.n_dofs)
    assert abs(v @ m.vmult(u) - u @ m.vmult(v)) < 1e-12 * abs(v @ m.vmult(u))


def test_constrained_rows_identity_and_vmult_add():
    m = OracleMesh(3, 2, 2)
    u = sm64(5, m.n_dofs)
    a = m.vmult(u)
    con = m.constrained
    assert np.array_equal(a[con], u[con])
    d0 = sm64(6, m.n_dofs)
    b = m.vmult_add(d0, u)
    assert np.allclose(b, d0 + a, rtol=1e-14, atol=1e-14)
    assert np.array_equal(b[con], d0[con] + u[con])


def test_threaded_baseline_matches_scalar():
    m = OracleMesh(3, 4, 2)
    u = sm64(1, m.n_dofs)
    a, b = m.vmult(u), m.vmult(u, threaded=True)
    assert np.linalg.norm(a - b) <= 1e-14 * np.linalg.norm(a)


@pytest.mark.parametrize("dim,p,r", [(3, 4, 2), (2, 5, 2), (3, 1, 3), (3, 8, 1), (2, 2, 4)])
def test_simd_cpu_baseline_matches_scalar(dim, p, r):
    """the timed CPU baseline (collocation form, 8-cell batches) is the same operator as the scalar restatement"""
    m = OracleMesh(dim, p, r)
    u = sm64(2, m.n_dofs)
    a, b = m.vmult(u), m.vmult(u, fast=True)
    assert np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(a)
    assert np.array_equal(a[m.constrained], b[m.constrained])


def test_golden_vectors_match_oracle():
    """tests/golden/*.npz were generated by tests/golden/make_golden.py from this oracle;
    they pin the oracle against accidental change and are what the -m gpu tests compare to."""
    path = os.path.join(GOLD, "apply_cases.json")
    cases = json.load(open(path))
    for c in cases:
        m = OracleMesh(c["dim"], c["p"], c["r"], c["left"], c["right"])
        g = np.load(os.path.join(GOLD, c["file"]))
        assert np.array_equal(m.loc2glob, g["loc2glob"])
        assert np.array_equal(m.constrained, g["constrained"])
        a = m.vmult(sm64(c["seed"], m.n_dofs))
        assert np.linalg.norm(a - g["Au"]) <= 1e-14 * np.linalg.norm(a)
