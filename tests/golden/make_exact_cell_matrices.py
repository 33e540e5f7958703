"""Exact 1-D mass and stiffness matrices of the Gauss-Lobatto Lagrange basis on [0,1] (FE_Q(p), p = 1..4), computed
symbolically with sympy (rational for p <= 2, algebraic numbers for p = 3, 4) and written as 40-digit decimal strings
(plus exact rationals where they exist) to exact_cell_matrices.json.

Why: the reference's oracle (deal.II MatrixFree, laplace_operator_cpu.cc:125-143) cannot be built here, so oracle/ is a
restatement.  Gauss(p+1) quadrature integrates grad(phi_i).grad(phi_j) on an affine cell exactly, hence for a constant
coefficient the cell matrix of the restatement must equal  K = S x M x M + M x S x M + M x M x S  (Kronecker products,
x fastest) built from these exact 1-D matrices -- an anchor that depends on neither deal.II nor the restatement.

Run:  python tests/golden/make_exact_cell_matrices.py   (needs sympy; the tests only read the JSON)
"""
import json
import os

import sympy as sp

x = sp.symbols("x")


def gll_nodes(p):
    """Gauss-Lobatto points on [0,1]: the end points and the roots of P_p'(2x-1)"""
    if p == 1:
        return [sp.Integer(0), sp.Integer(1)]
    t = sp.symbols("t")
    roots = sp.roots(sp.diff(sp.legendre(p, t), t), t, multiple=True)
    inner = sorted([(1 + sp.simplify(r)) / 2 for r in roots], key=lambda v: sp.N(v, 50))
    return [sp.Integer(0)] + inner + [sp.Integer(1)]


def lagrange(nodes):
    out = []
    for i, xi in enumerate(nodes):
        num, den = sp.Integer(1), sp.Integer(1)
        for j, xj in enumerate(nodes):
            if j != i:
                num *= (x - xj)
                den *= (xi - xj)
        out.append(sp.expand(num / den))
    return out


def main():
    data = {}
    for p in range(1, 5):
        nodes = gll_nodes(p)
        ell = lagrange(nodes)
        n = p + 1
        M = [[sp.nsimplify(sp.simplify(sp.integrate(ell[i] * ell[j], (x, 0, 1)))) for j in range(n)] for i in range(n)]
        S = [[sp.nsimplify(sp.simplify(sp.integrate(sp.diff(ell[i], x) * sp.diff(ell[j], x), (x, 0, 1)))) for j in range(n)] for i in range(n)]
        entry = dict(nodes=[str(sp.N(v, 40)) for v in nodes],
                     mass=[[str(sp.N(v, 40)) for v in row] for row in M],
                     stiffness=[[str(sp.N(v, 40)) for v in row] for row in S])
        if all(v.is_Rational for row in M + S for v in row):
            entry["mass_exact"] = [[str(v) for v in row] for row in M]
            entry["stiffness_exact"] = [[str(v) for v in row] for row in S]
        # sanity: partition of unity
        assert all(sp.simplify(sum(M[i][j] for j in range(n)) - sp.integrate(ell[i], (x, 0, 1))) == 0 for i in range(n))
        assert all(sp.simplify(sum(S[i][j] for j in range(n))) == 0 for i in range(n))
        data[str(p)] = entry
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "exact_cell_matrices.json")
    with open(out, "w") as fh:
        json.dump(data, fh, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
