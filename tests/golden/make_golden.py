"""Generates the golden fixtures in tests/golden/ from the CPU oracle (oracle/mf_oracle.c).

The oracle itself is pinned against SURVEY.md Appendix A.3/B and an independent
numpy assembled operator (tests/test_oracle.py).  The reference (deal.II based)
cannot be built or imported in this image, so these are oracle outputs, not
outputs of the reference binary.   Usage: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import OracleMesh, sm64  # noqa: E402

CASES = [  # dim, p, r, left, right, seed
    (2, 4, 2, 0.0, 1.0, 1),   # the test_laplace_op.cu configuration (2D Q4, 16 cells, 289 DoFs)
    (3, 4, 1, -1.0, 1.0, 1),
    (3, 4, 2, -1.0, 1.0, 1),
    (3, 2, 2, -1.0, 1.0, 1),
    (2, 3, 3, -1.0, 1.0, 1),
    (3, 1, 2, -1.0, 1.0, 2),
    (3, 3, 1, -1.0, 1.0, 2),
]

out = []
for dim, p, r, lo, hi, seed in CASES:
    m = OracleMesh(dim, p, r, lo, hi)
    u = sm64(seed, m.n_dofs)
    name = "apply_d%d_p%d_r%d.npz" % (dim, p, r)
    np.savez_compressed(os.path.join(HERE, name), loc2glob=m.loc2glob, constrained=m.constrained, Au=m.vmult(u),
                        bmop3=m.bmop(3), inv_diag=m.inverse_diagonal())
    out.append(dict(dim=dim, p=p, r=r, left=lo, right=hi, seed=seed, file=name, n_dofs=int(m.n_dofs)))
json.dump(out, open(os.path.join(HERE, "apply_cases.json"), "w"), indent=1)
print("wrote", len(out), "cases")
