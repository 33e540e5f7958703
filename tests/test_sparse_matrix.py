"""The assembled-matrix competitor (CUDAWrappers::SparseMatrix, bmop_spm.cu / test_spm.cu): host assembly in the library
(mfg_csr_assemble_laplace) against the oracle operator on the CPU (the device matrix and its CSR kernel: tests/test_z_late_gpu_additions.py)."""
import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.oracle import OracleMesh, sm64


def csr_matvec(rp, col, val, u):
    import scipy.sparse as sp
    return sp.csr_matrix((val, col.astype(np.int64), rp.astype(np.int64)), shape=(rp.size - 1, rp.size - 1)) @ u


@pytest.mark.parametrize("dim,p,r", [(2, 1, 3), (2, 4, 2), (3, 1, 2), (3, 2, 2), (3, 4, 1), (3, 3, 2)])
def test_assembled_matrix_equals_matrix_free_operator(dim, p, r):
    """test_laplace_op.cu:183-195 as a known-answer test: assembled sparse matrix times u == matrix-free operator"""
    o = OracleMesh(dim, p, r)
    h = 2.0 / (1 << r)
    rp, col, val = mf.assemble_laplace_csr(dim, p, o.loc2glob, o.n_dofs, np.full(o.n_cells, 1.0 / h), o.coefficient, o.constrained)
    assert rp[-1] == val.size and np.all(np.diff(rp.astype(np.int64)) >= 1)
    u = sm64(3, o.n_dofs)
    want = o.vmult(u)
    got = csr_matvec(rp, col, val, u)
    free = np.ones(o.n_dofs, bool); free[o.constrained] = False
    # constrained columns are eliminated: compare on an input that is zero there, and the identity rows separately
    u0 = u.copy(); u0[o.constrained] = 0.0
    assert np.linalg.norm(csr_matvec(rp, col, val, u0)[free] - o.vmult(u0)[free]) <= 1e-13 * np.linalg.norm(want)
    assert np.array_equal(got[o.constrained], u[o.constrained])
    # symmetric, columns sorted inside a row
    import scipy.sparse as sp
    A = sp.csr_matrix((val, col.astype(np.int64), rp.astype(np.int64)), shape=(o.n_dofs, o.n_dofs))
    assert abs(A - A.T).max() <= 1e-13 * abs(A).max()
    assert all(np.all(np.diff(col[rp[i]:rp[i + 1]].astype(np.int64)) > 0) for i in range(0, o.n_dofs, max(1, o.n_dofs // 50)))
