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


def test_csr_kernel_emulated():
    """the device code of the CSR kernel (one warp per row, __shfl_down_sync reduction) on the CPU emulation of tests/emu: the
    assembled matrix times u equals the oracle operator -- kernel and host assembly together, before their first run on hardware"""
    import ctypes as C
    import os
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "dealii_cuda_b200", "csrc", "sparse_matrix.cu")).read()
    a, b = src.index("template <typename T>\n__global__ void csr_vmult_warp_per_row"), src.index("void assemble(int dim")
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "csr_kernel_device_part.h"), "w").write(src[a:b])
        so = os.path.join(tmp, "libemu_csr.so")
        subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-I", tmp, "-I", os.path.join(root, "tests", "emu"), "-o", so,
                               os.path.join(root, "tests", "emu", "emu_csr.cc")])
        lib = C.CDLL(so)
        u32, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
        lib.emu_csr_vmult.argtypes = [C.c_uint32, u32, u32, dp, dp, dp]
        for dim, p, r in [(2, 3, 2), (3, 2, 1)]:
            o = OracleMesh(dim, p, r)
            rp, col, val = mf.assemble_laplace_csr(dim, p, o.loc2glob, o.n_dofs, np.full(o.n_cells, (1 << r) / 2.0), o.coefficient, o.constrained)
            u = sm64(3, o.n_dofs); u[o.constrained] = 0.0
            y = np.zeros(o.n_dofs)
            lib.emu_csr_vmult(o.n_dofs, rp.ctypes.data_as(u32), col.ctypes.data_as(u32), val.ctypes.data_as(dp), u.ctypes.data_as(dp), y.ctypes.data_as(dp))
            want = o.vmult(u)
            assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)
