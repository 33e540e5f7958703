"""Header-only generic path (include/dealii_cuda_b200/fee_gpu.cuh): FEEvaluationGpu + cell_loop with USER-WRITTEN functors
compiled by nvcc (examples/generic_ops.cu -> examples/_build/libgeneric_ops.so, built by __graft_entry__.build()).
Every functor is checked against a numpy restatement of what the reference's FEEvaluationGpu methods compute
(fee_gpu.cuh:197-365): mass operator, the reference's Laplace LocalOperator, a right-hand side integral."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle.oracle import OracleMesh, sm64

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    import dealii_cuda_b200 as mf
    return mf.Context(0, None)


@pytest.fixture(scope="module")
def gen():
    path = os.path.join(os.environ.get("MFG_EXAMPLES_BUILD") or os.path.join(ROOT, "examples", "_build"), "libgeneric_ops.so")
    assert os.path.exists(path), "examples/_build/libgeneric_ops.so is missing: run __graft_entry__.build()"
    lib = C.CDLL(path)
    lib.generic_apply.restype = C.c_int
    lib.generic_apply.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.generic_last_error.restype = C.c_char_p

    def apply(mfree, which, dim, p, dtype, dst, src, coef=None):
        rc = lib.generic_apply(mfree.h, which, dim, p, 1 if np.dtype(dtype) == np.float64 else 0, C.c_void_p(dst.getData()),
                               C.c_void_p(src.getData()), C.c_void_p(coef.getData() if coef is not None else 0))
        assert rc == 0, lib.generic_last_error().decode()
    return apply


def rel_err(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300)


def tables(o, dim, p):
    n = p + 1
    q = np.arange(n ** dim)
    q_idx = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    N = np.asarray(o.shape_values)
    Nq = np.ones((n ** dim, n ** dim))
    for e in range(dim):
        Nq *= N[q_idx[None, :, e], q_idx[:, None, e]]   # [q][i] = prod_e phi_{i_e}(x_{q_e})
    return q_idx, Nq


def uniform_geometry(o, dim, p, left=-1.0, right=1.0):
    import dealii_cuda_b200 as mf
    _, _, xq, wq = mf.shape_info(p)
    q_idx, Nq = tables(o, dim, p)
    h = (right - left) / round(o.n_cells ** (1.0 / dim))
    cc = np.asarray(o.cell_coords)[:, :dim].astype(np.float64)
    X = left + h * (cc[:, None, :] + xq[q_idx][None, :, :])
    JxW = h ** dim * np.prod(wq[q_idx], axis=1)
    return X, JxW, Nq


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("coloring", [False, True])
@pytest.mark.parametrize("dim,p,r", [(2, 1, 3), (2, 2, 2), (2, 4, 2), (3, 1, 2), (3, 2, 1), (3, 3, 1), (3, 4, 1), (3, 4, 2)])
def test_mass_operator_and_rhs(ctx, gen, dim, p, r, coloring, dtype):
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    mesh = mf.HyperCubeMesh(ctx, dim, p, r)
    mfree = mf.MatrixFreeGpu(ctx, dtype)
    mfree.reinit(mesh, use_coloring=coloring)
    X, JxW, Nq = uniform_geometry(o, dim, p)
    l2g = np.asarray(o.loc2glob).astype(np.int64)
    u = sm64(4, o.n_dofs)
    Mloc = Nq.T @ (JxW[:, None] * Nq)
    want = np.zeros(o.n_dofs)
    np.add.at(want, l2g, u[l2g] @ Mloc.T)
    src, dst = mf.GpuVector.from_numpy(ctx, u.astype(dtype)), mf.GpuVector(ctx, o.n_dofs, dtype)
    dst.fill(0.0)
    gen(mfree, 0, dim, p, dtype, dst, src)
    tol = 1e-12 if dtype == np.float64 else 2e-5
    assert rel_err(dst.toVector(), want) <= tol
    # sum of the mass-matrix apply of the constant 1 = volume of the domain
    one = mf.GpuVector(ctx, o.n_dofs, dtype); one.fill(1.0)
    dst.fill(0.0)
    gen(mfree, 0, dim, p, dtype, dst, one)
    assert abs(dst.toVector().astype(np.float64).sum() - 2.0 ** dim) <= (1e-11 if dtype == np.float64 else 1e-4) * 2.0 ** dim
    # right-hand side (phi_i, f), f = 1 + x0 + 2 x1 [+ 3 x2] at the quadrature points
    f = 1.0 + (X * (1 + np.arange(dim))).sum(-1)
    bw = np.zeros(o.n_dofs)
    np.add.at(bw, l2g, (f * JxW[None, :]) @ Nq)
    dst.fill(0.0)
    gen(mfree, 2, dim, p, dtype, dst, src)
    assert rel_err(dst.toVector(), bw) <= tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,p,r", [(2, 2, 3), (2, 4, 2), (3, 1, 2), (3, 2, 2), (3, 3, 1), (3, 4, 1), (3, 4, 2)])
def test_user_written_laplace_matches_oracle(ctx, gen, dim, p, r, dtype):
    """the reference's own LocalOperator (laplace_operator_gpu.h:247-282) written against the generic FEEvaluationGpu:
    must give the unconstrained variable-coefficient Laplace apply of the C oracle"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    coef = np.asarray(o.coefficient).copy()
    o.clear_constraints()
    mesh = mf.HyperCubeMesh(ctx, dim, p, r)
    mfree = mf.MatrixFreeGpu(ctx, dtype)
    mfree.reinit(mesh)
    u = sm64(8, o.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u.astype(dtype)), mf.GpuVector(ctx, o.n_dofs, dtype)
    cdev = mf.GpuVector.from_numpy(ctx, coef.reshape(-1).astype(dtype))
    dst.fill(0.0)
    gen(mfree, 1, dim, p, dtype, dst, src, cdev)
    assert rel_err(dst.toVector(), o.vmult(u)) <= (1e-12 if dtype == np.float64 else 2e-5)


def test_user_written_laplace_general_geometry(ctx, gen):
    """generic path on a deformed mesh (full J^-1 per quadrature point) against the dense numpy restatement"""
    import dealii_cuda_b200 as mf
    import importlib.util
    spec = importlib.util.spec_from_file_location("general_geometry_checker", os.path.join(ROOT, "tests", "test_gpu_general_geometry.py"))
    gg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gg)
    geometry, reference_apply = gg.geometry, gg.reference_apply
    dim, p, r = 3, 2, 1
    o = OracleMesh(dim, p, r)
    K, JxW, xq, q_idx = geometry(o, dim, p, eps=0.08)
    coef = 1.0 / (0.05 + 2.0 * (xq ** 2).sum(-1))
    mfree = mf.MatrixFreeGpu(ctx, np.float64)
    mfree.reinit(dict(dim=dim, degree=p, n_dofs=o.n_dofs, loc2glob=np.asarray(o.loc2glob), inv_jac=K, JxW=JxW))
    u = sm64(12, o.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs)
    cdev = mf.GpuVector.from_numpy(ctx, coef.reshape(-1))
    dst.fill(0.0)
    gen(mfree, 1, dim, p, np.float64, dst, src, cdev)

    class O:
        pass
    ov = O()
    ov.n_cells, ov.n_dofs, ov.constrained, ov.loc2glob = o.n_cells, o.n_dofs, np.zeros(0, np.uint32), np.asarray(o.loc2glob)
    ov.shape_values, ov.shape_gradients = o.shape_values, o.shape_gradients
    assert rel_err(dst.toVector(), reference_apply(ov, dim, p, K, JxW, coef, q_idx, u)) <= 1e-12
