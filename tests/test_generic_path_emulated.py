"""The DEVICE code of the header-only generic path (include/dealii_cuda_b200/fee_gpu.cuh: FEEvaluationGpu read_dof_values with the
hanging-node interpolation, evaluate, quadrature-point operation, integrate, distribute_local_to_global with the transposed
interpolation and atomic adds, apply_kernel_shmem) run on the CPU: tests/emu/cuda_emu.h supplies threadIdx / __syncthreads /
atomicAdd / shared memory (CUDA threads as fibers with real barriers), tests/emu/emu_generic.cc the functors of examples/generic_ops.cu and
the two launches of cell_loop.  Compared with the numpy statements the GPU tests use -- so the kernel logic (not just its index
arithmetic) has run before its first run on hardware."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.adaptive import AdaptiveMesh, resolve_hanging_nodes
from oracle.oracle import sm64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    build = tmp_path_factory.mktemp("emu")
    src = open(os.path.join(ROOT, "include", "dealii_cuda_b200", "fee_gpu.cuh")).read()
    cut = src.index("inline void fee_check(int rc")                   # everything behind is host launch code (kernel<<<...>>>)
    dev = src[:cut] + "\n}  // namespace dealii_cuda_b200\n"
    for inc in ('#include <cuda_runtime.h>\n', '#include "../mfgpu.h"\n', '#include "matrix_free_gpu.h"\n'):
        assert inc in dev
        dev = dev.replace(inc, "")
    (build / "fee_gpu_device_part.h").write_text(dev)
    so = build / "libemu_generic.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-I", str(build), "-I", os.path.join(ROOT, "tests", "emu"),
                           "-o", str(so), os.path.join(ROOT, "tests", "emu", "emu_generic.cc")])
    lib = C.CDLL(str(so))
    u32, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    lib.emu_generic_apply.restype = C.c_int
    lib.emu_generic_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, u32, dp, dp, u32, dp, dp, dp, dp, dp, dp]

    def apply(which, dim, p, a, u, coefficient=None):
        """a: arrays of mf.AdaptiveMesh.arrays(); kernel cell order = cells without a mask first (stable), as mfg_mf_reinit sorts them"""
        mask = a["constraint_mask"]
        perm = np.concatenate([np.nonzero(mask == 0)[0], np.nonzero(mask != 0)[0]])
        n = p + 1
        val, grad, _, wq = (np.asarray(t, dtype=np.float64) for t in mf.shape_info(p))
        q = np.arange(n ** dim)
        wref = np.prod(wq[np.stack([(q // n ** e) % n for e in range(dim)], axis=1)], axis=1)
        l2g = np.ascontiguousarray(a["loc2glob"][perm], dtype=np.uint32)
        inv_jac = np.ascontiguousarray(a["inv_jac"][perm])
        jxw = np.ascontiguousarray((1.0 / inv_jac)[:, None] ** dim * wref[None, :])
        m = np.ascontiguousarray(mask[perm], dtype=np.uint32)
        hang = np.ascontiguousarray(mf.hanging_node_weights(p))
        coef = np.ascontiguousarray(coefficient[perm]) if coefficient is not None else np.zeros(1)
        src = np.ascontiguousarray(u, dtype=np.float64)
        dst = np.zeros_like(src)
        P = lambda x, t: x.ctypes.data_as(t)
        rc = lib.emu_generic_apply(which, dim, p, l2g.shape[0], int((mask == 0).sum()), P(l2g, u32), P(jxw, dp), P(inv_jac, dp), P(m, u32),
                                   P(np.ascontiguousarray(val), dp), P(np.ascontiguousarray(grad), dp), P(hang, dp), P(coef, dp), P(src, dp), P(dst, dp))
        assert rc == 0
        return dst
    return apply


def adaptive_case(dim, p):
    am = mf.AdaptiveMesh(dim, p).refine_global(1)
    am.mark_cells_in_annulus(0.9, 0.0, None); am.execute_coarsening_and_refinement()
    am.mark_cells_in_annulus(0.5, 0.0, (-0.1, -0.2, -0.3)); am.execute_coarsening_and_refinement()
    am.distribute_dofs()
    return am, am.arrays(), AdaptiveMesh(dim, p, 0, [], cells=am.active_cells().tolist())


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2)])
def test_emulated_laplace_functor_with_hanging_nodes(emu, dim, p):
    """the reference's Laplace LocalOperator on the generic path, hanging-node mesh: the free rows of the oracle's operator"""
    am, a, o = adaptive_case(dim, p)
    assert a["constraint_mask"].max() > 0
    u = sm64(31, am.n_dofs); u[o.constrained] = 0.0
    got = emu(1, dim, p, a, u, a["coefficient"])
    free = np.ones(am.n_dofs, bool); free[o.constrained] = False
    want = o.vmult(u)
    assert np.linalg.norm(got[free] - want[free]) <= 1e-12 * np.linalg.norm(want[free])


@pytest.mark.parametrize("dim,p", [(2, 2), (3, 1), (3, 2)])
def test_emulated_mass_functor_with_hanging_nodes(emu, dim, p):
    """a user-written mass operator: gather through the rewritten map, interpolate, local mass matrix, transposed interpolation"""
    am, a, o = adaptive_case(dim, p)
    n = p + 1
    val, _, _, wq = (np.asarray(t) for t in mf.shape_info(p))
    q = np.arange(n ** dim)
    qi = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    Nq = np.ones((n ** dim, n ** dim))
    for e in range(dim):
        Nq *= val[qi[None, :, e], qi[:, None, e]]
    wref = np.prod(wq[qi], axis=1)
    u = sm64(32, am.n_dofs)
    want = np.zeros(am.n_dofs)
    for ci in range(am.n_cells):
        row, mask = a["loc2glob"][ci].astype(np.int64), int(a["constraint_mask"][ci])
        ul = resolve_hanging_nodes(u[row].reshape((n,) * dim), mask, p, dim).ravel()
        v = Nq.T @ ((o.h[ci] ** dim * wref) * (Nq @ ul))
        np.add.at(want, row, resolve_hanging_nodes(v.reshape((n,) * dim), mask, p, dim, transpose=True).ravel())
    got = emu(0, dim, p, a, u)
    assert np.linalg.norm(got - want) <= 1e-12 * np.linalg.norm(want)
