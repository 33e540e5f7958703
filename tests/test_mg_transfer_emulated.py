"""The DEVICE code of the multigrid transfer (csrc/mg_transfer.cu: mg_kernel with the block weight table added for adaptive
hierarchies) run on the CPU emulation of tests/emu/cuda_emu.h, on the blocks the library's host hierarchy produces
(mfg_amesh_build_mg), against the GEOMETRIC prolongation matrix of oracle/adaptive_mg.py -- kernel and data together, before their
first run on hardware.  Also the globally refined form (closed-form valence weights) against the same kind of matrix."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import dealii_cuda_b200 as mf
from oracle.adaptive import AdaptiveMesh as OracleAdaptive
from oracle.adaptive_mg import AdaptiveMultigridOracle
from oracle.oracle import shape_1d, sm64

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    build = tmp_path_factory.mktemp("emu_mg")
    src = open(os.path.join(ROOT, "dealii_cuda_b200", "csrc", "mg_transfer.cu")).read()
    a = src.index("struct PMat { double P[17 * 9]; };")
    b = src.index("__global__ void mark_coarse")
    dev = src[a:b]
    # fine_lattice_points is a setup kernel that is not needed here; keep pass / weights / mg_kernel
    c, d = dev.index("__global__ void fine_lattice_points"), dev.index("// 1-D pass along direction d")
    dev = dev[:c] + dev[d:]
    (build / "mg_kernel_device_part.h").write_text(dev)
    so = build / "libemu_mg.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-shared", "-fPIC", "-pthread", "-I", str(build), "-I", os.path.join(ROOT, "tests", "emu"),
                           "-o", str(so), os.path.join(ROOT, "tests", "emu", "emu_mg_transfer.cc")])
    lib = C.CDLL(str(so))
    u32, dp = C.POINTER(C.c_uint32), C.POINTER(C.c_double)
    lib.emu_mg_transfer.restype = C.c_int
    lib.emu_mg_transfer.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, u32, u32, dp, u32, u32, dp, dp, dp, C.c_uint32]

    def transfer(prolong, dim, p, coarse_idx, fine_idx, weights, src, n_dst, dst0=None, cell_xyz=None, nc=None):
        n, nf = p + 1, 2 * p + 1
        _, _, xn, _, _ = shape_1d(p)
        P1 = np.zeros((nf, n))
        for f in range(nf):
            x = xn[f] / 2 if f <= p else 0.5 + xn[f - p] / 2
            for i in range(n):
                P1[f, i] = np.prod([(x - xn[m]) / (xn[i] - xn[m]) for m in range(n) if m != i])
        ci, fi = np.ascontiguousarray(coarse_idx, np.uint32), np.ascontiguousarray(fine_idx, np.uint32)
        w = np.ascontiguousarray(weights, np.float64) if weights is not None else None
        xyz = np.ascontiguousarray(cell_xyz if cell_xyz is not None else np.zeros((ci.shape[0], 3)), np.uint32)
        ncv = np.ascontiguousarray(nc if nc is not None else [1, 1, 1], np.uint32)
        s = np.ascontiguousarray(src, np.float64)
        dst = np.zeros(n_dst) if dst0 is None else np.ascontiguousarray(dst0, np.float64).copy()
        P = lambda x, t: x.ctypes.data_as(t) if x is not None else None
        rc = lib.emu_mg_transfer(int(prolong), dim, p, ci.shape[0], P(ci, u32), P(fi, u32), P(w, dp), P(xyz, u32), P(ncv, u32), P(np.ascontiguousarray(P1), dp),
                                 P(s, dp), P(dst, dp), n_dst)
        assert rc == 0
        return dst
    return transfer


@pytest.mark.parametrize("dim,p,base,steps", [(2, 2, 2, [(0.6, 0.0, None), (0.4, 0.1, (-0.1, -0.2))]), (2, 3, 1, [(0.9, 0.0, None), (0.5, 0.0, None)]),
                                               (3, 1, 1, [(0.9, 0.0, None), (0.5, 0.0, (-0.1, -0.2, -0.3))]), (3, 2, 1, [(0.9, 0.0, None)])])
def test_emulated_transfer_kernel_on_adaptive_blocks(emu, dim, p, base, steps):
    am = mf.AdaptiveMesh(dim, p, limit_level_difference_at_vertices=True).refine_global(base)
    for R, r, c in steps:
        am.mark_cells_in_annulus(R, r, c)
        am.execute_coarsening_and_refinement()
    am.distribute_dofs().build_mg(0)
    o = OracleAdaptive(dim, p, 0, [], cells=am.active_cells().tolist())
    lc = {l: [tuple(int(v) for v in row) for row in am.level_cells(l)] for l in range(am.n_levels)}
    mg = AdaptiveMultigridOracle(dim, p, lc, o, smoother_degree=0, n_eig=1)
    for l in range(1, am.n_levels):
        lv, Pm = am.mg_level(l), mg.P[l]
        nc_, nf_ = mg.levels[l - 1].n_dofs, mg.levels[l].n_dofs
        uc, rf, d0 = sm64(10 + l, nc_), sm64(20 + l, nf_), sm64(30 + l, nc_)
        fine = emu(1, dim, p, lv["coarse_idx"], lv["fine_idx"], lv["weights"], uc, nf_)
        assert np.linalg.norm(fine - Pm @ uc) <= 1e-13 * np.linalg.norm(Pm @ uc)
        coarse = emu(0, dim, p, lv["coarse_idx"], lv["fine_idx"], lv["weights"], rf, nc_, dst0=d0)
        want = d0 + Pm.T @ rf
        assert np.linalg.norm(coarse - want) <= 1e-13 * np.linalg.norm(want)
        assert np.array_equal(coarse[mg.levels[l - 1].boundary], d0[mg.levels[l - 1].boundary])     # boundary rows untouched
