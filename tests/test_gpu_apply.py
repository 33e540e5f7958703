"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(ctypes binding of include/mfgpu.h), against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): DoF maps / constraint lists bit-exact; operator
output within 1e-12 relative (FP64) or 1e-5 (FP32)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.oracle import OracleMesh, sm64  # noqa: E402  (checker only)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    s = max(np.abs(b).max(), 1e-300) if b.size else 1.0  # scale first: the raw bmop loop reaches 1e200
    return np.linalg.norm(a / s - b / s) / max(np.linalg.norm(b / s), 1e-300)


def max_rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


MESH_CASES = [(2, 1, 1), (2, 2, 1), (2, 4, 2), (2, 3, 3), (2, 8, 2), (2, 5, 0), (3, 1, 1), (3, 1, 3), (3, 2, 2), (3, 3, 2),
              (3, 4, 1), (3, 4, 2), (3, 4, 3), (3, 5, 1), (3, 6, 1), (3, 7, 1), (3, 8, 1), (3, 4, 0)]


@pytest.mark.parametrize("dim,p,r", MESH_CASES)
def test_dof_map_and_constraints_bit_exact(ctx, dim, p, r):
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    m = mf.HyperCubeMesh(ctx, dim, p, r)
    assert (m.n_cells, m.n_dofs, m.n_constrained) == (o.n_cells, o.n_dofs, o.n_constrained)
    assert np.array_equal(m.loc2glob(), o.loc2glob)
    assert np.array_equal(m.constrained_dofs(), o.constrained)
    assert np.array_equal(m.cell_coords(), o.cell_coords)
    # lattice lookup agrees with the oracle's DoF -> lattice table
    lat = o.dof_lattice
    assert np.array_equal(m.lattice_to_dof(lat), np.arange(o.n_dofs, dtype=np.uint32))


def test_shape_info_matches_oracle():
    import dealii_cuda_b200 as mf
    from oracle.oracle import shape_1d
    for p in range(1, 9):
        val, grad, xq, wq = mf.shape_info(p)
        oval, ograd, _, oxq, owq = shape_1d(p)
        assert np.allclose(val, oval, rtol=0, atol=1e-15)
        assert np.allclose(grad, ograd, rtol=0, atol=2e-13)
        assert np.allclose(xq, oxq, rtol=0, atol=1e-16) and np.allclose(wq, owq, rtol=0, atol=1e-16)


APPLY_CASES = [(2, p, 3) for p in range(1, 9)] + [(3, p, 2) for p in range(1, 9)] + [(3, 4, 3), (3, 4, 0), (2, 4, 0), (3, 1, 4), (2, 1, 6)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("coloring", [False, True])
@pytest.mark.parametrize("dim,p,r", APPLY_CASES)
def test_vmult_matches_oracle(ctx, dim, p, r, coloring, dtype):
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    m = mf.HyperCubeMesh(ctx, dim, p, r)
    op = mf.LaplaceOperatorGpu(ctx, dtype, use_coloring=coloring)
    op.reinit(m)
    assert op.m() == o.n_dofs
    u = sm64(1, o.n_dofs)
    src = mf.GpuVector.from_numpy(ctx, u.astype(dtype))
    dst = mf.GpuVector(ctx, o.n_dofs, dtype)
    dst.fill(123.0)  # vmult must overwrite
    op.vmult(dst, src)
    got = dst.toVector()
    want = o.vmult(u.astype(dtype).astype(np.float64))
    assert rel_err(got, want) <= TOL[dtype]
    assert max_rel_err(got, want) <= 4 * TOL[dtype]
    # constrained rows are the identity, bit-exact; src is left untouched bit-exactly
    con = o.constrained
    assert np.array_equal(got[con], u.astype(dtype)[con])
    assert np.array_equal(src.toVector(), u.astype(dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("variant", [1, 40, 51, 52, 53, 54])
@pytest.mark.parametrize("p,r", [(1, 2), (1, 3), (2, 2), (3, 2), (4, 1), (4, 2), (4, 3), (3, 3), (2, 0), (5, 2), (5, 0)])
def test_kernel_variants_match_oracle(ctx, p, r, variant, dtype):
    """variant 1 = column kernel, 40 = staged kernel, 51..54 = slab3 kernel in its four flavours (register / cp.async
    gather x no / early face merges); cell counts that are not a multiple of the warp group size exercise the tail
    handling; the second call reuses the kernel's private arrays."""
    import dealii_cuda_b200 as mf
    if variant == 40 and p == 1:
        pytest.skip("the staged kernel starts at degree 2")
    o = OracleMesh(3, p, r)
    m = mf.HyperCubeMesh(ctx, 3, p, r)
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(m)
    op.set_variant(variant)
    assert op.active_variant() == (variant if variant in (1, 40) else 50)
    u = sm64(7, o.n_dofs).astype(dtype)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs, dtype)
    for _ in range(2):
        dst.fill(-3.0)
        op.vmult(dst, src)
        got = dst.toVector()
        assert rel_err(got, o.vmult(u.astype(np.float64))) <= TOL[dtype]
        assert np.array_equal(got[o.constrained], u[o.constrained])
    d0 = sm64(8, o.n_dofs).astype(dtype)
    dst.fromHost(d0)
    op.vmult_add(dst, src)
    assert rel_err(dst.toVector(), o.vmult_add(d0.astype(np.float64), u.astype(np.float64))) <= TOL[dtype]


def test_unknown_variant_is_rejected(ctx):
    import dealii_cuda_b200 as mf
    m = mf.HyperCubeMesh(ctx, 3, 4, 1)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    op.set_variant(9)  # (a slab2 configuration of round 1)
    a, b = mf.GpuVector(ctx, m.n_dofs), mf.GpuVector(ctx, m.n_dofs)
    with pytest.raises(mf.MfgError):
        op.vmult(a, b)


STAGE_CASES = [(2, 2), (2, 3), (3, 2), (3, 3), (4, 0), (4, 1), (4, 2), (4, 3), (5, 1), (5, 2)]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("p,r", STAGE_CASES)
def test_staged_kernel_matches_oracle(ctx, p, r, dtype):
    """variant 40 = staged kernel (kernels_stage.cuh: asynchronous staged gather, register face merges, staged scatter with
    plain stores for group-interior DoFs): vmult into a dirty vector (the plain stores and the zero pass must cover every
    DoF), vmult_add, and agreement with the slab3 kernel it shares its contractions with"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, p, r)
    m = mf.HyperCubeMesh(ctx, 3, p, r)
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(m)
    op.set_variant(40)
    assert op.active_variant() == 40
    u = sm64(3, o.n_dofs).astype(dtype)
    src = mf.GpuVector.from_numpy(ctx, u)
    dst = mf.GpuVector(ctx, o.n_dofs, dtype)
    dst.fill(-7.5)
    op.vmult(dst, src)
    want = o.vmult(u.astype(np.float64))
    assert rel_err(dst.toVector(), want) <= TOL[dtype]
    assert np.array_equal(src.toVector(), u)  # src untouched
    op.vmult_add(dst, src)
    assert rel_err(dst.toVector(), 2.0 * want) <= 2 * TOL[dtype]
    st = op.stage_stats()
    assert st["groups"] == (o.n_cells + 32 // (p + 1) - 1) // (32 // (p + 1)) and st["staged"] + st["fallback"] == st["groups"]
    op6 = mf.LaplaceOperatorGpu(ctx, dtype)
    op6.reinit(m)
    op6.set_variant(51)
    d6 = mf.GpuVector(ctx, o.n_dofs, dtype)
    op6.vmult(d6, src)
    assert rel_err(dst.toVector(), 2.0 * d6.toVector().astype(np.float64)) <= (1e-14 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("variant", [40, 50])
def test_grouped_kernels_repeated_applies_and_box(ctx, variant):
    """the bmop loop (bmop.cu:135-153) on a non-cubic box whose last group is partial, against the oracle"""
    import dealii_cuda_b200 as mf
    box = dict(log2_cells=(2, 1, 3), origin=(-1.0, -0.5, 0.0), h=0.25)
    o = OracleMesh(3, 4, box=box)
    m = mf.HyperCubeMesh(ctx, 3, 4, box=box)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    op.set_variant(variant)
    assert op.active_variant() == variant
    a, b = mf.GpuVector(ctx, o.n_dofs), mf.GpuVector(ctx, o.n_dofs)
    op.bmop(a, b, 3, 0.1)
    assert rel_err(a.toVector(), o.bmop(3, 0.1)) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("p,r", [(1, 2), (1, 3), (2, 2), (2, 3), (3, 2), (3, 3), (4, 0), (4, 1), (4, 2), (4, 3), (5, 1), (5, 2)])
def test_slab3_kernel_matches_oracle(ctx, p, r, dtype):
    """variant 50 = slab3 kernel (kernels_slab3.cuh: asynchronous gather one group ahead, face merges in registers right
    after the contraction across the face), the default for 3D degree 1..5 with the atomic scatter"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, p, r)
    m = mf.HyperCubeMesh(ctx, 3, p, r)
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(m)
    assert op.active_variant() == 50
    u = sm64(4, o.n_dofs).astype(dtype)
    src = mf.GpuVector.from_numpy(ctx, u)
    dst = mf.GpuVector(ctx, o.n_dofs, dtype)
    dst.fill(-7.5)
    op.vmult(dst, src)
    want = o.vmult(u.astype(np.float64))
    assert rel_err(dst.toVector(), want) <= TOL[dtype]
    assert np.array_equal(src.toVector(), u)
    op.vmult_add(dst, src)
    assert rel_err(dst.toVector(), 2.0 * want) <= 2 * TOL[dtype]


def test_slab_variants_rejected_where_unsupported(ctx):
    import dealii_cuda_b200 as mf
    for dim, p, coloring, auto in [(2, 4, False, 1), (3, 5, False, 50), (3, 4, True, 1), (3, 6, False, 1)]:
        m = mf.HyperCubeMesh(ctx, dim, p, 1)
        op = mf.LaplaceOperatorGpu(ctx, np.float64, use_coloring=coloring)
        op.reinit(m)
        assert op.active_variant() == auto  # auto: slab3 kernel for 3D degree 1..5 with atomics, column kernel otherwise
        if auto == 1:
            for v in (40, 50):
                op.set_variant(v)
                a, b = mf.GpuVector(ctx, m.n_dofs), mf.GpuVector(ctx, m.n_dofs)
                with pytest.raises(mf.MfgError):
                    op.vmult(a, b)


def test_golden_fixtures(ctx):
    import dealii_cuda_b200 as mf
    for c in json.load(open(os.path.join(GOLD, "apply_cases.json"))):
        g = np.load(os.path.join(GOLD, c["file"]))
        m = mf.HyperCubeMesh(ctx, c["dim"], c["p"], c["r"], c["left"], c["right"])
        assert np.array_equal(m.loc2glob(), g["loc2glob"])
        assert np.array_equal(m.constrained_dofs(), g["constrained"])
        op = mf.LaplaceOperatorGpu(ctx, np.float64)
        op.reinit(m)
        src = mf.GpuVector.from_numpy(ctx, sm64(c["seed"], m.n_dofs))
        dst = mf.GpuVector(ctx, m.n_dofs)
        op.vmult(dst, src)
        assert rel_err(dst.toVector(), g["Au"]) <= 1e-12
        # bmop loop, 3 applications from dst = 0.1
        a, b = mf.GpuVector(ctx, m.n_dofs), mf.GpuVector(ctx, m.n_dofs)
        op.bmop(a, b, 3, 0.1)
        assert rel_err(a.toVector(), g["bmop3"]) <= 1e-11  # three chained applications
        op.compute_diagonal()
        assert rel_err(op.get_diagonal_inverse().toVector(), g["inv_diag"]) <= 1e-12


@pytest.mark.parametrize("dim,p,r,k", [(3, 4, 2, 100), (3, 4, 2, 3), (2, 4, 3, 3)])
def test_bmop_100_applications(ctx, dim, p, r, k):
    """bmop.cu:135-153: the raw loop grows like lambda_max^100 -> relative comparison."""
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    want = o.bmop(k)
    m = mf.HyperCubeMesh(ctx, dim, p, r)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    a, b = mf.GpuVector(ctx, m.n_dofs), mf.GpuVector(ctx, m.n_dofs)
    ms = op.bmop(a, b, k, 0.1)
    assert ms > 0
    got = a.toVector()
    assert np.all(np.isfinite(got))
    # 100 chained applications amplify roundoff (the loop is an un-normalised power iteration): on the
    # CPU alone the oracle's matrix-free loop differs from the loop with its own assembled matrix by 1.6e-10
    # (3D Q4 r=2), FP64 from long double by 4e-11.  Calibrate the tolerance with exactly that CPU-only
    # difference; single applications are held to 1e-12 in test_vmult_matches_oracle.
    if o.n_dofs > 6000:  # dense calibration matrix too large: use the value calibrated at r=2
        assert rel_err(got, want) <= 1e-8 and max_rel_err(got, want) <= 1e-8
        return
    K = o.assemble_dense()
    x = np.full(o.n_dofs, 0.1)
    for _ in range(k):
        x = K @ x
    cal = max(rel_err(x, want), max_rel_err(x, want))
    assert rel_err(got, want) <= max(1e-10, 20 * cal)
    assert max_rel_err(got, want) <= max(1e-10, 20 * cal)


def test_bmop_fp32_renormalised(ctx):
    """FP32 overflows in the raw 100-loop: compare a per-step renormalised loop.  The loop amplifies
    roundoff by ~50x within 3 steps and ~1e5x within 10 (calibrated in FP64 on the CPU, see above), so only k = 2
    chained FP32 steps can be held to the 1e-5 of a single application."""
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, 4, 2)
    m = mf.HyperCubeMesh(ctx, 3, 4, 2)
    op = mf.LaplaceOperatorGpu(ctx, np.float32)
    op.reinit(m)
    u = np.full(o.n_dofs, 0.1)
    d, s = mf.GpuVector(ctx, o.n_dofs, np.float32), mf.GpuVector(ctx, o.n_dofs, np.float32)
    d.fill(0.1)
    for _ in range(2):
        d.swap(s)
        op.vmult(d, s)
        d *= 1.0 / d.l2_norm()
        u = o.vmult(u); u /= np.linalg.norm(u)
    assert rel_err(d.toVector(), u) <= 1e-5


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_vmult_add_and_constrained_rows(ctx, dtype):
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, 3, 2)
    m = mf.HyperCubeMesh(ctx, 3, 3, 2)
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(m)
    u, d0 = sm64(5, o.n_dofs).astype(dtype), sm64(6, o.n_dofs).astype(dtype)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector.from_numpy(ctx, d0)
    op.vmult_add(dst, src)
    got = dst.toVector()
    want = o.vmult_add(d0.astype(np.float64), u.astype(np.float64))
    assert rel_err(got, want) <= TOL[dtype]
    con = o.constrained
    assert np.array_equal(got[con], (d0[con] + u[con]).astype(dtype))


def test_inverse_diagonal(ctx):
    import dealii_cuda_b200 as mf
    for dim, p, r in [(2, 4, 2), (3, 2, 2), (3, 4, 1), (3, 7, 1)]:
        o = OracleMesh(dim, p, r)
        m = mf.HyperCubeMesh(ctx, dim, p, r)
        op = mf.LaplaceOperatorGpu(ctx, np.float64)
        op.reinit(m)
        with pytest.raises(mf.MfgError):
            op.get_diagonal_inverse()  # Assert(diagonal_is_available)
        op.compute_diagonal()
        assert rel_err(op.get_diagonal_inverse().toVector(), o.inverse_diagonal()) <= 1e-12


def test_explicit_array_reinit_matches_mesh_path(ctx):
    """mfg_mf_reinit (what a deal.II based caller uses): loc2glob / inv_jac / coefficient as host arrays."""
    import dealii_cuda_b200 as mf
    for dim, p, r, coloring in [(3, 4, 2, False), (2, 3, 3, False), (3, 2, 2, True)]:
        o = OracleMesh(dim, p, r)
        l2g, coef = o.loc2glob.copy(), o.coefficient.copy()
        arrays = dict(dim=dim, degree=p, n_dofs=o.n_dofs, loc2glob=l2g, inv_jac=np.full(o.n_cells, (1 << r) / 2.0))
        if coloring:
            cc = o.cell_coords
            col = (cc[:, 0] & 1) + 2 * (cc[:, 1] & 1) + 4 * (cc[:, 2] & 1)
            order = np.argsort(col, kind="stable")
            arrays["loc2glob"], coef = l2g[order], coef[order]
            arrays["color_offsets"] = np.concatenate(([0], np.cumsum(np.bincount(col, minlength=1 << dim)))).astype(np.uint32)
        data = mf.MatrixFreeGpu(ctx, np.float64)
        data.reinit(arrays, use_coloring=coloring)
        assert data.n_dofs == o.n_dofs and data.n_cells_tot == o.n_cells
        ch = mf.ConstraintHandlerGpu(ctx, np.float64)
        ch.reinit(o.constrained, o.n_dofs)
        op = mf.LaplaceOperatorGpu(ctx, np.float64, use_coloring=coloring)
        op.reinit(data, ch, coefficient=coef)
        u = sm64(9, o.n_dofs)
        src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs)
        op.vmult(dst, src)
        assert rel_err(dst.toVector(), o.vmult(u)) <= 1e-12


def test_vmult_host_and_ptr_entry_points(ctx):
    import torch
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, 4, 2)
    m = mf.HyperCubeMesh(ctx, 3, 4, 2)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    u = sm64(4, o.n_dofs)
    want = o.vmult(u)
    out = np.empty_like(u)
    op.vmult_host(out, u)
    assert rel_err(out, want) <= 1e-12
    # pipelined host API: two slots in flight, different inputs
    u2 = sm64(5, o.n_dofs)
    hs = [torch.from_numpy(u).pin_memory(), torch.from_numpy(u2).pin_memory()]
    hd = [torch.empty(o.n_dofs, dtype=torch.float64).pin_memory() for _ in range(2)]
    for rep in range(3):
        for k in range(2):
            op.vmult_host_async(hd[k].numpy(), hs[k].numpy(), k)
    op.host_sync()
    assert rel_err(hd[0].numpy(), want) <= 1e-12 and rel_err(hd[1].numpy(), o.vmult(u2)) <= 1e-12
    ts = torch.from_numpy(u).cuda(); td = torch.empty_like(ts)
    torch.cuda.synchronize()
    op.vmult_ptr(td.data_ptr(), ts.data_ptr())
    ctx.synchronize()
    assert rel_err(td.cpu().numpy(), want) <= 1e-12
    # wrapped torch tensors as GpuVectors
    vs, vd = mf.GpuVector.wrap(ctx, ts), mf.GpuVector.wrap(ctx, td)
    td.zero_(); torch.cuda.synchronize()
    op.vmult(vd, vs); ctx.synchronize()
    assert rel_err(td.cpu().numpy(), want) <= 1e-12


def test_cpp_facade_bmop_example():
    """examples/bmop.cc: the reference's bmop.cu driver on the header-only C++ facade (host code built with g++)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["make", "-C", os.path.join(root, "examples"), "-s"])
    out = subprocess.run([os.path.join(root, "examples", "_build", "bmop"), "3", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert [r[:3] for r in rows] == [["3", "4", "4913"], ["3", "4", "35937"]]  # dim, degree, n_dofs as bmop.cu:152 prints them
    assert all(float(r[3]) > 0 for r in rows)


def test_error_behaviour(ctx):
    import dealii_cuda_b200 as mf
    with pytest.raises(mf.MfgError):
        mf.HyperCubeMesh(ctx, 3, 9, 1)  # degree > 8
    with pytest.raises(mf.MfgError):
        mf.HyperCubeMesh(ctx, 4, 2, 1)
    m = mf.HyperCubeMesh(ctx, 3, 2, 1)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(m)
    a, b = mf.GpuVector(ctx, m.n_dofs), mf.GpuVector(ctx, m.n_dofs + 1)
    with pytest.raises(mf.MfgError):
        op.vmult(a, b)  # size mismatch
    with pytest.raises(mf.MfgError):
        op.vmult(a, a)  # aliasing
    c = mf.GpuVector(ctx, m.n_dofs, np.float32)
    with pytest.raises(mf.MfgError):
        op.vmult(a, c)  # dtype mismatch


# ---- full-size cases: size-independent properties + direct comparison with the threaded oracle ----

@pytest.mark.parametrize("coloring", [False, True])
def test_full_size_r5_against_threaded_oracle(ctx, coloring):
    """BASELINE configs[0]: 3D Q4 r=5, 2,146,689 DoFs."""
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, 4, 5)
    assert o.n_dofs == 2146689 and o.n_constrained == 98306
    m = mf.HyperCubeMesh(ctx, 3, 4, 5)
    assert np.array_equal(m.loc2glob(), o.loc2glob)
    assert np.array_equal(m.constrained_dofs(), o.constrained)
    op = mf.LaplaceOperatorGpu(ctx, np.float64, use_coloring=coloring)
    op.reinit(m)
    u = sm64(1, o.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs)
    op.vmult(dst, src)
    want = o.vmult(u, threaded=True)
    assert rel_err(dst.toVector(), want) <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_full_size_r6_against_threaded_oracle(ctx, dtype):
    """the headline configuration itself (3D Q4 r=6, 16,974,593 DoFs) against the oracle, default kernel"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(3, 4, 6)
    m = mf.HyperCubeMesh(ctx, 3, 4, 6)
    assert m.n_dofs == o.n_dofs == 16974593
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(m)
    u = sm64(1, o.n_dofs).astype(dtype)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs, dtype)
    op.vmult(dst, src)
    want = o.vmult(u.astype(np.float64), threaded=True)
    assert rel_err(dst.toVector(), want) <= TOL[dtype]
    assert max_rel_err(dst.toVector(), want) <= 10 * TOL[dtype]


def test_full_size_r6_properties(ctx):
    """1-GPU roofline point: 3D Q4 r=6, 16,974,593 DoFs.  Symmetry, linearity, constrained-row identity,
    atomic == colored, A*1 = 0 on rows away from the boundary for a constant coefficient."""
    import dealii_cuda_b200 as mf
    m = mf.HyperCubeMesh(ctx, 3, 4, 6)
    assert m.n_dofs == 16974593 and m.n_cells == 262144 and m.n_constrained == 257 ** 3 - 255 ** 3
    n = m.n_dofs
    opa = mf.LaplaceOperatorGpu(ctx, np.float64); opa.reinit(m)
    opc = mf.LaplaceOperatorGpu(ctx, np.float64, use_coloring=True); opc.reinit(m)
    u, v = sm64(1, n), sm64(2, n)
    gu, gv = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector.from_numpy(ctx, v)
    Au, Av, Auc, w = (mf.GpuVector(ctx, n) for _ in range(4))
    opa.vmult(Au, gu); opa.vmult(Av, gv); opc.vmult(Auc, gu)
    # symmetry
    vAu, uAv = gv.dot(Au), gu.dot(Av)
    assert abs(vAu - uAv) <= 1e-11 * abs(vAu)
    # atomics vs coloring: same operator up to summation order
    d = Au.toVector() - Auc.toVector()
    assert np.linalg.norm(d) <= 1e-13 * Au.l2_norm()
    # linearity: A(2u - 3v) = 2Au - 3Av
    w.equ(2.0, gu); w.add(-3.0, gv)
    Aw = mf.GpuVector(ctx, n); opa.vmult(Aw, w)
    Aw.add(-2.0, Au); Aw.add(3.0, Av)
    assert Aw.l2_norm() <= 1e-12 * Au.l2_norm()
    # constrained rows: identity, bit-exact
    con = m.constrained_dofs()
    assert np.array_equal(Au.toVector()[con], u[con])
    # constant coefficient: A*1 vanishes on all rows whose cells touch no constrained DoF
    opa.set_coefficient(np.ones((m.n_cells, m.dofs_per_cell)))
    one = mf.GpuVector(ctx, n); one.fill(1.0)
    opa.vmult(Au, one)
    r = Au.toVector()
    l2g = m.loc2glob()
    cflag = np.zeros(n, dtype=bool); cflag[con] = True
    touched = np.zeros(n, dtype=bool)
    bcells = cflag[l2g].any(axis=1)
    touched[l2g[bcells].ravel()] = True
    assert np.abs(r[~touched]).max() <= 1e-12


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,p", [(2, 2), (3, 4)])
def test_empty_mesh_from_arrays(ctx, dim, p, dtype):
    """an operator without cells (empty partition): vmult is the identity on the constrained rows and zero elsewhere,
    vmult_add leaves dst + src[c]; the diagonal is 1 on constrained rows"""
    import dealii_cuda_b200 as mf
    npc = (p + 1) ** dim
    data = dict(dim=dim, degree=p, n_dofs=7, loc2glob=np.zeros((0, npc), np.uint32), inv_jac=np.zeros(0))
    mfree = mf.MatrixFreeGpu(ctx, dtype)
    mfree.reinit(data)
    ch = mf.ConstraintHandlerGpu(ctx, dtype)
    ch.reinit(np.array([1, 4], np.uint32), 7)
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(mfree, ch, coefficient=np.zeros((0, npc)))
    u = np.arange(1, 8).astype(dtype)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, 7, dtype)
    dst.fill(9.0)
    op.vmult(dst, src)
    assert np.array_equal(dst.toVector(), np.array([0, 2, 0, 0, 5, 0, 0], dtype))
    op.vmult_add(dst, src)
    assert np.array_equal(dst.toVector(), np.array([0, 4, 0, 0, 10, 0, 0], dtype))
    op.compute_diagonal()
    assert np.array_equal(op.get_diagonal_inverse().toVector()[[1, 4]], np.ones(2, dtype))
