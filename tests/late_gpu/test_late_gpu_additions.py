"""GPU tests added late in round 2, AFTER the round's GPU budget was spent: they exercise code that compiles and whose host
logic is covered on the CPU, but they have not run on hardware yet.  They are not collected by the main suite: the wrapper
tests/test_z_late_gpu_additions.py (last in the suite) runs this file in a CHILD pytest process and reports every test function,
so that nothing here -- not even a crash of the process -- can mask or take down the hardware-verified tests in front of it.
  * generic FEEvaluationGpu path: MatrixFreeGpu::cell_loop(dst, loc_op) (matrix_free_gpu.h:382-393), evaluate_on_cells<Op>
    (:415-435), hanging-node interpolation in read_dof_values / distribute_local_to_global (fee_gpu.cuh:333-351);
  * the restated deal.II graph coloring (coloring.cc:8-33) driving the atomics-free scatter;
  * the library's adaptive-mesh substrate (mfg_amesh_*) feeding LaplaceOperatorGpu, the C++ facade on it;
  * multigrid with local smoothing on adaptive meshes (mfg_amg_*) against oracle/adaptive_mg.py;
  * the assembled sparse-matrix competitor (mfg_spm_*)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # tests/ (the generic-path fixtures live there)
from oracle.oracle import OracleMesh, sm64  # noqa: E402  (checker)
# the compiled examples (examples/_build, built by __graft_entry__.build()); the emulation run points this at its own builds
EXAMPLES_BUILD = os.environ.get("MFG_EXAMPLES_BUILD") or os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                                                      "examples", "_build")
from test_gpu_generic_path import gen, rel_err  # noqa: E402,F401  (fixture: the compiled example functors)


@pytest.mark.parametrize("dim,p,r", [(2, 2, 2), (3, 2, 1), (3, 4, 1)])
def test_dst_only_cell_loop_and_evaluate_on_cells(ctx, gen, dim, p, r):
    """MatrixFreeGpu::cell_loop(dst, loc_op) (matrix_free_gpu.h:382-393) and evaluate_on_cells<LocalCoeffOp> (:415-435,
    laplace_operator_gpu.h:191-211): the dst-only loop gives the same right-hand side as the loop with a source vector, the
    evaluated coefficient is 1 / (0.05 + 2 |x_q|^2) at the oracle's quadrature points"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    m = mf.HyperCubeMesh(ctx, dim, p, r)
    mfree = mf.MatrixFreeGpu(ctx, np.float64)
    mfree.reinit(m)
    dummy = mf.GpuVector(ctx, o.n_dofs)
    a, b = mf.GpuVector(ctx, o.n_dofs), mf.GpuVector(ctx, o.n_dofs)
    a.fill(0.0); b.fill(0.0)
    gen(mfree, 2, dim, p, np.float64, a, dummy)
    gen(mfree, 3, dim, p, np.float64, b, dummy)
    assert rel_err(b.toVector(), a.toVector()) <= 1e-14
    coef = mf.GpuVector(ctx, o.n_cells * (p + 1) ** dim)
    gen(mfree, 4, dim, p, np.float64, coef, dummy)
    assert rel_err(coef.toVector(), np.asarray(o.coefficient).ravel()) <= 1e-14


@pytest.mark.parametrize("dim,p,r", [(2, 3, 3), (3, 2, 2), (3, 4, 2)])
def test_operator_with_restated_dealii_coloring(ctx, dim, p, r):
    """cells sorted by the restated deal.II colors, atomics-free scatter (use_coloring): same operator as the oracle"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    l2g = np.asarray(o.loc2glob)
    color, nc = mf.graph_coloring(l2g, o.n_dofs)
    perm = np.argsort(color, kind="stable")
    offsets = np.concatenate([[0], np.cumsum(np.bincount(color, minlength=nc))]).astype(np.uint32)
    n = p + 1
    h = 2.0 / round(o.n_cells ** (1.0 / dim))
    data = mf.MatrixFreeGpu(ctx, np.float64)
    data.reinit(dict(dim=dim, degree=p, n_dofs=o.n_dofs, loc2glob=l2g[perm], inv_jac=np.full(o.n_cells, 1.0 / h), color_offsets=offsets),
                use_coloring=True)
    assert data.num_colors == nc
    ch = mf.ConstraintHandlerGpu(ctx, np.float64)
    ch.reinit(np.asarray(o.constrained), o.n_dofs)
    op = mf.LaplaceOperatorGpu(ctx, np.float64, use_coloring=True)
    op.reinit(data, ch, coefficient=np.asarray(o.coefficient)[perm])
    u = sm64(5, o.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs)
    op.vmult(dst, src)
    want = o.vmult(u)
    assert np.linalg.norm(dst.toVector() - want) <= 1e-12 * np.linalg.norm(want)


@pytest.mark.parametrize("dim,p", [(2, 3), (3, 2), (3, 4)])
def test_operator_on_library_built_adaptive_mesh(ctx, dim, p):
    """LaplaceOperatorGpu::reinit on an adaptive mesh built by the LIBRARY's host substrate (mfg_amesh_*, tests/test_adaptive_mesh.py
    pins its arrays to the oracle on the CPU): vmult, the Jacobi diagonal and a CG solve against oracle/adaptive.py on the same cells"""
    import dealii_cuda_b200 as mf
    from oracle.adaptive import AdaptiveMesh
    am = mf.AdaptiveMesh(dim, p).refine_global(2)
    am.mark_cells_in_annulus(0.8, 0.0, None); am.execute_coarsening_and_refinement()
    am.mark_cells_in_annulus(0.45, 0.1, (-0.1, -0.2, -0.3)); am.execute_coarsening_and_refinement()
    am.distribute_dofs()
    o = AdaptiveMesh(dim, p, 0, [], cells=am.active_cells().tolist())
    assert o.mask.max() > 0 and am.n_dofs == o.n_dofs
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(am)
    u = sm64(11, o.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, o.n_dofs)
    op.vmult(dst, src)
    want = o.vmult(u)
    assert np.linalg.norm(dst.toVector() - want) <= 1e-12 * np.linalg.norm(want)
    op.compute_diagonal()
    assert rel_err(op.get_diagonal_inverse().toVector(), o.inverse_diagonal()) <= 1e-12


def test_bmop_driver_on_the_pseudo_adaptive_mesh():
    """examples/bmop.cc built with -DADAPTIVE_GRID (bmop.cu:170-181): pseudo_adaptive_refinement + hanging-node operator through
    the C++ facade; the DoF counts are those of the library's host substrate (checked on the CPU in tests/test_adaptive_mesh.py)"""
    import os
    import subprocess
    exe = os.path.join(EXAMPLES_BUILD, "bmop_adaptive")
    assert os.path.exists(exe), "examples/_build/bmop_adaptive is missing: run __graft_entry__.build()"
    out = subprocess.run([exe, "4", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert [int(r[2]) for r in rows] == [729, 57142]
    assert all(float(r[3]) > 0 for r in rows)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,p,r", [(2, 3, 3), (3, 2, 2), (3, 4, 2)])
def test_sparse_matrix_vmult(ctx, dim, p, r, dtype):
    """SparseMatrix::vmult (cuda_sparse_matrix.cu:414-429) against the oracle and against the matrix-free operator"""
    import dealii_cuda_b200 as mf
    o = OracleMesh(dim, p, r)
    mesh = mf.HyperCubeMesh(ctx, dim, p, r)
    S = mf.SparseMatrixGpu(ctx, dtype)
    S.reinit(mesh)
    assert S.m() == o.n_dofs and S.n_nonzero_elements() > o.n_dofs
    u = sm64(3, o.n_dofs); u[o.constrained] = 0.0
    src, dst = mf.GpuVector.from_numpy(ctx, u.astype(dtype)), mf.GpuVector(ctx, o.n_dofs, dtype)
    S.vmult(dst, src)
    want = o.vmult(u)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    assert np.linalg.norm(dst.toVector().astype(np.float64) - want) <= tol * np.linalg.norm(want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,p,r", [(2, 2, 1), (2, 4, 1), (3, 1, 1), (3, 2, 1), (3, 4, 0)])
def test_operator_on_the_ball_mesh(ctx, dim, p, r, dtype):
    """LaplaceOperatorGpu::reinit on the library's BALL_GRID substrate (mfg_umesh_*: hyper_ball + SphericalManifold + refine_global,
    poisson_common.h:59-72; tests/test_ball_mesh.py pins its arrays on the CPU): general geometry, against the dense numpy operator
    built from the same loc2glob / K / JxW / coefficient (fee_gpu.cuh:236-240, 276-280), identity on the Dirichlet rows"""
    import dealii_cuda_b200 as mf
    bm = mf.BallMesh(dim, p, r).distribute_dofs()
    a = bm.arrays()
    n, npc = p + 1, (p + 1) ** dim
    val, grad, _, _ = (np.asarray(t) for t in mf.shape_info(p))
    q = np.arange(npc)
    qi = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    B = np.zeros((dim, npc, npc))                        # [d][q][i]
    for d in range(dim):
        t = np.ones((npc, npc))
        for e in range(dim):
            t *= (grad if e == d else val)[qi[None, :, e], qi[:, None, e]]
        B[d] = t
    A = np.zeros((bm.n_dofs, bm.n_dofs))
    for c in range(bm.n_cells):
        gx = np.einsum("qed,eqi->qdi", a["inv_jac"][c], B)             # grad_x phi_i at q = K^T grad_xi phi_i
        row = a["loc2glob"][c].astype(np.int64)
        A[np.ix_(row, row)] += np.einsum("qdi,q,qdj->ij", gx, a["coefficient"][c] * a["JxW"][c], gx)
    con = a["boundary"].astype(np.int64)
    A[con, :] = 0.0; A[:, con] = 0.0; A[con, con] = 1.0
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(bm)
    u = sm64(13, bm.n_dofs)
    src, dst = mf.GpuVector.from_numpy(ctx, u.astype(dtype)), mf.GpuVector(ctx, bm.n_dofs, dtype)
    op.vmult(dst, src)
    want = A @ u
    assert rel_err(dst.toVector(), want) <= (1e-12 if dtype == np.float64 else 2e-5)
    assert np.array_equal(dst.toVector()[con], u.astype(dtype)[con])
    if dtype == np.float64:
        op.compute_diagonal()
        assert rel_err(op.get_diagonal_inverse().toVector(), 1.0 / np.diag(A)) <= 1e-12


def test_bmop_driver_on_the_ball_mesh():
    """examples/bmop.cc built with -DBALL_GRID (bmop.cu:164-168) through the C++ facade"""
    import subprocess
    exe = os.path.join(EXAMPLES_BUILD, "bmop_ball")
    assert os.path.exists(exe), "examples/_build/bmop_ball is missing: run __graft_entry__.build()"
    out = subprocess.run([exe, "2", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert [int(r[2]) for r in rows] == [3817, 29521] and all(float(r[3]) > 0 for r in rows)


@pytest.mark.parametrize("dim,p,rmin,rmax", [(2, 2, 2, 4), (2, 4, 1, 3), (3, 2, 1, 3), (3, 4, 1, 2)])
def test_poisson_on_the_ball_converges(dim, p, rmin, rmax):
    """examples/poisson.cu with domain = BALL (poisson_common.h:65-70): non-affine cells in the precompiled operator (general
    geometry) and in the user-written functors of the generic path.  The numpy restatement on the same mesh arrays gives L2 error
    ratios 7.5 / 7.7 (2D Q2), 21 / 36 (2D Q4), 6.6 / 6.9 (3D Q2), 19 (3D Q4, still pre-asymptotic) per refinement."""
    import subprocess
    exe = os.path.join(EXAMPLES_BUILD, "poisson")
    out = subprocess.run([exe, str(dim), str(p), str(rmin), str(rmax), "ball"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(rows) == rmax - rmin + 1
    errs = [float(r[5]) for r in rows]
    for a, b in zip(errs[:-1], errs[1:]):
        assert 0.4 * 2 ** (p + 1) <= a / b <= 2.0 * 2 ** (p + 1), (errs, "expected about a factor 2^(p+1) per refinement")


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,p", [(2, 2), (2, 4), (3, 1), (3, 2), (3, 3)])
def test_generic_path_interpolates_hanging_nodes(ctx, gen, dim, p, dtype):
    """read_dof_values / distribute_local_to_global of the generic FEEvaluationGpu path apply resolve_hanging_nodes_shmem
    (fee_gpu.cuh:333-335, 349-351) on the cells that carry a constraint mask: a user-written MASS operator on an adaptive mesh
    against numpy (gather through the rewritten map, interpolate, local mass matrix, transposed interpolation, scatter)"""
    import dealii_cuda_b200 as mf
    from oracle.adaptive import AdaptiveMesh, resolve_hanging_nodes
    am = mf.AdaptiveMesh(dim, p).refine_global(2 if dim == 2 else 1)
    am.mark_cells_in_annulus(0.9, 0.0, None); am.execute_coarsening_and_refinement()
    am.mark_cells_in_annulus(0.5, 0.0, (-0.1, -0.2, -0.3)); am.execute_coarsening_and_refinement()
    am.distribute_dofs()
    a = am.arrays()
    assert a["constraint_mask"].max() > 0
    o = AdaptiveMesh(dim, p, 0, [], cells=am.active_cells().tolist())     # (resolve_hanging_nodes needs no mesh; o gives h per cell)
    n = p + 1
    _, _, xq, wq = mf.shape_info(p)
    N = np.asarray(mf.shape_info(p)[0])                                    # [i][q]
    q = np.arange(n ** dim)
    q_idx = np.stack([(q // n ** e) % n for e in range(dim)], axis=1)
    Nq = np.ones((n ** dim, n ** dim))
    for e in range(dim):
        Nq *= N[q_idx[None, :, e], q_idx[:, None, e]]                      # [q][i]
    wref = np.prod(wq[q_idx], axis=1)
    u = sm64(31, am.n_dofs)
    want = np.zeros(am.n_dofs)
    for ci in range(am.n_cells):
        row, mask = a["loc2glob"][ci].astype(np.int64), int(a["constraint_mask"][ci])
        ul = resolve_hanging_nodes(u[row].reshape((n,) * dim), mask, p, dim, transpose=False).ravel()
        v = Nq.T @ ((o.h[ci] ** dim * wref) * (Nq @ ul))
        v = resolve_hanging_nodes(v.reshape((n,) * dim), mask, p, dim, transpose=True).ravel()
        np.add.at(want, row, v)
    mfree = mf.MatrixFreeGpu(ctx, dtype)
    mfree.reinit(dict(dim=dim, degree=p, n_dofs=am.n_dofs, loc2glob=a["loc2glob"], inv_jac=a["inv_jac"], constraint_mask=a["constraint_mask"]))
    src, dst = mf.GpuVector.from_numpy(ctx, u.astype(dtype)), mf.GpuVector(ctx, am.n_dofs, dtype)
    dst.fill(0.0)
    gen(mfree, 0, dim, p, dtype, dst, src)
    assert rel_err(dst.toVector(), want) <= (1e-12 if dtype == np.float64 else 2e-5)
    # the reference's own Laplace LocalOperator on the same path: the free rows of the oracle's hanging-node operator.  Per-point user
    # arrays are indexed in kernel cell order = cells without a mask first (stable), get_global_q
    perm = np.concatenate([np.nonzero(a["constraint_mask"] == 0)[0], np.nonzero(a["constraint_mask"] != 0)[0]])
    u0 = u.copy(); u0[o.constrained] = 0.0
    cdev = mf.GpuVector.from_numpy(ctx, a["coefficient"][perm].reshape(-1).astype(dtype))
    src0 = mf.GpuVector.from_numpy(ctx, u0.astype(dtype))
    dst.fill(0.0)
    gen(mfree, 1, dim, p, dtype, dst, src0, cdev)
    free = np.ones(am.n_dofs, bool); free[o.constrained] = False
    assert rel_err(dst.toVector()[free], o.vmult(u0)[free]) <= (1e-12 if dtype == np.float64 else 2e-5)


@pytest.mark.parametrize("dim,p,rmin,rmax", [(2, 2, 3, 6), (2, 4, 2, 5), (3, 2, 2, 4), (3, 4, 2, 4)])
def test_poisson_on_a_locally_refined_mesh_converges_at_the_optimal_rate(dim, p, rmin, rmax):
    """examples/poisson.cu with grid_refinement = NONUNIFORM (poisson_common.h:76-92): hanging-node constraints in the precompiled
    operator, in the user-written right-hand-side / error functors of the generic path and in the constraint handler at once.
    Against the ANALYTIC solution of poisson_common.cc the L2 error must still fall like h^(p+1) per refinement of the family."""
    import subprocess
    exe = os.path.join(EXAMPLES_BUILD, "poisson")
    assert os.path.exists(exe), "examples/_build/poisson is missing: run __graft_entry__.build()"
    out = subprocess.run([exe, str(dim), str(p), str(rmin), str(rmax), "nonuniform"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(rows) == rmax - rmin + 1
    errs = [float(r[5]) for r in rows]
    for a, b in zip(errs[:-1], errs[1:]):
        assert 0.5 * 2 ** (p + 1) <= a / b <= 2.0 * 2 ** (p + 1), (errs, "expected a factor 2^(p+1) per refinement")
    assert all(int(r[4]) > 0 for r in rows)


def _adaptive_mg_case(ctx, dim, p, base, steps, min_level):
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import AdaptiveMultigrid
    from oracle.adaptive import AdaptiveMesh
    from oracle.adaptive_mg import AdaptiveMultigridOracle
    am = mf.AdaptiveMesh(dim, p, limit_level_difference_at_vertices=True).refine_global(base)
    for R, r, c in steps:
        am.mark_cells_in_annulus(R, r, c)
        am.execute_coarsening_and_refinement()
    am.distribute_dofs()
    mg = AdaptiveMultigrid(ctx, am, min_level=min_level)
    o = AdaptiveMesh(dim, p, 0, [], cells=am.active_cells().tolist())
    lc = {l: [tuple(int(v) for v in row) for row in am.level_cells(l)] for l in range(am.n_levels)}
    return am, mg, o, AdaptiveMultigridOracle(dim, p, lc, o, min_level=min_level)


AMG_CASES = [(2, 2, 2, [(0.6, 0.0, None), (0.4, 0.1, (-0.1, -0.2))], 1), (2, 3, 1, [(0.9, 0.0, None), (0.5, 0.0, None)], 1),
             (3, 2, 1, [(0.9, 0.0, None), (0.5, 0.0, (-0.1, -0.2, -0.3))], 1)]


@pytest.mark.parametrize("dim,p,base,steps,min_level", AMG_CASES)
def test_adaptive_multigrid_building_blocks(ctx, dim, p, base, steps, min_level):
    """mfg_amg_* (csrc/multigrid.cu) piece by piece against oracle/adaptive_mg.py: level operators with boundary + refinement-edge
    constraints (laplace_operator_gpu.h:153-186), vmult_interface_down / up (:306-352), prolongate / restrict_and_add on the
    adaptive blocks (mg_transfer_matrix_free_gpu.cu:592-654), copy_to_mg / copy_from_mg (.cu:688-757), Chebyshev eigenvalues"""
    import dealii_cuda_b200 as mf
    am, mg, o, omg = _adaptive_mg_case(ctx, dim, p, base, steps, min_level)
    assert any(mg.n_edge[l] for l in mg.levels)
    V = lambda a: mf.GpuVector.from_numpy(ctx, np.ascontiguousarray(a, dtype=np.float64))
    for l in mg.levels:
        lm = omg.levels[l]
        assert mg.n_dofs[l] == lm.n_dofs and mg.n_edge[l] == lm.edge.size
        u = sm64(20 + l, lm.n_dofs)
        src, dst = V(u), mf.GpuVector(ctx, lm.n_dofs)
        mg.ops[l].vmult(dst, src)
        assert rel_err(dst.toVector(), lm.vmult(u)) <= 1e-12
        mg.vmult_interface_down(l, dst, src)
        want = lm.vmult_interface_down(u)
        assert np.linalg.norm(dst.toVector() - want) <= 1e-12 * max(np.linalg.norm(want), 1e-300) and (np.linalg.norm(want) > 0) == (lm.edge.size > 0)
        mg.vmult_interface_up(l, dst, src)
        want = lm.vmult_interface_up(u)
        assert np.linalg.norm(dst.toVector() - want) <= 1e-12 * max(np.linalg.norm(want), 1e-300)
        # copies between the active mesh and the level
        ua = sm64(40 + l, o.n_dofs)
        vl = mf.GpuVector(ctx, lm.n_dofs); vl.fill(3.0)
        mg.copy_to_level(l, vl, V(ua))
        assert np.array_equal(vl.toVector(), omg.copy_to_mg(ua)[l])
        va = mf.GpuVector(ctx, o.n_dofs); va.fill(0.0)
        mg.copy_from_level(l, va, src)
        g, lv = omg.copy[l]
        want = np.zeros(o.n_dofs); want[g] = u[lv]
        assert np.array_equal(va.toVector(), want)
        if l > mg.levels[0]:
            lc = omg.levels[l - 1]
            uc = sm64(60 + l, lc.n_dofs)
            fine = mf.GpuVector(ctx, lm.n_dofs); fine.fill(7.0)
            mg.prolongate(l, fine, V(uc))
            assert rel_err(fine.toVector(), omg.P[l] @ uc) <= 1e-13
            d0 = sm64(80 + l, lc.n_dofs)
            coarse = V(d0)
            mg.restrict_and_add(l, coarse, src)
            assert rel_err(coarse.toVector(), d0 + omg.P[l].T @ u) <= 1e-13
            assert abs(mg.lambda_max[l] - omg.smoothers[l][0]) <= 1e-8 * omg.smoothers[l][0]


@pytest.mark.parametrize("dim,p,base,steps,min_level", AMG_CASES)
def test_adaptive_multigrid_vcycle_and_cg(ctx, dim, p, base, steps, min_level):
    """PreconditionMG::vmult (Multigrid::level_v_step with the edge matrices, poisson_mg.cu:365-375) and the V-cycle-preconditioned
    SolverCG on the active operator (:504-518) against the numpy oracle: the same V-cycle output, the same iteration count"""
    import dealii_cuda_b200 as mf
    from oracle.adaptive_mg import cg_preconditioned
    am, mg, o, omg = _adaptive_mg_case(ctx, dim, p, base, steps, min_level)
    r = sm64(7, o.n_dofs); r[o.constrained] = 0.0
    src, dst = mf.GpuVector.from_numpy(ctx, r), mf.GpuVector(ctx, o.n_dofs)
    mg.vmult(dst, src)
    want = omg.vmult(r)
    assert np.linalg.norm(dst.toVector() - want) <= 1e-7 * np.linalg.norm(want)      # (coarse solve: CG to 1e-10 vs a direct solve)
    ue = sm64(5, o.n_dofs); ue[o.constrained] = 0.0
    b = o.vmult(ue)
    tol = 1e-10 * np.linalg.norm(b)
    _, it_ref, hist_ref = cg_preconditioned(o.vmult, omg.vmult, b, tol)
    vb, vx = mf.GpuVector.from_numpy(ctx, b), mf.GpuVector(ctx, o.n_dofs)
    it, res, hist = mg.solve_cg(vx, vb, tol, 100, history=True)
    assert abs(it - it_ref) <= 1 and it <= 12, (it, it_ref)
    k = min(it, it_ref)
    assert np.allclose(hist[:k // 2 + 1], hist_ref[:k // 2 + 1], rtol=1e-5)
    assert np.linalg.norm(vx.toVector() - ue) <= 1e-8 * np.linalg.norm(ue)
    assert mg.coarse_iterations > 0


def test_adaptive_multigrid_through_the_cxx_facade():
    """examples/bmop.cc -DADAPTIVE_GRID, mode `mg`: AdaptiveMesh<3>(limit_level_difference_at_vertices) + pseudo_adaptive_refinement +
    AdaptiveMultigrid<3,double>::solve_cg through the header-only facade (poisson_mg.cu:430-552 on the adaptive grid)"""
    import os
    import re
    import subprocess
    exe = os.path.join(EXAMPLES_BUILD, "bmop_adaptive")
    out = subprocess.run([exe, "4", "4", "mg"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    m = re.search(r"(\d+) iterations.*error ([-0-9.e+]+)", out.stdout)
    assert m, out.stdout
    assert int(m.group(1)) <= 20 and float(m.group(2)) <= 1e-7, out.stdout


# ---------------------------------------------------------------------------------------------------------------------------------
# multigrid over the box partition (dealii_cuda_b200/partitioned_mg.py): all boxes in this process (LocalWorldLevel: the exchange is
# staged through the host), the same algorithm a torchrun job runs with one box per rank

def _partitioned_mg(ctx, world, dim, p, r, strong=False, collapse=True):
    from dealii_cuda_b200.partitioned_mg import LocalWorldLevel, PartitionedMultigrid
    return PartitionedMultigrid(lambda l: LocalWorldLevel(ctx, world, dim, p, l, np.float64, strong), 1, r, collapse_coarse=collapse)


@pytest.mark.parametrize("dim,p,r", [(2, 2, 3), (3, 2, 2)])
def test_partitioned_multigrid_with_one_box_is_the_library_vcycle(ctx, dim, p, r):
    """world = 1: the field algorithm of partitioned_mg.py (Chebyshev with the Lanczos estimate, level_v_step, coarse CG) gives the
    V-cycle of the library's GeometricMultigrid (mfg_mg_*, hardware-verified) -- same eigenvalue estimates, same result"""
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.multigrid import GeometricMultigrid
    pm = _partitioned_mg(ctx, 1, dim, p, r)
    gm = GeometricMultigrid(ctx, dim, p, 1, r)
    for l in range(2, r + 1):
        assert abs(pm.smoothers[l].lambda_max - gm.lambda_max[l]) <= 1e-10 * gm.lambda_max[l]
    n = pm.finest.parts[0].n
    src = sm64(3, n)
    src[pm.finest.parts[0].mesh.constrained_dofs()] = 0.0
    a, b1, b2 = mf.GpuVector.from_numpy(ctx, src), mf.GpuVector(ctx, n), mf.GpuVector(ctx, n)
    pm.vmult([b1], [a])
    gm.vmult(b2, a)
    want = b2.toVector()
    assert np.linalg.norm(b1.toVector() - want) <= 1e-8 * np.linalg.norm(want)


@pytest.mark.parametrize("world,dim,p,r,strong,collapse", [(2, 2, 2, 3, False, True), (2, 2, 2, 3, False, False), (4, 2, 3, 2, False, True),
                                                           (2, 3, 2, 2, False, True), (8, 3, 2, 1, False, False), (8, 3, 2, 2, False, True),
                                                           (4, 2, 2, 3, True, True), (4, 3, 4, 2, True, True)])
def test_partitioned_multigrid_solves_the_global_problem(ctx, world, dim, p, r, strong, collapse):
    """several boxes: MG-preconditioned CG over the partition solves A u = b of the GLOBAL mesh (oracle operator on the global box) in a
    handful of iterations, the V-cycle is symmetric in the global inner product, replicas of interface DoFs stay equal; the coarse
    problem on one replicated global mesh (collapse) or by CG over the partition"""
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.partition import box_for_rank
    from test_partition import global_box, local_to_global_map
    pm = _partitioned_mg(ctx, world, dim, p, r, strong, collapse)
    L = pm.finest
    gbox, _ = global_box(world, dim, r, strong=strong)
    og = OracleMesh(dim, p, box=gbox)
    assert og.n_dofs == L.n_global
    maps = []
    for rank, part in enumerate(L.parts):
        box, me, _ = box_for_rank(rank, world, dim, r, strong=strong)
        maps.append(local_to_global_map(OracleMesh(dim, p, box=box), og, me, p, r, dim, world, strong))
    u = sm64(11, og.n_dofs)
    u[np.asarray(og.constrained)] = 0.0
    b_g = og.vmult(u)
    field = lambda g: [mf.GpuVector.from_numpy(ctx, np.ascontiguousarray(g[m])) for m in maps]
    b, x = field(b_g), L.new_field()
    for v in x:
        v.fill(0.0)
    # the partitioned operator is the global one
    t = L.new_field()
    L.vmult(t, field(u))
    for v, m in zip(t, maps):
        assert np.linalg.norm(v.toVector() - b_g[m]) <= 1e-12 * np.linalg.norm(b_g)
    assert abs(L.dot(b, b) - b_g @ b_g) <= 1e-12 * (b_g @ b_g)
    hist = []
    its, res = pm.solve_cg(x, b, 1e-10 * np.sqrt(L.dot(b, b)), 50, history=hist)
    assert its <= 12, (its, hist)
    got = np.full(og.n_dofs, np.nan)
    for v, m in zip(x, maps):
        part = v.toVector()
        seen = ~np.isnan(got[m])
        assert np.allclose(got[m][seen], part[seen], rtol=0, atol=1e-12 * np.abs(u).max())      # replicas agree
        got[m] = part
    assert np.linalg.norm(got - u) <= 1e-8 * np.linalg.norm(u)
    # symmetry of the preconditioner: <M r1, r2> = <r1, M r2>
    r1_g, r2_g = sm64(12, og.n_dofs), sm64(13, og.n_dofs)
    r1_g[np.asarray(og.constrained)] = 0.0; r2_g[np.asarray(og.constrained)] = 0.0
    r1, r2, m1, m2 = field(r1_g), field(r2_g), L.new_field(), L.new_field()
    pm.vmult(m1, r1); pm.vmult(m2, r2)
    s12, s21 = L.dot(m1, r2), L.dot(r1, m2)
    assert abs(s12 - s21) <= 1e-7 * abs(s12)


@pytest.mark.parametrize("world,p,r,mode", [(2, 2, 3, "weak"), (8, 2, 2, "weak"), (4, 2, 3, "strong")])
def test_partitioned_multigrid_on_real_ranks(world, p, r, mode):
    """the same multigrid with ONE box per rank (torchrun, NCCL / NVLink exchange, all-reduced dots): tests/multirank_mg_worker.py"""
    import subprocess
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    port = 29900 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "tests", "multirank_mg_worker.py"), str(p), str(r), mode]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=root)
    assert out.returncode == 0 and out.stdout.count("MULTIRANK_MG_OK") == world, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.parametrize("args", [("2", "3"), ("8", "2"), ("4", "3", "strong")])
def test_partitioned_multigrid_through_the_cxx_facade(args):
    """examples/partitioned_mg.cc: the same multigrid as C++ host code on the facade (include/dealii_cuda_b200/partitioned_mg.h:
    LocalWorldLevel, PartitionedChebyshev, PartitionedMultigrid), 3D Q4, all boxes on this device"""
    import re
    import subprocess
    exe = os.path.join(EXAMPLES_BUILD, "partitioned_mg")
    assert os.path.exists(exe), "examples/_build/partitioned_mg is missing: run __graft_entry__.build()"
    out = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"(\d+) iterations.*error ([-0-9.e+]+)", out.stdout)
    assert m and int(m.group(1)) <= 10 and float(m.group(2)) <= 1e-8, out.stdout
