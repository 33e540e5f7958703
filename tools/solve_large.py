"""Solver-level numbers of BASELINE.json configs[2] ("poisson": 3D Q4, ~100M DoFs) and the multigrid variant on one B200:
CG with the Jacobi (= Chebyshev degree 0) preconditioner at r=6 and r=7, MG-preconditioned CG at r=6 (and r=7 with --mg7).
Right-hand side b = A u for a seeded random u; stop at |r| <= 1e-10 |b|."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dealii_cuda_b200 as mf
from dealii_cuda_b200.multigrid import GeometricMultigrid, solver_cg_preconditioned

ctx = mf.Context(0, torch.cuda.current_stream().cuda_stream)


def problem(op, n):
    ue = mf.GpuVector.wrap(ctx, torch.rand((n,), dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1)))
    b, x = mf.GpuVector(ctx, n), mf.GpuVector(ctx, n)
    op.vmult(b, ue)
    return ue, b, x


for r in (() if "--mg-only" in sys.argv else (6, 7)):
    mesh = mf.HyperCubeMesh(ctx, 3, 4, r)
    op = mf.LaplaceOperatorGpu(ctx, np.float64); op.reinit(mesh)
    n = mesh.n_dofs
    ue, b, x = problem(op, n)
    op.compute_diagonal()
    mf.solver_cg(op, x, b, 0.0, 3); x.fill(0.0); ctx.synchronize()
    t0 = time.perf_counter()
    its, res = mf.solver_cg(op, x, b, 1e-10 * b.l2_norm(), 20000)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    x.add(-1.0, ue)
    print(json.dumps(dict(solver="cg+jacobi", r=r, n_dofs=n, iterations=its, seconds=dt, ms_per_iteration=1e3 * dt / its,
                          rel_error=x.l2_norm() / ue.l2_norm(), gdofs_iter=n * its / dt / 1e9)), flush=True)
    del op, mesh, ue, b, x
    torch.cuda.empty_cache()

for r in ((6, 7) if "--mg7" in sys.argv else (6,)):
    if "--mg-only" in sys.argv or True:
        pass
    t0 = time.perf_counter()
    mg = GeometricMultigrid(ctx, 3, 4, 1, r)
    ctx.synchronize()
    setup = time.perf_counter() - t0
    op = mg.ops[r]
    n = op.m()
    ue, b, x = problem(op, n)
    ctx.synchronize()
    t0 = time.perf_counter()
    it, res_ = mg.solve_cg(x, b, 1e-10 * b.l2_norm(), 100)   # the library's loop (mfg_mg_solve_cg)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    x.add(-1.0, ue)
    print(json.dumps(dict(solver="cg+gmg(V-cycle, Chebyshev(5))", r=r, n_dofs=n, iterations=it, seconds=dt, setup_seconds=setup,
                          rel_error=x.l2_norm() / ue.l2_norm())), flush=True)
    del mg, op, ue, b, x
    torch.cuda.empty_cache()
