import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dealii_cuda_b200 as mf
ctx = mf.Context(0, torch.cuda.current_stream().cuda_stream)
mesh = mf.HyperCubeMesh(ctx, 3, 4, 6)
op = mf.LaplaceOperatorGpu(ctx, np.float64); op.reinit(mesh)
n = mesh.n_dofs
ue = mf.GpuVector.wrap(ctx, torch.rand((n,), dtype=torch.float64, device="cuda"))
b, x = mf.GpuVector(ctx, n), mf.GpuVector(ctx, n)
op.vmult(b, ue); op.compute_diagonal(); ctx.synchronize()
for mi in (int(sys.argv[1]),) * 2:
    x.fill(0.0); ctx.synchronize()
    t0 = time.perf_counter()
    its, res = mf.solver_cg(op, x, b, 1e-30, mi)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print("iters", its, "ms/iter", 1e3 * dt / its)
if len(sys.argv) > 2:
    hs = [torch.full((n,), 0.1, dtype=torch.float64).pin_memory() for _ in range(2)]
    hd = [torch.empty((n,), dtype=torch.float64).pin_memory() for _ in range(2)]
    if sys.argv[2] == "e2e":
        for k in range(6):
            op.vmult_host_async(hd[k % 2].numpy(), hs[k % 2].numpy(), k % 2)
        op.host_sync()
    for rep in range(3):
        x.fill(0.0); ctx.synchronize()
        t0 = time.perf_counter()
        its, res = mf.solver_cg(op, x, b, 1e-30, 300)
        ctx.synchronize()
        print(sys.argv[2], "iters", its, "ms/iter", 1e3 * (time.perf_counter() - t0) / its)
