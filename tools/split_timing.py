import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dealii_cuda_b200 as mf
from dealii_cuda_b200.partition import box_for_rank, build_exchange_plan
main = torch.cuda.Stream(); torch.cuda.set_stream(main)
ctx = mf.Context(0, main.cuda_stream)
box, me, grid = box_for_rank(0, 2, 3, 6, -1.0, 1.0)
mesh = mf.HyperCubeMesh(ctx, 3, 4, box=box)
op = mf.LaplaceOperatorGpu(ctx, np.float64); op.reinit(mesh)
plan = build_exchange_plan(0, 2, 3, 4, 6, mesh.lattice_to_dof, mesh.n_dofs)
n = mesh.n_dofs
a = torch.full((n,), 0.1, dtype=torch.float64, device="cuda"); b = torch.zeros_like(a)
k = op.set_interface_dofs(plan.pack_idx)
print("interface groups", k, "n_send", plan.n_send)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
pa, pb = a.data_ptr(), b.data_ptr()
print("whole      %.1f us" % timeit(lambda: op.vmult_ptr(pa, pb)))
print("part 0     %.1f us" % timeit(lambda: op.vmult_part_ptr(pa, pb, 0)))
print("part 1     %.1f us" % timeit(lambda: op.vmult_part_ptr(pa, pb, 1)))
print("part 2     %.1f us" % timeit(lambda: op.vmult_part_ptr(pa, pb, 2)))
print("parts 0+1+2 %.1f us" % timeit(lambda: (op.vmult_part_ptr(pa, pb, 0), op.vmult_part_ptr(pa, pb, 1), op.vmult_part_ptr(pa, pb, 2))))
print("parts 1+2 %.1f us" % timeit(lambda: (op.vmult_part_ptr(pa, pb, 1), op.vmult_part_ptr(pa, pb, 2))))
