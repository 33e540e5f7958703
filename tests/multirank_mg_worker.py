"""Worker of the late GPU test test_partitioned_multigrid_on_real_ranks: one process per GPU (torchrun, NCCL).  Multigrid over the box
partition (dealii_cuda_b200/partitioned_mg.py) with one box per rank -- DistributedLevel: the operator's NVLink / NCCL exchange,
dots all-reduced -- solves A u = b of the GLOBAL mesh; the solution is compared with the u the right-hand side was made from (global
oracle operator).  Prints one line 'MULTIRANK_MG_OK ...' per rank on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.partition import box_for_rank
    from dealii_cuda_b200.partitioned_mg import DistributedLevel, PartitionedMultigrid
    from oracle.oracle import OracleMesh, sm64  # checker
    from test_partition import global_box, local_to_global_map
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dim, p, r = 3, int(sys.argv[1]), int(sys.argv[2])
    strong = len(sys.argv) > 3 and sys.argv[3] == "strong"
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    main_stream = torch.cuda.Stream()
    torch.cuda.set_stream(main_stream)
    ctx = mf.Context(local_rank, main_stream.cuda_stream)
    pm = PartitionedMultigrid(lambda l: DistributedLevel(ctx, rank, world, dim, p, l, np.float64, strong), 1, r)
    L = pm.finest
    gbox, _ = global_box(world, dim, r, strong=strong)
    og = OracleMesh(dim, p, box=gbox)
    box, me, _ = box_for_rank(rank, world, dim, r, strong=strong)
    l2g = local_to_global_map(OracleMesh(dim, p, box=box), og, me, p, r, dim, world, strong)
    u = sm64(11, og.n_dofs)
    u[np.asarray(og.constrained)] = 0.0
    b_g = og.vmult(u)
    b, x = [mf.GpuVector.from_numpy(ctx, np.ascontiguousarray(b_g[l2g]))], L.new_field()
    x[0].fill(0.0)
    bb = L.dot(b, b)
    assert abs(bb - b_g @ b_g) <= 1e-12 * (b_g @ b_g), (bb, b_g @ b_g)
    hist = []
    its, res = pm.solve_cg(x, b, 1e-10 * np.sqrt(bb), 50, history=hist)
    torch.cuda.synchronize()
    got = x[0].toVector()
    err = np.linalg.norm(got - u[l2g]) / np.linalg.norm(u[l2g])
    assert its <= 12 and err <= 1e-8, (rank, its, err, hist)
    print("MULTIRANK_MG_OK rank %d/%d: Q%d r=%d %s, %d global DoFs, %d MG-CG iterations, rel error %.2e"
          % (rank, world, p, r, "strong" if strong else "weak", L.n_global, its, err), flush=True)
    sys.stdout.flush()
    # (as tests/multirank_worker.py: the ranks leave without ncclCommDestroy, which can hang on this stack; a last barrier, itself
    # under a timer, keeps a fast rank from leaving while a peer still waits for its pushes)
    import threading
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        os._exit(0)


if __name__ == "__main__":
    main()
