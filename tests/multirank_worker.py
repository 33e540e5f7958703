"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun, NCCL).  The N-rank apply of
DistributedLaplaceOperator -- interface cell groups, NVLink P2P pushes into the neighbours' symmetric-memory buffers (or the
NCCL fallback), ordered accumulate, all overlapped with the interior cell groups -- against the GLOBAL oracle mesh, and the
distributed CG against the solution it was set up for.  Prints one line 'MULTIRANK_OK ...' per rank on success."""
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
    from dealii_cuda_b200 import distributed as mfd
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
    worst = 0.0
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-5)):
        dop = mfd.DistributedLaplaceOperator(ctx, rank, world, dim, p, r, dtype, strong=strong)
        gbox, _ = global_box(world, dim, r, strong=strong)
        og = OracleMesh(dim, p, box=gbox)
        ol = OracleMesh(dim, p, box=mfd.box_for_rank(rank, world, dim, r, strong=strong)[0])
        assert np.array_equal(dop.mesh.loc2glob(), ol.loc2glob), "local DoF map differs from the oracle"
        l2g = local_to_global_map(ol, og, dop.me, p, r, dim, world, strong)
        u_g = sm64(21, og.n_dofs).astype(dtype)
        want = og.vmult(u_g.astype(np.float64))[l2g]
        src = mf.GpuVector.from_numpy(ctx, u_g[l2g])
        dst = mf.GpuVector(ctx, ol.n_dofs, dtype)
        for _ in range(3):  # repeated applies reuse the receive buffers: the barriers must keep them apart
            dst.fill(7.0)
            dop.vmult(dst, src)
        torch.cuda.synchronize()
        got = dst.toVector().astype(np.float64)
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert err <= tol, (rank, dtype, err)
        worst = max(worst, err / tol)
        # replicas of interface DoFs are bit-identical on all ranks
        full = np.full(og.n_dofs, np.nan)
        full[l2g] = got
        gathered = [None] * world
        dist.all_gather_object(gathered, full)
        for other in gathered:
            both = ~np.isnan(other) & ~np.isnan(full)
            assert np.array_equal(other[both], full[both]), "replicas of interface DoFs differ between ranks"
        # global dot product over owned DoFs
        v_g = sm64(22, og.n_dofs).astype(dtype)
        dv = mf.GpuVector.from_numpy(ctx, v_g[l2g])
        ref = float(np.dot(u_g.astype(np.float64), v_g.astype(np.float64)))
        assert abs(dop.dot(src, dv) - ref) <= (1e-13 if dtype == np.float64 else 1e-5) * abs(ref)
        if dtype == np.float64:
            # CG over the partition: b = A u, solve, compare with u
            b = mf.GpuVector.from_numpy(ctx, og.vmult(u_g.astype(np.float64))[l2g])
            x = mf.GpuVector(ctx, ol.n_dofs, dtype)
            x.fill(0.0)
            bn = dop.dot(b, b) ** 0.5
            its, res = mfd.solver_cg_distributed(dop, x, b, 1e-11 * bn, 3000)
            xe = x.toVector()
            e = np.linalg.norm(xe - u_g[l2g]) / np.linalg.norm(u_g[l2g])
            assert 0 < its < 3000 and e <= 1e-8, (its, res, e)
    print("MULTIRANK_OK rank %d of %d: worst error / tolerance %.3f, P2P %s" % (rank, world, worst, dop.exchange.symm is not None), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)  # (see distributed.bench_main: the NCCL teardown can hang on this stack)


if __name__ == "__main__":
    main()
