"""Multi-GPU partition logic.
CPU part (-m "not gpu"): world_size-2 run over gloo -- exchange plan, ordered accumulation, owned-DoF dot product --
with the CPU oracle standing in for the local cell loops.  GPU part: 2/4/8 partitions emulated in ONE process on
one GPU (no rank waits on another: B200_PROFILING.md), CUDA pack / accumulate kernels, against the global oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.oracle import OracleMesh, sm64  # noqa: E402


def global_box(world, dim, r, left=-1.0, right=1.0, strong=False):
    from dealii_cuda_b200.partition import rank_coords
    _, g = rank_coords(0, world, dim)
    lg = [r + (0 if strong else int(np.log2(g[d]))) for d in range(dim)]
    return dict(log2_cells=lg, origin=[left] * dim, h=(right - left) / (1 << r), dirichlet_faces=0x3f), g


def local_to_global_map(olocal, oglobal, me, p, r, dim, world=None, strong=False):
    """global DoF index of every local DoF, via lattice coordinates"""
    from dealii_cuda_b200.partition import local_log2
    lg = local_log2(world, dim, r, strong) if world is not None else [r] * dim
    lat = olocal.dof_lattice.astype(np.int64) + np.array([me[d] * p * (1 << lg[d]) if d < dim else 0 for d in range(3)])
    glat = oglobal.dof_lattice.astype(np.int64)
    dims = glat.max(axis=0) + 1
    key = lambda a: a[:, 0] + dims[0] * (a[:, 1] + dims[1] * a[:, 2])
    inv = np.full(int(np.prod(dims)), -1, dtype=np.int64)
    inv[key(glat)] = np.arange(oglobal.n_dofs)
    out = inv[key(lat)]
    assert (out >= 0).all()
    return out


def oracle_lattice_to_dof(olocal):
    lat = olocal.dof_lattice.astype(np.int64)
    dims = lat.max(axis=0) + 1
    inv = np.full(int(np.prod(dims)), -1, dtype=np.int64)
    inv[lat[:, 0] + dims[0] * (lat[:, 1] + dims[1] * lat[:, 2])] = np.arange(olocal.n_dofs)

    def f(pts):
        pts = np.asarray(pts, dtype=np.int64).reshape(-1, 3)
        return inv[pts[:, 0] + dims[0] * (pts[:, 1] + dims[1] * pts[:, 2])].astype(np.uint32)
    return f


def numpy_accumulate(plan, vec, recv):
    out = vec.copy()
    for u, d in enumerate(plan.shared_dofs):
        acc = 0.0
        for j in range(plan.offsets[u], plan.offsets[u + 1]):
            s = plan.slots[j]
            acc += vec[d] if s < 0 else recv[s]
        out[d] = acc
    return out


def _gloo_worker(rank, world, port, dim, p, r, q, strong=False):
    import torch
    import torch.distributed as dist
    from dealii_cuda_b200.partition import box_for_rank, build_exchange_plan, global_n_dofs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gbox, _ = global_box(world, dim, r, strong=strong)
        og = OracleMesh(dim, p, box=gbox)
        assert og.n_dofs == global_n_dofs(world, dim, p, r, strong)
        box, me, g = box_for_rank(rank, world, dim, r, strong=strong)
        ol = OracleMesh(dim, p, box=box)
        l2gmap = local_to_global_map(ol, og, me, p, r, dim, world, strong)
        plan = build_exchange_plan(rank, world, dim, p, r, oracle_lattice_to_dof(ol), ol.n_dofs, strong)
        u_g = sm64(11, og.n_dofs)
        want = og.vmult(u_g)
        # local cell loop (oracle as the stand-in for the CUDA kernel), then the exchange
        part = ol.vmult(u_g[l2gmap])
        send = torch.from_numpy(part[plan.pack_idx].copy())
        recv = torch.empty_like(send)
        reqs, so, ro = [], 0, 0
        for qn in plan.neighbors:
            nq = plan.lists[qn].size
            reqs.append(dist.isend(send[so:so + nq], qn)); so += nq
        for qn in plan.neighbors:
            nq = plan.lists[qn].size
            reqs.append(dist.irecv(recv[ro:ro + nq], qn)); ro += nq
        for rq in reqs:
            rq.wait()
        got = numpy_accumulate(plan, part, recv.numpy())
        err = np.linalg.norm(got - want[l2gmap]) / np.linalg.norm(want[l2gmap])
        # replicas of interface DoFs are bit-identical on all ranks
        full = np.full(og.n_dofs, np.nan)
        full[l2gmap] = got
        gathered = [None] * world
        dist.all_gather_object(gathered, full)
        consistent = True
        for other in gathered:
            both = ~np.isnan(other) & ~np.isnan(full)
            consistent &= bool(np.array_equal(other[both], full[both]))
        # owned-DoF dot product
        v_g = sm64(12, og.n_dofs)
        loc = float(np.dot((u_g[l2gmap] * v_g[l2gmap])[plan.owned_mask.astype(bool)], np.ones(int(plan.owned_mask.sum()))))
        t = torch.tensor([loc], dtype=torch.float64)
        dist.all_reduce(t)
        dot_err = abs(float(t) - float(np.dot(u_g, v_g))) / abs(float(np.dot(u_g, v_g)))
        owned_total = torch.tensor([int(plan.owned_mask.sum())]); dist.all_reduce(owned_total)
        q.put((rank, err, consistent, dot_err, int(owned_total), og.n_dofs, plan.n_send))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,dim,p,r,strong", [(2, 3, 2, 1, False), (2, 3, 4, 1, False), (2, 2, 3, 2, False), (2, 3, 3, 2, True)])
def test_partition_exchange_gloo(world, dim, p, r, strong):
    """weak scaling (one 2^r cube per rank) and strong scaling (the refine_global(r) cube cut into the rank grid)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(rk, world, port, dim, p, r, q, strong)) for rk in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, err, consistent, dot_err, owned_total, n_global, n_send in res:
        assert err <= 1e-13, (rank, err)
        assert consistent
        assert dot_err <= 1e-14
        assert owned_total == n_global  # every global DoF is owned by exactly one rank
        assert n_send > 0


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_counts_single_process(world):
    """plans of all ranks built in one process: symmetric lists, interface sizes, ownership partition"""
    from dealii_cuda_b200.partition import box_for_rank, build_exchange_plan, global_n_dofs
    dim, p, r = 3, 2, 1
    plans, maps = [], []
    gbox, g = global_box(world, dim, r)
    og = OracleMesh(dim, p, box=gbox)
    for rank in range(world):
        box, me, _ = box_for_rank(rank, world, dim, r)
        ol = OracleMesh(dim, p, box=box)
        plans.append(build_exchange_plan(rank, world, dim, p, r, oracle_lattice_to_dof(ol), ol.n_dofs))
        maps.append(local_to_global_map(ol, og, me, p, r, dim))
    for a in range(world):
        for b in plans[a].neighbors:
            assert a in plans[b].neighbors
            # both sides list the same global DoFs in the same order
            assert np.array_equal(maps[a][plans[a].lists[b]], maps[b][plans[b].lists[a]])
    # receive offsets (what a neighbour needs to store straight into this rank's receive buffer, mfg_exchange_push_stream):
    # the blocks of the neighbours tile [0, n_send) in ascending rank order
    for a in range(world):
        o = 0
        for b in plans[a].neighbors:
            assert plans[a].recv_off[b] == o and plans[b].lists[a].size == plans[a].lists[b].size
            o += plans[a].lists[b].size
        assert o == plans[a].n_send
    owned = np.zeros(og.n_dofs, dtype=int)
    for a in range(world):
        np.add.at(owned, maps[a][plans[a].owned_mask.astype(bool)], 1)
    assert (owned == 1).all() and og.n_dofs == global_n_dofs(world, dim, p, r)


@pytest.mark.gpu
@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("world,p,r,dtype", [(2, 4, 1, np.float64), (4, 2, 2, np.float64), (8, 4, 1, np.float64), (8, 3, 2, np.float32),
                                             (2, 4, 3, np.float64), (4, 5, 2, np.float64)])
def test_partitions_emulated_on_one_gpu(ctx, world, p, r, dtype, split):
    """All ranks' partitions live in this process; the collective is replaced by device-side slicing.
    Exercises mfg_mesh_create_box, the CUDA cell loop per partition, mfg_exchange_pack / _accumulate."""
    import torch
    import dealii_cuda_b200 as mf
    from dealii_cuda_b200.distributed import InterfaceExchange
    from dealii_cuda_b200.partition import box_for_rank, build_exchange_plan
    import ctypes as C
    dim = 3
    tol = 1e-12 if dtype == np.float64 else 1e-5
    gbox, g = global_box(world, dim, r)
    og = OracleMesh(dim, p, box=gbox)
    u_g = sm64(21, og.n_dofs).astype(dtype)
    want = og.vmult(u_g.astype(np.float64))
    parts = []
    for rank in range(world):
        box, me, _ = box_for_rank(rank, world, dim, r)
        mesh = mf.HyperCubeMesh(ctx, dim, p, box=box)
        ol = OracleMesh(dim, p, box=box)
        assert np.array_equal(mesh.loc2glob(), ol.loc2glob) and np.array_equal(mesh.constrained_dofs(), ol.constrained)
        op = mf.LaplaceOperatorGpu(ctx, dtype); op.reinit(mesh)
        plan = build_exchange_plan(rank, world, dim, p, r, mesh.lattice_to_dof, mesh.n_dofs)
        ex = InterfaceExchange(ctx, plan, dtype)
        m = local_to_global_map(ol, og, me, p, r, dim)
        src = mf.GpuVector.from_numpy(ctx, u_g[m]); dst = mf.GpuVector(ctx, mesh.n_dofs, dtype)
        if split:
            # the overlapped form: interface cell groups, pack, then the other groups (must not touch packed DoFs)
            k = op.set_interface_dofs(plan.pack_idx)
            assert (k > 0) == (op.active_variant() in (6, 40, 50) and plan.n_send > 0)
            dst.fill(7.0)
            op.vmult_part_ptr(dst.getData(), src.getData(), 0)
            op.vmult_part_ptr(dst.getData(), src.getData(), 1)
            if k == 0:  # kernel without a work list: all cells are in part 2, nothing to overlap
                op.vmult_part_ptr(dst.getData(), src.getData(), 2)
            mf.check(mf.lib.mfg_exchange_pack(ex.h, C.c_void_p(dst.getData()), C.c_void_p(ex.send.data_ptr())))
            if k > 0:
                op.vmult_part_ptr(dst.getData(), src.getData(), 2)
        else:
            op.vmult(dst, src)
            mf.check(mf.lib.mfg_exchange_pack(ex.h, C.c_void_p(dst.getData()), C.c_void_p(ex.send.data_ptr())))
        parts.append(dict(mesh=mesh, op=op, plan=plan, ex=ex, map=m, dst=dst, src=src))
    ctx.synchronize(); torch.cuda.synchronize()
    # "all_to_all": recv buffer of rank a = concatenation over neighbours b (ascending) of b's block for a
    for a, pa in enumerate(parts):
        chunks = []
        for b in pa["plan"].neighbors:
            pb = parts[b]["plan"]
            off = sum(pb.lists[q].size for q in pb.neighbors if q < a)
            chunks.append(parts[b]["ex"].send[off:off + pb.lists[a].size])
        if chunks:
            pa["ex"].recv[:pa["plan"].n_send] = torch.cat(chunks)
    torch.cuda.synchronize()
    full = {}
    for a, pa in enumerate(parts):
        mf.check(mf.lib.mfg_exchange_accumulate(pa["ex"].h, C.c_void_p(pa["dst"].getData()), C.c_void_p(pa["ex"].recv.data_ptr())))
        got = pa["dst"].toVector()
        assert np.linalg.norm(got - want[pa["map"]]) <= tol * np.linalg.norm(want[pa["map"]])
        full[a] = got
    # replicas bit-identical
    ref = np.full(og.n_dofs, np.nan)
    for a, pa in enumerate(parts):
        seen = ~np.isnan(ref[pa["map"]])
        assert np.array_equal(ref[pa["map"]][seen], full[a].astype(np.float64)[seen])
        ref[pa["map"]] = full[a]


@pytest.mark.parametrize("strong", [False, True])
@pytest.mark.parametrize("dim,world", [(3, 1), (3, 2), (3, 4), (3, 8), (2, 2), (2, 4)])
def test_library_partition_equals_numpy_statement(dim, world, strong):
    """the library's C++ host code (csrc/partition.cu, mfg_partition_*) against the numpy statement of the same partition
    (tests/partition_numpy_reference.py), array by array, for every rank of every grid"""
    import partition_numpy_reference as ref
    from dealii_cuda_b200 import partition as lib
    p, r = 2, 2
    assert lib.global_n_dofs(world, dim, p, r, strong) == ref.global_n_dofs(world, dim, p, r, strong)
    assert lib.local_log2(world, dim, r, strong) == ref.local_log2(world, dim, r, strong)
    for rank in range(world):
        assert lib.rank_coords(rank, world, dim) == ref.rank_coords(rank, world, dim)
        bl, mel, gl = lib.box_for_rank(rank, world, dim, r, -1.0, 1.0, strong)
        br, mer, gr = ref.box_for_rank(rank, world, dim, r, -1.0, 1.0, strong)
        assert bl == br and tuple(mel) == tuple(mer) and tuple(gl) == tuple(gr)
        o = OracleMesh(dim, p, box=bl)
        l2d = oracle_lattice_to_dof(o)
        a = lib.build_exchange_plan(rank, world, dim, p, r, l2d, o.n_dofs, strong)
        b = ref.build_exchange_plan(rank, world, dim, p, r, l2d, o.n_dofs, strong)
        assert a.neighbors == b.neighbors and a.splits == b.splits and a.n_send == b.n_send and a.recv_off == b.recv_off
        for name in ("pack_idx", "shared_dofs", "offsets", "slots", "owned_mask"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
            assert getattr(a, name).dtype == getattr(b, name).dtype, name
        assert all(np.array_equal(a.lists[q], b.lists[q]) for q in a.neighbors)


@pytest.mark.parametrize("world,dim,p,r,strong", [(8, 3, 2, 2, 0), (4, 3, 1, 2, 1), (2, 3, 3, 1, 0), (4, 2, 3, 3, 1), (1, 3, 2, 1, 0)])
def test_cxx_facade_partition_plan(world, dim, p, r, strong):
    """include/dealii_cuda_b200/distributed.h (BoxPartition, ExchangePlan) through the host-only example: the same plans as the Python
    binding for the same lattice -> DoF map, and every global DoF owned exactly once"""
    import re
    import subprocess
    from dealii_cuda_b200 import partition as lib
    exe = os.path.join(ROOT, "examples", "_build", "partition_plan")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "-s", "_build/partition_plan"])
    out = subprocess.run([exe, str(world), str(dim), str(p), str(r), str(strong)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert len(out.stdout.strip().splitlines()) == world
    total_owned = 0
    for rank, line in enumerate(out.stdout.strip().splitlines()):
        f = {k: int(v) for k, v in re.findall(r"(n_local|neighbors|n_send|shared|slots|owned|global) (\d+)", line)}
        lg = lib.local_log2(world, dim, r, bool(strong))
        M = [p * (1 << lg[d]) + 1 if d < dim else 1 for d in range(3)]
        l2d = lambda pts: (pts[:, 0] + M[0] * (pts[:, 1] + M[1] * pts[:, 2])).astype(np.uint32)
        plan = lib.build_exchange_plan(rank, world, dim, p, r, l2d, M[0] * M[1] * M[2], bool(strong))
        assert f["n_local"] == M[0] * M[1] * M[2] and f["neighbors"] == len(plan.neighbors) and f["n_send"] == plan.n_send
        assert f["shared"] == plan.shared_dofs.size and f["slots"] == plan.slots.size and f["owned"] == int(plan.owned_mask.sum())
        assert f["global"] == lib.global_n_dofs(world, dim, p, r, bool(strong))
        total_owned += f["owned"]
    assert total_owned == lib.global_n_dofs(world, dim, p, r, bool(strong))
