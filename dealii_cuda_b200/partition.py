"""Box partition of the uniform mesh over the GPUs of one node and the interface-DoF exchange plan: binding of the library's
C++ host code (csrc/partition.cu, mfg_partition_*; pure host code, runs without a device -- tested with gloo on the CPU).

New capability (the reference is single-GPU: GpuVector::compress is a no-op and locally_owned_elements() is the complete
index set, gpu_vec.h:174-175).  Layout: ranks ordered x fastest on 1 -> 1x1x1, 2 -> 1x1x2 (split in z),
4 -> 1x2x2, 8 -> 2x2x2.  Every rank stores all DoFs its cells touch; DoFs on partition interfaces are replicated."""
import ctypes as C
import itertools

import numpy as np

from . import _capi
from ._capi import check, lib


def rank_coords(rank, world, dim=3):
    me, g = (C.c_int * 3)(), (C.c_int * 3)()
    check(lib.mfg_partition_rank_coords(int(rank), int(world), int(dim), me, g))
    return tuple(me), tuple(g)


def box_for_rank(rank, world, dim, r, left=-1.0, right=1.0, strong=False):
    """Local box descriptor (mfg_box_desc) of a rank: weak scaling: a 2^r cube of cells of edge h = (right-left)/2^r;
    strong scaling: this rank's part of the refine_global(r) mesh of [left,right]^dim."""
    me, g = rank_coords(rank, world, dim)
    d = _capi.BoxDesc()
    check(lib.mfg_partition_box(int(rank), int(world), int(dim), 1, int(r), float(left), float(right), int(bool(strong)), C.byref(d)))
    return dict(log2_cells=list(d.log2_cells), origin=list(d.origin), h=d.h, dirichlet_faces=d.dirichlet_faces), me, g


def local_log2(world, dim, r, strong=False):
    """log2 of the cells per direction of one rank's box"""
    return box_for_rank(0, world, dim, r, strong=strong)[0]["log2_cells"][:dim]


def global_n_dofs(world, dim, degree, r, strong=False):
    n = C.c_uint64()
    check(lib.mfg_partition_global_n_dofs(int(world), int(dim), int(degree), int(r), int(bool(strong)), C.byref(n)))
    return n.value


class ExchangePlan:
    """Index lists of one rank (mfg_partition_plan).  neighbors: ascending ranks sharing DoFs; send/recv counts are symmetric."""

    def __init__(self, rank, world, lists, n_local, replicated=None):
        self.rank, self.world, self.n_local = rank, world, n_local

        def flatten(d):
            ranks = sorted(d)
            arrs = [np.ascontiguousarray(d[q], dtype=np.uint32).ravel() for q in ranks]
            start = np.concatenate([[0], np.cumsum([a.size for a in arrs])]).astype(np.uintp)
            flat = np.ascontiguousarray(np.concatenate(arrs) if arrs else np.zeros(0, np.uint32), dtype=np.uint32)
            return (C.c_int * max(1, len(ranks)))(*ranks), start, flat, len(ranks)

        lr, ls, lf, nl = flatten(lists)
        rr, rs, rf, nr = flatten(replicated or {})
        h = C.c_void_p()
        u32, szp = C.POINTER(C.c_uint32), C.POINTER(C.c_size_t)
        check(lib.mfg_partition_plan_create(int(rank), int(world), int(n_local), nl, lr, ls.ctypes.data_as(szp), lf.ctypes.data_as(u32),
                                            nr, rr, rs.ctypes.data_as(szp), rf.ctypes.data_as(u32), C.byref(h)))
        try:
            sizes = (C.c_size_t * 5)()
            check(lib.mfg_partition_plan_sizes(h, sizes))
            n_send, n_shared, n_slots, n_nb, _ = list(sizes)
            nb = (C.c_int * max(1, n_nb))()
            splits, recv_off = np.zeros(world, np.uint32), np.zeros(world, np.uint32)
            self.pack_idx, self.shared_dofs = np.zeros(n_send, np.uint32), np.zeros(n_shared, np.uint32)
            self.offsets, self.slots = np.zeros(n_shared + 1, np.uint32), np.zeros(n_slots, np.int32)
            self.owned_mask = np.zeros(n_local, np.uint8)
            check(lib.mfg_partition_plan_get(h, nb, splits.ctypes.data_as(u32), recv_off.ctypes.data_as(u32), self.pack_idx.ctypes.data_as(u32),
                                             self.shared_dofs.ctypes.data_as(u32), self.offsets.ctypes.data_as(u32),
                                             self.slots.ctypes.data_as(C.POINTER(C.c_int32)), self.owned_mask.ctypes.data_as(C.POINTER(C.c_uint8))))
        finally:
            lib.mfg_partition_plan_destroy(h)
        self.neighbors = [int(nb[i]) for i in range(n_nb)]
        self.lists = {q: np.ascontiguousarray(lists[q], dtype=np.uint32) for q in self.neighbors}
        self.splits = [int(v) for v in splits]
        self.n_send = int(n_send)
        self.recv_off = {q: int(recv_off[q]) for q in self.neighbors}  # where neighbour q's block starts in this rank's receive buffer


def interface_points(rank, world, dim, degree, r, delta, drop_dirichlet, strong=False):
    """(neighbour rank or -1, (m,3) uint32 local lattice points shared with the neighbour at grid offset delta)"""
    dl = (C.c_int * 3)(*(list(delta) + [0] * (3 - len(delta))))
    nb, cnt = C.c_int(), C.c_size_t()
    check(lib.mfg_partition_interface_points(int(rank), int(world), int(dim), int(degree), int(r), int(bool(strong)), dl, int(bool(drop_dirichlet)),
                                             C.byref(nb), C.byref(cnt), None))
    pts = np.zeros((cnt.value, 3), dtype=np.uint32)
    if cnt.value:
        check(lib.mfg_partition_interface_points(int(rank), int(world), int(dim), int(degree), int(r), int(bool(strong)), dl, int(bool(drop_dirichlet)),
                                                 C.byref(nb), C.byref(cnt), pts.ctypes.data_as(C.POINTER(C.c_uint32))))
    return nb.value, pts


def build_exchange_plan(rank, world, dim, degree, r, lattice_to_dof, n_local, strong=False):
    """lattice_to_dof: callable mapping an (m,3) uint32 array of LOCAL lattice points (0..p*2^r_d per direction)
    to local DoF indices (HyperCubeMesh.lattice_to_dof on the GPU; an oracle-based map in the CPU tests)."""
    lists, replicated = {}, {}
    for delta in itertools.product((-1, 0, 1), repeat=dim):
        if not any(delta):
            continue
        nb_rank, pts_all = interface_points(rank, world, dim, degree, r, delta, False, strong)
        if nb_rank < 0:
            continue
        replicated[nb_rank] = lattice_to_dof(pts_all)
        _, pts = interface_points(rank, world, dim, degree, r, delta, True, strong)
        if pts.shape[0]:
            lists[nb_rank] = lattice_to_dof(pts)
    return ExchangePlan(rank, world, lists, n_local, replicated)
