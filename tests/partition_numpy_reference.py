"""TEST INFRASTRUCTURE: the numpy statement of the box partition and the exchange plan that the library's C++ host code
(csrc/partition.cu, mfg_partition_*) is compared with, array by array, in tests/test_partition.py.

Box partition of the uniform mesh over the GPUs of one node and the interface-DoF exchange plan.

New capability (the reference is single-GPU: GpuVector::compress is a no-op and
locally_owned_elements() is the complete index set, gpu_vec.h:174-175).

Layout: the global mesh is a box of 2^r-cell cubes, one cube per rank, ranks ordered x fastest:
1 -> 1x1x1, 2 -> 1x1x2 (split in z), 4 -> 1x2x2, 8 -> 2x2x2 (the octants of one cube = contiguous
Morton ranges of the refine_global(r+1) mesh).  Every rank stores all DoFs its cells touch; DoFs on
partition interfaces are replicated.  Pure numpy: no CUDA needed (tested with gloo on the CPU).
"""
import itertools

import numpy as np

GRIDS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}


def rank_coords(rank, world, dim=3):
    g = GRIDS[world]
    if dim == 2:
        g = {1: (1, 1, 1), 2: (1, 2, 1), 4: (2, 2, 1)}[world]
    return (rank % g[0], (rank // g[0]) % g[1], rank // (g[0] * g[1])), g


def local_log2(world, dim, r, strong=False):
    """log2 of the cells per direction of one rank's box: weak scaling = a 2^r cube per rank (the domain grows with the
    ranks), strong scaling = the refine_global(r) cube [left,right]^dim cut into the rank grid"""
    _, g = rank_coords(0, world, dim)
    lg = [r - (int(np.log2(g[d])) if strong else 0) for d in range(dim)]
    assert min(lg) >= 0, "more ranks than cells in a direction"
    return lg


def box_for_rank(rank, world, dim, r, left=-1.0, right=1.0, strong=False):
    """Local box descriptor (mfg_box_desc) of a rank: weak scaling: a 2^r cube of cells of edge h = (right-left)/2^r;
    strong scaling: this rank's part of the refine_global(r) mesh of [left,right]^dim."""
    (ix, iy, iz), g = rank_coords(rank, world, dim)
    me = (ix, iy, iz)
    faces = 0
    for d in range(dim):
        if me[d] == 0:
            faces |= 1 << (2 * d)
        if me[d] == g[d] - 1:
            faces |= 1 << (2 * d + 1)
    h = (right - left) / (1 << r)
    lg = local_log2(world, dim, r, strong)
    return dict(log2_cells=lg + [0] * (3 - dim), origin=[left + me[d] * h * (1 << lg[d]) for d in range(dim)] + [0.0] * (3 - dim),
                h=h, dirichlet_faces=faces), me, g


def global_n_dofs(world, dim, degree, r, strong=False):
    _, g = rank_coords(0, world, dim)
    lg = local_log2(world, dim, r, strong)
    n = 1
    for d in range(dim):
        n *= degree * (1 << lg[d]) * g[d] + 1
    return n


class ExchangePlan:
    """Index lists of one rank.  neighbors: ascending ranks sharing DoFs; send/recv counts are symmetric."""

    def __init__(self, rank, world, lists, n_local, replicated=None):
        self.rank, self.world, self.n_local = rank, world, n_local
        self.neighbors = sorted(lists)
        self.lists = {q: np.ascontiguousarray(lists[q], dtype=np.uint32) for q in self.neighbors}
        self.splits = [int(self.lists[q].size) if q in self.lists else 0 for q in range(world)]
        self.pack_idx = (np.concatenate([self.lists[q] for q in self.neighbors]) if self.neighbors else np.zeros(0, np.uint32)).astype(np.uint32)
        self.n_send = int(self.pack_idx.size)
        recv_off, o = {}, 0
        for q in self.neighbors:
            recv_off[q] = o
            o += self.lists[q].size
        self.recv_off = {int(q): int(v) for q, v in recv_off.items()}  # where neighbour q's block starts in this rank's receive buffer
        # CSR of contributions per shared DoF, ascending rank order, -1 = own partial sum
        contrib = {}
        for q in self.neighbors:
            for pos, d in enumerate(self.lists[q].tolist()):
                contrib.setdefault(d, []).append((q, recv_off[q] + pos))
        self.shared_dofs = np.array(sorted(contrib), dtype=np.uint32)
        offsets, slots = [0], []
        owned = np.ones(n_local, dtype=np.uint8)
        for d in self.shared_dofs.tolist():
            items = sorted(contrib[d] + [(rank, -1)])
            slots.extend(s for _, s in items)
            offsets.append(len(slots))
            if items[0][0] != rank:
                owned[d] = 0  # owner = lowest rank touching the DoF
        # constrained interface DoFs are replicated too but take no part in the exchange
        for q, dofs in (replicated or {}).items():
            if q < rank:
                owned[np.asarray(dofs, dtype=np.int64)] = 0
        self.offsets = np.array(offsets, dtype=np.uint32)
        self.slots = np.array(slots, dtype=np.int32)
        self.owned_mask = owned


def build_exchange_plan(rank, world, dim, degree, r, lattice_to_dof, n_local, strong=False):
    """lattice_to_dof: callable mapping an (m,3) uint32 array of LOCAL lattice points (0..p*2^r_d per direction)
    to local DoF indices (HyperCubeMesh.lattice_to_dof on the GPU; an oracle-based map in the CPU tests)."""
    me, g = rank_coords(rank, world, dim)
    lg = local_log2(world, dim, r, strong)
    Md = [degree * (1 << lg[d]) for d in range(dim)]  # last lattice index per direction of the local box
    lists, replicated = {}, {}
    for delta in itertools.product((-1, 0, 1), repeat=dim):
        if not any(delta):
            continue
        nb = [me[d] + delta[d] for d in range(dim)]
        if any(nb[d] < 0 or nb[d] >= g[d] for d in range(dim)):
            continue
        nb_rank = nb[0] + g[0] * ((nb[1] if dim > 1 else 0) + g[1] * (nb[2] if dim > 2 else 0))
        def points(drop_dirichlet):
            ranges = []
            for d in range(dim):
                M = Md[d]
                if delta[d] == 1:
                    ranges.append(np.array([M]))
                elif delta[d] == -1:
                    ranges.append(np.array([0]))
                else:
                    xs = np.arange(M + 1)
                    # points on the global (Dirichlet) boundary are constrained on every replica and carry
                    # dst[c] = src[c], which must not be summed across ranks
                    lo = 1 if (drop_dirichlet and me[d] == 0) else 0
                    hi = M if (drop_dirichlet and me[d] == g[d] - 1) else M + 1
                    ranges.append(xs[lo:hi])
            if any(rg.size == 0 for rg in ranges):
                return None
            while len(ranges) < 3:
                ranges.append(np.array([0]))
            # lexicographic (x fastest) order of the shared points: identical on both sides
            Z, Y, X = np.meshgrid(ranges[2], ranges[1], ranges[0], indexing="ij")
            return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.uint32)

        pts_all = points(False)
        replicated[nb_rank] = lattice_to_dof(pts_all)
        pts = points(True)
        if pts is not None:
            lists[nb_rank] = lattice_to_dof(pts)
    return ExchangePlan(rank, world, lists, n_local, replicated)
