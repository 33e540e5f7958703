"""Geometric multigrid over the box partition (BASELINE configs[4]: multigrid-preconditioned CG on several GPUs; the reference is
single-GPU, SURVEY 8e -- the algorithm is the one of poisson_mg.cu:430-552 / csrc/multigrid.cu, the data layout the one of
partition.py).

Every level l of the hierarchy is partitioned like the finest one: a rank's box on level l - 1 is its box on level l coarsened once,
so the transfer between two levels is LOCAL to a rank (MGTransferMatrixFreeGpu on the rank's two box meshes) and the only
communication is the interface exchange the operator already has:
  * level operator  = local cell loop + exchange_add (DistributedLaplaceOperator);
  * smoother        = Chebyshev on Dinv A with the exchanged diagonal, eigenvalue estimate by CG / Lanczos with owned-DoF dots;
  * restriction     = scale the fine residual by 1 / (number of ranks holding the DoF), restrict locally, exchange_add the coarse
                      result (a fine interface DoF interpolates from coarse DoFs of the same face only, so the local transpose sees its
                      whole row; the scaling counts it once);
  * prolongation    = local (replicas of interface DoFs see the same coarse values);
  * coarse problem  = the global level-1 mesh replicated on every process (CollapsedCoarseSolver: one all-reduce of its right-hand
                      side, the library's CG locally), or unpreconditioned CG over the partition (collapse_coarse = False).
A field is a list of GpuVectors, one per box this process holds: ONE in a torchrun job (DistributedLevel: NCCL / NVLink exchange), ALL
of them in LocalWorldLevel, which stages the exchange through the host and needs neither torch nor a second device -- that form runs
the identical algorithm in tests and in the CPU emulation build."""
import ctypes as C

import numpy as np

from . import GpuVector, HyperCubeMesh, LaplaceOperatorGpu, _capi, check, lib, solver_cg
from .multigrid import MGTransferMatrixFreeGpu
from .partition import box_for_rank, build_exchange_plan, global_n_dofs, rank_coords


class _Part:
    """one box on one level: mesh, operator, exchange plan and the per-DoF weights derived from it"""


class _LevelBase:
    """what the multigrid algorithm needs from a level; subclasses supply parts, exchange_add and the sum over processes"""

    def new_field(self):
        return [GpuVector(self.ctx, p.n, self.dtype) for p in self.parts]

    def _finish_setup(self):
        for p in self.parts:
            p.owned = GpuVector.from_numpy(self.ctx, p.plan.owned_mask.astype(self.dtype))
        ones = self.new_field()
        for v in ones:
            v.fill(1.0)
        self.exchange_add(ones)                       # number of ranks holding the DoF (constrained replicas are not exchanged: 1)
        for v in ones:
            v.invert()
        self.inv_mult = ones
        diag = self.new_field()
        for p, v in zip(self.parts, diag):
            p.op.compute_diagonal()
            v.assign(p.op.get_diagonal_inverse())
            v.invert()                                # local diagonal
        self.exchange_add(diag)                       # interface rows: the sum over the sharing ranks
        for v in diag:
            v.invert()
        self.inv_diag = diag
        self._tmp = self.new_field()

    def dot(self, a, b):
        """global dot product: every DoF counted on the rank that owns it"""
        s = 0.0
        for p, x, y, t in zip(self.parts, a, b, self._tmp):
            t.assign(x)
            t.scale(p.owned)
            s += t.dot(y)
        return self._sum_over_processes(s)

    def vmult(self, dst, src):
        for p, d, s in zip(self.parts, dst, src):
            p.op.vmult(d, s)
        self.exchange_add(dst)


class _HostStagedExchange:
    """mfg_exchange_* of one box with device buffers owned here; the routing between the boxes happens on the host (LocalWorldLevel)"""

    def __init__(self, ctx, plan, dtype):
        code = _capi.F64 if np.dtype(dtype) == np.float64 else _capi.F32
        h, u32 = C.c_void_p(), C.POINTER(C.c_uint32)
        check(lib.mfg_exchange_create(ctx.h, code, plan.pack_idx.ctypes.data_as(u32), plan.n_send, plan.shared_dofs.ctypes.data_as(u32),
                                      plan.shared_dofs.size, plan.offsets.ctypes.data_as(u32),
                                      plan.slots.ctypes.data_as(C.POINTER(C.c_int32)), plan.slots.size, C.byref(h)))
        self.h = h
        self.send, self.recv = GpuVector(ctx, max(1, plan.n_send), dtype), GpuVector(ctx, max(1, plan.n_send), dtype)

    def __del__(self):
        try:
            if self.h:
                lib.mfg_exchange_destroy(self.h)
                self.h = None
        except Exception:
            pass


class LocalWorldLevel(_LevelBase):
    """all `world` boxes of a level in this process, on one device"""

    def __init__(self, ctx, world, dim, degree, level, dtype=np.float64, strong=False, left=-1.0, right=1.0):
        self.ctx, self.world, self.level, self.dtype = ctx, world, level, dtype
        self.dim, self.degree, self.strong, self.left, self.right = dim, degree, strong, left, right
        self.parts = []
        for rank in range(world):
            p = _Part()
            box, p.me, _ = box_for_rank(rank, world, dim, level, left, right, strong)
            p.box = box
            p.mesh = HyperCubeMesh(ctx, dim, degree, box=box)
            p.op = LaplaceOperatorGpu(ctx, dtype)
            p.op.reinit(p.mesh)
            p.n = p.mesh.n_dofs
            p.plan = build_exchange_plan(rank, world, dim, degree, level, p.mesh.lattice_to_dof, p.n, strong)
            p.ex = _HostStagedExchange(ctx, p.plan, dtype)
            self.parts.append(p)
        self.n_global = global_n_dofs(world, dim, degree, level, strong)
        self._finish_setup()

    def _sum_over_processes(self, s):
        return s

    def _sum_array_over_processes(self, a):
        return a

    def exchange_add(self, field):
        if self.world == 1:
            return
        sends = []
        for p, v in zip(self.parts, field):
            if p.plan.n_send:
                check(lib.mfg_exchange_pack(p.ex.h, C.c_void_p(v.getData()), C.c_void_p(p.ex.send.getData())))
            sends.append(p.ex.send.toVector())
        for a, (p, v) in enumerate(zip(self.parts, field)):
            if not p.plan.n_send:
                continue
            # receive buffer of box a = the neighbours' blocks for a, in ascending rank order (what all_to_all_single delivers)
            chunks = []
            for b in p.plan.neighbors:
                pb = self.parts[b].plan
                off = sum(pb.lists[q].size for q in pb.neighbors if q < a)
                chunks.append(sends[b][off:off + pb.lists[a].size])
            recv = np.concatenate(chunks).astype(self.dtype)
            assert recv.size == p.plan.n_send
            check(lib.mfg_vec_from_host(p.ex.recv.h, recv.ctypes.data_as(C.c_void_p), recv.size))
            check(lib.mfg_exchange_accumulate(p.ex.h, C.c_void_p(v.getData()), C.c_void_p(p.ex.recv.getData())))


class DistributedLevel(_LevelBase):
    """this rank's box of a level in a torch.distributed job: DistributedLaplaceOperator (overlapped NVLink exchange inside vmult
    where the box is large enough to have interior groups, NCCL all_to_all for the plain exchange_add)"""

    def __init__(self, ctx, rank, world, dim, degree, level, dtype=np.float64, strong=False, left=-1.0, right=1.0, overlap=True, dop=None):
        from .distributed import DistributedLaplaceOperator
        self.ctx, self.world, self.level, self.dtype = ctx, world, level, dtype
        # (dop: an operator the caller already holds for this level, e.g. the finest one of a benchmark)
        self.dop = dop if dop is not None else DistributedLaplaceOperator(ctx, rank, world, dim, degree, level, dtype, left, right, strong=strong,
                                                                          overlap=overlap)
        self.dim, self.degree, self.strong, self.left, self.right = dim, degree, strong, left, right
        p = _Part()
        p.mesh, p.op, p.n, p.plan = self.dop.mesh, self.dop.op, self.dop.n_local, self.dop.plan
        p.box, p.me, _ = box_for_rank(rank, world, dim, level, left, right, strong)
        self.parts = [p]
        self.n_global = self.dop.n_global
        self._finish_setup()

    def _sum_over_processes(self, s):
        if self.world == 1:
            return s
        import torch
        import torch.distributed as dist
        t = torch.tensor([s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, group=self.dop.exchange.group)
        return float(t.item())

    def _sum_array_over_processes(self, a):
        if self.world == 1:
            return a
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        dist.all_reduce(t, group=self.dop.exchange.group)
        return t.cpu().numpy()

    def dot(self, a, b):
        return self.dop.dot(a[0], b[0])          # one masked-dot kernel (mfg_vec_dot_masked) + all-reduce

    def exchange_add(self, field):
        self.dop.exchange.add_interface_contributions(field[0].getData())

    def vmult(self, dst, src):
        self.dop.vmult(dst[0], src[0])


class PartitionedChebyshev:
    """PreconditionChebyshev on Dinv A of a partitioned level: csrc/multigrid.cu's mfg_cheb (deal.II's procedure) on fields"""

    def __init__(self, level, degree=5, smoothing_range=15.0, eig_iterations=15, eig_cg_residual=1e-2):
        self.level, self.degree = level, degree
        self.d, self.t = level.new_field(), level.new_field()
        self.lambda_max = self.lambda_min = 1.0
        self.eig_iterations = 0
        if eig_iterations > 0:
            self._estimate(eig_iterations, eig_cg_residual)
        beta = 1.2 * self.lambda_max
        alpha = self.lambda_max / smoothing_range if smoothing_range > 1.0 else min(0.9 * self.lambda_max, self.lambda_min)
        self.delta, self.theta = 0.5 * (beta - alpha), 0.5 * (beta + alpha)

    def _estimate(self, n_iterations, eig_cg_residual):
        L = self.level
        g, h, dd = L.new_field(), L.new_field(), L.new_field()
        for k, (p, v) in enumerate(zip(L.parts, g)):
            a = np.full(p.n, -1.0 / np.sqrt(float(L.n_global)), dtype=L.dtype)   # x = 0: g = -b, b = 1 / sqrt(n) ...
            if p.plan.rank == 0 and p.n:
                a[0] = 0.0                                                       # ... with entry 0 = 0 (the domain corner: held by rank 0 only)
            check(lib.mfg_vec_from_host(v.h, a.ctypes.data_as(C.c_void_p), a.size))

        def precondition():
            for x, y, di in zip(h, g, L.inv_diag):
                x.assign(y)
                x.scale(di)
        precondition()
        for x, y in zip(dd, h):
            x.equ(-1.0, y)
        gh, res = L.dot(g, h), np.sqrt(L.dot(g, g))
        alphas, betas = [], []
        it = 1
        while it <= n_iterations and res > eig_cg_residual:
            L.vmult(h, dd)
            alpha = gh / L.dot(dd, h)
            alphas.append(alpha)
            for x, y in zip(g, h):
                x.add(alpha, y)
            res = np.sqrt(L.dot(g, g))
            precondition()
            gh_new = L.dot(g, h)
            beta, gh = gh_new / gh, gh_new
            betas.append(beta)
            for x, y in zip(dd, h):
                x.sadd(beta, -1.0, y)
            it += 1
        self.eig_iterations = len(alphas)
        if not alphas:
            return
        k = len(alphas)
        T = np.zeros((k, k))
        for j in range(k):
            T[j, j] = 1.0 / alphas[j] + (betas[j - 1] / alphas[j - 1] if j else 0.0)
            if j + 1 < k:
                T[j, j + 1] = T[j + 1, j] = np.sqrt(betas[j]) / alphas[j]
        ev = np.linalg.eigvalsh(T)
        self.lambda_max, self.lambda_min = float(ev[-1]), float(ev[0])

    def _update(self, x, b, f1, f2, zero_start, first):
        # t = A x on entry (ignored with zero_start):  r = Dinv (b - t);  d = f1 d + f2 r;  x += d   (ONE kernel per box: cheb_update of
        # csrc/multigrid.cu through mfg_vec_chebyshev_update)
        for xi, bi, ti, di, dinv in zip(x, b, self.t, self.d, self.level.inv_diag):
            check(lib.mfg_vec_chebyshev_update(self.level.ctx.h, xi.h, di.h, ti.h, bi.h, dinv.h, float(f1), float(f2), int(bool(zero_start)), int(bool(first))))

    def apply(self, x, b, zero_start):
        """PreconditionChebyshev::vmult (zero_start) / ::step"""
        rhok, sigma = self.delta / self.theta, self.theta / self.delta
        if not zero_start:
            self.level.vmult(self.t, x)
        self._update(x, b, 0.0, 1.0 / self.theta, zero_start, True)
        for _ in range(self.degree):
            self.level.vmult(self.t, x)
            rhokp = 1.0 / (2.0 * sigma - rhok)
            f1, f2, rhok = rhokp * rhok, 2.0 * rhokp / self.delta, rhokp
            self._update(x, b, f1, f2, False, False)


def field_cg(level, x, b, abs_tol, max_iter, precond=None, history=None):
    """SolverCG on fields (csrc/multigrid.cu cg_preconditioned: same control flow); precond(dst, src) or None.  Returns (iterations, residual)."""
    g, h, d = level.new_field(), level.new_field(), level.new_field()
    level.vmult(g, x)
    for gi, bi in zip(g, b):
        gi.add(-1.0, bi)
    res = np.sqrt(level.dot(g, g))
    if history is not None:
        history.append(res)
    it = 0
    if res > abs_tol:
        def apply_precond():
            if precond is None:
                for hi, gi in zip(h, g):
                    hi.assign(gi)
            else:
                precond(h, g)
        apply_precond()
        for di, hi in zip(d, h):
            di.equ(-1.0, hi)
        gh = level.dot(g, h)
        for it in range(1, max_iter + 1):
            level.vmult(h, d)
            alpha = gh / level.dot(d, h)
            for xi, di in zip(x, d):
                xi.add(alpha, di)
            for gi, hi in zip(g, h):
                gi.add(alpha, hi)
            res = np.sqrt(level.dot(g, g))
            if history is not None:
                history.append(res)
            if res <= abs_tol:
                break
            apply_precond()
            gh_new = level.dot(g, h)
            beta, gh = gh_new / gh, gh_new
            for di, hi in zip(d, h):
                di.sadd(beta, -1.0, hi)
    return it, res


class CollapsedCoarseSolver:
    """The coarsest level on ONE mesh: the boxes' right-hand sides are summed into the vector of the global level-`min_level` mesh
    (every DoF from the rank that owns it, one all-reduce of a few thousand numbers), every process solves that small system with the
    library's CG (mfg_solver_cg, PreconditionIdentity, reduction 1e-10: bmop_mg.cu:75-82) -- identical arithmetic everywhere, so the
    replicas agree -- and reads its box back.  Replaces a few hundred partitioned CG iterations, each of them latency-bound (an
    exchange and two all-reduced dots for a handful of cells per rank), by one collective."""

    def __init__(self, L):
        self.L = L
        dim, deg = L.dim, L.degree
        _, g = rank_coords(0, L.world, dim)
        lg = [L.level + (0 if L.strong else int(round(np.log2(g[d])))) for d in range(dim)] + [0] * (3 - dim)
        box = dict(log2_cells=lg, origin=[L.left] * 3, h=(L.right - L.left) / (1 << L.level), dirichlet_faces=0x3f)
        self.mesh = HyperCubeMesh(L.ctx, dim, deg, box=box)
        assert self.mesh.n_dofs == L.n_global, (self.mesh.n_dofs, L.n_global)
        self.op = LaplaceOperatorGpu(L.ctx, L.dtype)
        self.op.reinit(self.mesh)
        self.maps = []
        for p in L.parts:
            lgl = p.box["log2_cells"]
            npts = [deg * (1 << lgl[d]) + 1 if d < dim else 1 for d in range(3)]
            grid = np.stack(np.meshgrid(*[np.arange(n, dtype=np.int64) for n in npts], indexing="ij"), axis=-1).reshape(-1, 3)
            off = np.array([p.me[d] * deg * (1 << lgl[d]) if d < dim else 0 for d in range(3)], dtype=np.int64)
            loc = p.mesh.lattice_to_dof(grid.astype(np.uint32))
            m = np.full(p.n, -1, dtype=np.int64)
            m[loc] = self.mesh.lattice_to_dof((grid + off).astype(np.uint32))
            assert (m >= 0).all()
            self.maps.append(m)
        self.bg, self.xg = GpuVector(L.ctx, self.mesh.n_dofs, L.dtype), GpuVector(L.ctx, self.mesh.n_dofs, L.dtype)
        self.iterations = 0

    def solve(self, x, b):
        L = self.L
        acc = np.zeros(self.mesh.n_dofs)
        for p, m, v in zip(L.parts, self.maps, b):
            own = p.plan.owned_mask.astype(bool)
            acc[m[own]] += v.toVector().astype(np.float64)[own]
        acc = L._sum_array_over_processes(acc)
        a = np.ascontiguousarray(acc, dtype=L.dtype)
        check(lib.mfg_vec_from_host(self.bg.h, a.ctypes.data_as(C.c_void_p), a.size))
        self.xg.fill(0.0)
        tol = (1e-10 if np.dtype(L.dtype) == np.float64 else 1e-4) * max(float(np.linalg.norm(acc)), 1e-300)
        its, _ = solver_cg(self.op, self.xg, self.bg, tol, 10000, use_jacobi=False)[:2]
        self.iterations += its
        xs = self.xg.toVector()
        for m, v in zip(self.maps, x):
            part = np.ascontiguousarray(xs[m])
            check(lib.mfg_vec_from_host(v.h, part.ctypes.data_as(C.c_void_p), part.size))


class PartitionedMultigrid:
    """V-cycle (Multigrid::level_v_step, csrc/multigrid.cu mfg_mg::cycle) on partitioned levels min_level..max_level and the CG it
    preconditions.  make_level(l) returns the level object (LocalWorldLevel or DistributedLevel)."""

    def __init__(self, make_level, min_level, max_level, smoother_degree=5, smoothing_range=15.0, eig_iterations=15, collapse_coarse=True):
        assert 1 <= min_level <= max_level, "the coarsest partitioned level needs at least 2 cells per direction and box"
        self.min_level, self.max_level = min_level, max_level
        self.levels = {l: make_level(l) for l in range(min_level, max_level + 1)}
        top = self.levels[max_level]
        self.ctx, self.dtype = top.ctx, top.dtype
        self.smoothers = {l: PartitionedChebyshev(self.levels[l], smoother_degree, smoothing_range, eig_iterations) for l in range(min_level + 1, max_level + 1)}
        self.transfers = {}
        for l in range(min_level + 1, max_level + 1):
            ts = []
            for pc, pf in zip(self.levels[l - 1].parts, self.levels[l].parts):
                t = MGTransferMatrixFreeGpu(self.ctx, self.dtype)
                t.build({l - 1: pc.mesh, l: pf.mesh})
                ts.append(t)
            self.transfers[l] = ts
        self.x = {l: L.new_field() for l, L in self.levels.items()}
        self.b = {l: L.new_field() for l, L in self.levels.items()}
        self.t = {l: L.new_field() for l, L in self.levels.items()}
        self.coarse_iterations = 0
        # the coarse problem: on one replicated global mesh (default), or by CG over the partition
        self.coarse = CollapsedCoarseSolver(self.levels[min_level]) if collapse_coarse else None

    @property
    def finest(self):
        return self.levels[self.max_level]

    def restrict_and_add(self, level, dst_coarse, src_fine):
        """dst_coarse += R src_fine with every fine DoF counted once; src_fine is scaled in place"""
        for v, w in zip(src_fine, self.levels[level].inv_mult):
            v.scale(w)
        for t, dc, sf in zip(self.transfers[level], dst_coarse, src_fine):
            t.restrict_and_add(level, dc, sf)
        self.levels[level - 1].exchange_add(dst_coarse)

    def prolongate(self, level, dst_fine, src_coarse):
        for t, df, sc in zip(self.transfers[level], dst_fine, src_coarse):
            t.prolongate(level, df, sc)

    def _cycle(self, l):
        L, x, b, t = self.levels[l], self.x[l], self.b[l], self.t[l]
        if l == self.min_level and self.coarse is not None:
            before = self.coarse.iterations
            self.coarse.solve(x, b)
            self.coarse_iterations += self.coarse.iterations - before
            return
        if l == self.min_level:
            for v in x:
                v.fill(0.0)
            bn = np.sqrt(L.dot(b, b))
            its, _ = field_cg(L, x, b, (1e-10 if np.dtype(self.dtype) == np.float64 else 1e-4) * max(bn, 1e-300), 10000)
            self.coarse_iterations += its
            return
        self.smoothers[l].apply(x, b, True)             # pre-smoothing from a zero guess
        L.vmult(t, x)
        for ti, bi in zip(t, b):
            ti.sadd(-1.0, 1.0, bi)                      # t = b - A x
        for v in self.b[l - 1]:
            v.fill(0.0)
        self.restrict_and_add(l, self.b[l - 1], t)
        self._cycle(l - 1)
        self.prolongate(l, t, self.x[l - 1])
        for xi, ti in zip(x, t):
            xi.add(ti)
        self.smoothers[l].apply(x, b, False)            # post-smoothing

    def vmult(self, dst, src):
        """PreconditionMG::vmult: one V-cycle on the right-hand side src"""
        for bi, si in zip(self.b[self.max_level], src):
            bi.assign(si)
        self._cycle(self.max_level)
        for di, xi in zip(dst, self.x[self.max_level]):
            di.assign(xi)

    def solve_cg(self, x, b, abs_tol, max_iter=1000, history=None):
        return field_cg(self.finest, x, b, abs_tol, max_iter, precond=self.vmult, history=history)
