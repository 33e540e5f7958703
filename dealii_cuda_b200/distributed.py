"""Multi-GPU Laplace operator: one process per GPU (torch.distributed, NCCL), box partition of the mesh,
interface-DoF exchange after the local cell loop.  See partition.py for the layout."""
import ctypes as C
import json
import os
import time

import numpy as np

from . import (Context, GpuVector, HyperCubeMesh, LaplaceOperatorGpu, _capi, check, lib)
from .partition import box_for_rank, build_exchange_plan, global_n_dofs


class InterfaceExchange:
    """mfg_exchange + send/recv buffers + the collective."""

    def __init__(self, ctx, plan, dtype, group=None):
        import torch
        self.ctx, self.plan, self.group = ctx, plan, group
        code = _capi.F64 if np.dtype(dtype) == np.float64 else _capi.F32
        h = C.c_void_p()
        u32 = C.POINTER(C.c_uint32)
        check(lib.mfg_exchange_create(ctx.h, code, plan.pack_idx.ctypes.data_as(u32), plan.n_send, plan.shared_dofs.ctypes.data_as(u32),
                                      plan.shared_dofs.size, plan.offsets.ctypes.data_as(u32),
                                      plan.slots.ctypes.data_as(C.POINTER(C.c_int32)), plan.slots.size, C.byref(h)))
        self.h = h
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        self.send = torch.empty(max(1, plan.n_send), dtype=tdt, device="cuda")
        self.recv = torch.empty(max(1, plan.n_send), dtype=tdt, device="cuda")
        self.owned_mask = torch.from_numpy(plan.owned_mask).cuda()
        # second stream for the exchange when it is overlapped with the interior cells
        # (high priority: its interface cell kernel gets SM slots before the interior kernel launched next to it)
        self.side = torch.cuda.Stream(priority=-1) if plan.world > 1 and plan.n_send else None
        self.ev_ready, self.ev_done = torch.cuda.Event(), torch.cuda.Event()
        self.symm, self.recv_p2p = None, None
        import torch.distributed as dist
        if self.side is not None and os.environ.get("MFG_NO_P2P") is None and dist.is_available() and dist.is_initialized() \
                and dist.get_backend(group) == "nccl":
            self._setup_p2p(tdt)

    def _setup_p2p(self, tdt):
        """Receive buffer in symmetric memory: the neighbours store their partial sums straight into it over NVLink
        (mfg_exchange_push_stream) and a device-side barrier replaces the NCCL collective.  Falls back to NCCL
        (self.symm stays None) where peer mapping is not available."""
        import torch
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            plan = self.plan
            cap = torch.tensor([max(1, plan.n_send)], dtype=torch.int64, device="cuda")
            dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=self.group)
            recv = symm_mem.empty(int(cap.item()), dtype=tdt, device=torch.device("cuda", torch.cuda.current_device()))
            hdl = symm_mem.rendezvous(recv, self.group if self.group is not None else dist.group.WORLD)
            offs = [None] * plan.world
            dist.all_gather_object(offs, plan.recv_off, group=self.group)
            es = recv.element_size()
            ptrs = list(hdl.buffer_ptrs)
            nb = plan.neighbors
            self._peer_dst = (C.c_uint64 * len(nb))(*[ptrs[q] + offs[q][plan.rank] * es for q in nb])
            starts = np.concatenate([[0], np.cumsum([plan.splits[q] for q in nb])]).astype(np.uint32)
            self._chunk_start = (C.c_uint32 * starts.size)(*starts.tolist())
            self._n_chunks = len(nb)
            # (the NCCL path keeps its own receive buffer: a neighbour that is one apply ahead may already push into this one)
            self.recv_p2p, self.symm = recv, hdl
        except Exception as e:  # no peer access / symmetric memory on this system
            self.symm = None
            if self.plan.rank == 0:
                print("symmetric-memory exchange unavailable (%s: %s): NCCL all_to_all" % (type(e).__name__, e), file=__import__("sys").stderr)

    def __del__(self):
        try:
            if self.h:
                lib.mfg_exchange_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def add_interface_contributions(self, vec_ptr):
        """compress(add) + update_ghost_values in one step: every replica of an interface DoF ends with the
        sum of all partial sums, added in ascending rank order."""
        import torch.distributed as dist
        if self.plan.world == 1 or self.plan.n_send == 0:
            return
        check(lib.mfg_exchange_pack(self.h, C.c_void_p(vec_ptr), C.c_void_p(self.send.data_ptr())))
        n = self.plan.n_send
        dist.all_to_all_single(self.recv[:n], self.send[:n], self.plan.splits, self.plan.splits, group=self.group)
        check(lib.mfg_exchange_accumulate(self.h, C.c_void_p(vec_ptr), C.c_void_p(self.recv.data_ptr())))


    def run_overlapped(self, op, dst_ptr, src_ptr):
        """One apply with the exchange hidden behind the interior cells:
        main stream:  zero/constraint pass | interface cell groups | interior cell groups (programmatic dependent launch:
                                                                     its CTAs fill the SMs as the interface CTAs leave) | wait
        side stream:                                               | pack, all_to_all, ordered accumulate |"""
        import torch
        import torch.distributed as dist
        main = torch.cuda.current_stream()
        side = self.side
        op.vmult_part_ptr(dst_ptr, src_ptr, 0)
        op.vmult_part_ptr(dst_ptr, src_ptr, 1)
        self.ev_ready.record(main)
        op.vmult_part_ptr(dst_ptr, src_ptr, 2)   # adds into no exchanged DoF
        side.wait_event(self.ev_ready)
        if self.symm is not None:
            # P2P: stores into the neighbours' receive buffers, barrier (all stores landed), ordered accumulate, barrier
            # (everybody has read its buffer: the next apply may overwrite it).  No SM-hungry collective kernel.
            check(lib.mfg_exchange_push_stream(self.h, C.c_void_p(dst_ptr), self._peer_dst, self._chunk_start, self._n_chunks,
                                               C.c_void_p(side.cuda_stream)))
            with torch.cuda.stream(side):
                self.symm.barrier(0)
            check(lib.mfg_exchange_accumulate_stream(self.h, C.c_void_p(dst_ptr), C.c_void_p(self.recv_p2p.data_ptr()), C.c_void_p(side.cuda_stream)))
            with torch.cuda.stream(side):
                self.symm.barrier(1)
        else:
            check(lib.mfg_exchange_pack_stream(self.h, C.c_void_p(dst_ptr), C.c_void_p(self.send.data_ptr()), C.c_void_p(side.cuda_stream)))
            n = self.plan.n_send
            with torch.cuda.stream(side):
                dist.all_to_all_single(self.recv[:n], self.send[:n], self.plan.splits, self.plan.splits, group=self.group)
            check(lib.mfg_exchange_accumulate_stream(self.h, C.c_void_p(dst_ptr), C.c_void_p(self.recv.data_ptr()), C.c_void_p(side.cuda_stream)))
        self.ev_done.record(side)
        main.wait_event(self.ev_done)


class DistributedLaplaceOperator:
    """LaplaceOperatorGpu over a box partition: vmult = local cell loop + interface exchange."""

    def __init__(self, ctx, rank, world, dim, degree, r, dtype=np.float64, left=-1.0, right=1.0, variant=0, group=None, overlap=True,
                 strong=False):
        self.ctx, self.rank, self.world, self.strong = ctx, rank, world, strong
        box, self.me, self.grid = box_for_rank(rank, world, dim, r, left, right, strong)
        self.mesh = HyperCubeMesh(ctx, dim, degree, box=box)
        self.op = LaplaceOperatorGpu(ctx, dtype)
        self.op.reinit(self.mesh)
        if variant:
            self.op.set_variant(variant)
        self.plan = build_exchange_plan(rank, world, dim, degree, r, self.mesh.lattice_to_dof, self.mesh.n_dofs, strong)
        self.exchange = InterfaceExchange(ctx, self.plan, dtype, group)
        # overlap: cell groups that touch exchanged DoFs run first, the rest while the exchange is in flight
        self.n_iface_groups = self.op.set_interface_dofs(self.plan.pack_idx) if (overlap and world > 1 and self.plan.n_send) else 0
        self.n_local = self.mesh.n_dofs
        self.n_global = global_n_dofs(world, dim, degree, r, strong)

    def vmult_ptr(self, dst_ptr, src_ptr):
        if self.n_iface_groups:
            self.exchange.run_overlapped(self.op, dst_ptr, src_ptr)
        else:
            self.op.vmult_ptr(dst_ptr, src_ptr)
            self.exchange.add_interface_contributions(dst_ptr)

    def vmult(self, dst, src):
        self.vmult_ptr(dst.getData(), src.getData())

    def vmult_graphed(self, dst_ptr, src_ptr):
        """the apply (cell kernels, pushes, barriers, accumulate on two streams) replayed as ONE CUDA graph per (dst, src) pair:
        eager launches from Python are host-bound.  Needs a non-default current stream; falls back to eager launches."""
        import torch
        graphs = self.__dict__.setdefault("_graphs", {})
        g = graphs.get((dst_ptr, src_ptr))
        if g is None:
            self.vmult_ptr(dst_ptr, src_ptr)  # (this call computes the result; the capture below only records)
            if graphs.get("disabled") or os.environ.get("MFG_NO_GRAPH") is not None:
                return
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                    self.vmult_ptr(dst_ptr, src_ptr)
                graphs[(dst_ptr, src_ptr)] = g
            except Exception as e:
                graphs["disabled"] = True
                if self.rank == 0:
                    print("CUDA graph capture of the apply failed (%s): eager launches" % e, file=__import__("sys").stderr)
            return
        g.replay()

    def dot(self, a, b):
        """Global dot product: owned DoFs locally (deterministic two-pass reduction), then all_reduce."""
        import torch
        import torch.distributed as dist
        out = C.c_double()
        check(lib.mfg_vec_dot_masked(a.h, b.h, C.c_void_p(self.exchange.owned_mask.data_ptr()), C.byref(out)))
        if self.world == 1:
            return out.value
        t = torch.tensor([out.value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, group=self.exchange.group)
        return float(t.item())


def solver_cg_distributed(dop, x, b, abs_tol, max_iter=10000, use_jacobi=True, check_every=8):
    """SolverCG over the box partition (poisson.cu:233-260 control flow; the single-GPU form is mfg_solver_cg).  The operator is
    DistributedLaplaceOperator.vmult (replicas of interface DoFs stay bit-identical), the vector kernels are mfg_cgd_*: sums
    over the owned DoFs, scalars in a device block that is all-reduced in stream order (NCCL) -- no rank reads a scalar on the
    host inside the loop.  Every `check_every` iterations all ranks read the (identical) convergence flag; kernels of
    iterations behind the converged one return at once, so x is the iterate of the converged iteration.  The Jacobi
    diagonal is the exchanged sum of the local diagonals.  Returns (iterations, last residual)."""
    import torch
    import torch.distributed as dist
    ctx, n = dop.ctx, dop.n_local
    dt = np.float64 if x.dtype == np.float64 else np.float32
    code = _capi.F64 if dt == np.float64 else _capi.F32
    g, d, h = GpuVector(ctx, n, dt), GpuVector(ctx, n, dt), GpuVector(ctx, n, dt)
    minv = None
    if use_jacobi:
        if getattr(dop, "_inv_diag", None) is None:
            dop.op.compute_diagonal()
            diag = GpuVector(ctx, n, dt)
            diag.assign(dop.op.get_diagonal_inverse())
            diag.invert()                                        # local diagonal (constrained rows: 1)
            dop.exchange.add_interface_contributions(diag.getData())   # interface rows: sum over the sharing ranks
            diag.invert()
            dop._inv_diag = diag
        minv = dop._inv_diag
    world, group = dop.world, dop.exchange.group
    scal = torch.zeros(16, dtype=torch.float64, device="cuda")   # (kept alive by the captured graph)
    sp, own = C.c_void_p(scal.data_ptr()), C.c_void_p(dop.exchange.owned_mask.data_ptr())
    check(lib.mfg_cgd_init(ctx.h, sp))
    vp = lambda v: C.c_void_p(v.getData())

    def allreduce(lo, hi):
        if world > 1:
            dist.all_reduce(scal[lo:hi], group=group)

    # g = A x - b ; iteration 0: z = Minv g, |g|, g.z ; d = -z
    dop.vmult(g, x)
    g.add(-1.0, b)
    check(lib.mfg_cgd_residual(ctx.h, code, vp(g), vp(h), vp(minv) if minv is not None else None, own, n, sp, 1))
    allreduce(1, 3)
    check(lib.mfg_cgd_beta(ctx.h, sp, float(abs_tol), 0))
    check(lib.mfg_cgd_advance(ctx.h, code, vp(x), vp(d), vp(h), n, sp, 0))
    def iteration(it):
        dop.vmult_graphed(h.getData(), d.getData())                              # h = A d
        check(lib.mfg_cgd_dot(ctx.h, code, vp(d), vp(h), own, n, sp))
        allreduce(0, 1)
        check(lib.mfg_cgd_alpha(ctx.h, sp))
        check(lib.mfg_cgd_residual(ctx.h, code, vp(g), vp(h), vp(minv) if minv is not None else None, own, n, sp, 0))
        allreduce(1, 3)
        check(lib.mfg_cgd_beta(ctx.h, sp, float(abs_tol), it))
        check(lib.mfg_cgd_advance(ctx.h, code, vp(x), vp(d), vp(h), n, sp, it))

    # One iteration (apply, vector kernels, the two all-reduces) as ONE CUDA graph: the Python loop with its NCCL calls is
    # host-bound (1.05 ms per iteration at 2 x 17 M DoFs against 0.6 ms of device work).  it = -1: device-side counter.
    graph = None
    it = 0
    if max_iter >= 2 and os.environ.get("MFG_NO_GRAPH") is None and torch.cuda.current_stream() != torch.cuda.default_stream():
        iteration(1)                       # eager: warms every kernel and the apply's own graph
        it = 1
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=torch.cuda.current_stream()):
                dop.vmult_ptr(h.getData(), d.getData())
                check(lib.mfg_cgd_dot(ctx.h, code, vp(d), vp(h), own, n, sp))
                allreduce(0, 1)
                check(lib.mfg_cgd_alpha(ctx.h, sp))
                check(lib.mfg_cgd_residual(ctx.h, code, vp(g), vp(h), vp(minv) if minv is not None else None, own, n, sp, 0))
                allreduce(1, 3)
                check(lib.mfg_cgd_beta(ctx.h, sp, float(abs_tol), -1))
                check(lib.mfg_cgd_advance(ctx.h, code, vp(x), vp(d), vp(h), n, sp, -1))
        except Exception as e:
            graph = None
            if dop.rank == 0:
                print("CUDA graph capture of the CG iteration failed (%s): eager launches" % e, file=__import__("sys").stderr)
    done = False
    state = scal[:8].cpu()
    done = float(state[7]) >= 0
    while it < max_iter and not done:
        for _ in range(min(check_every, max_iter - it)):
            it += 1
            if graph is not None:
                graph.replay()
            else:
                iteration(it)
        state = scal[:8].cpu()                     # the only host read: identical on all ranks
        done = float(state[7]) >= 0
    its = int(state[7]) if done else it
    return its, float(state[3])


def parse_cpulist(text):
    """'0-3,8,10-11' (sysfs cpulist) -> set of CPU numbers"""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank, sysfs="/sys"):
    """Run this rank's host threads on the CPUs of the NUMA node its GPU hangs on, so that the pinned host buffers of the end-to-end
    leg (first touch) are local to the GPU's PCIe root: with one rank per GPU and 2 x 136 MB crossing PCIe per step and rank, buffers
    on the other socket put every copy on the inter-socket link.  Returns the node, or None where the topology is not visible
    (containers without sysfs NUMA information, a single node): nothing is changed then."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")).read())
        if node < 0 or not os.path.isdir(os.path.join(sysfs, "devices/system/node/node1")):
            return None
        cpus = parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read()) & os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def bench_main(args, metric):
    """bench.py --gpus N (N > 1): weak scaling, one 2^r cube of cells per GPU."""
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, b_alg, measured_peaks, cpu_reference_run, PROFILED_TRAFFIC
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL and the symmetric-memory rendezvous print to stdout; the contract is ONE JSON line there: everything else
    # goes to stderr, the line is written to the saved descriptor at the end
    import sys
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank)
    # The exchange runs beside the persistent interior cell kernel, which leaves 4 CTA slots free (MFG_SLAB2_RESERVE):
    # NCCL's send/recv kernel must fit into them, so few and narrow channels (measured on 2 x B200: 0.281 ms per apply
    # against 0.287 with NCCL's defaults, profiles/r01_multigpu_overlap.txt)
    if os.environ.get("MFG_NO_P2P") is not None:  # only the NCCL fallback of the exchange needs this
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
        os.environ.setdefault("NCCL_NTHREADS", "128")
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus
    # a non-default stream, so that a whole apply (cell kernels, pack, NCCL all_to_all on the side stream, accumulate)
    # can be captured into a CUDA graph and replayed without per-step host work
    main = torch.cuda.Stream()
    torch.cuda.set_stream(main)
    ctx = Context(local_rank, main.cuda_stream)
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    s = 8 if args.dtype == "f64" else 4
    strong = getattr(args, "scaling", "weak") == "strong"
    dop = DistributedLaplaceOperator(ctx, rank, world, args.dim, args.degree, args.refine, dtype, variant=args.variant, strong=strong)
    n = dop.n_local
    ta = torch.full((n,), 0.1, dtype=tdtype, device="cuda")
    tb = torch.zeros((n,), dtype=tdtype, device="cuda")
    pa, pb = ta.data_ptr(), tb.data_ptr()

    graphs = {}

    def apply_steps(k):
        nonlocal pa, pb
        for _ in range(k):
            pa, pb = pb, pa
            g = graphs.get((pa, pb))
            if g is not None:
                g.replay()
            else:
                dop.vmult_ptr(pa, pb)

    apply_steps(args.warmup)
    torch.cuda.synchronize()
    # self-check: the overlapped apply (P2P stores or NCCL on the side stream) against the plain sequence
    # cell loop -> pack -> NCCL all_to_all -> ordered accumulate on the same input (atomics: equal to rounding)
    selfcheck = None
    if dop.n_iface_groups:
        y1, y2 = torch.empty_like(ta), torch.empty_like(ta)
        xin = pa  # result of the warm-up applies: replicas of interface DoFs agree across ranks
        dop.vmult_ptr(y1.data_ptr(), xin)
        dop.op.vmult_ptr(y2.data_ptr(), xin)
        dop.exchange.add_interface_contributions(y2.data_ptr())
        torch.cuda.synchronize()
        err = torch.stack([(y1 - y2).abs().max(), y2.abs().max()])
        dist.all_reduce(err, op=dist.ReduceOp.MAX)
        selfcheck = {"overlapped_vs_sequential_max_rel_diff": float(err[0] / err[1])}
        assert selfcheck["overlapped_vs_sequential_max_rel_diff"] < (1e-12 if args.dtype == "f64" else 1e-5), selfcheck
        del y1, y2
    use_graphs = os.environ.get("MFG_NO_GRAPH") is None
    if use_graphs:
        try:
            for d_, s_ in ((pa, pb), (pb, pa)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=main):
                    dop.vmult_ptr(d_, s_)
                graphs[(d_, s_)] = g
        except Exception as e:  # keep going without graphs, say so in the result
            graphs.clear()
            use_graphs = False
            if rank == 0:
                print("CUDA graph capture of the apply failed (%s): eager launches" % e, file=__import__("sys").stderr)
        torch.cuda.synchronize()
        apply_steps(2)
        torch.cuda.synchronize()
    ta.fill_(0.1); tb.fill_(0.1)
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if rank == 0:
        sampler.start()
    # Clock samples (20 ms period) start before the timed region, under the same load.  The number of lead-in applies must be
    # THE SAME ON EVERY RANK (an apply synchronises with its neighbours through the exchange barriers: a rank that runs one
    # more apply than its peers waits for ever), so it is a fixed count, not a time: about 0.3 s at the r = 6 rate.
    n_max = torch.tensor([n], dtype=torch.int64, device="cuda")
    dist.all_reduce(n_max, op=dist.ReduceOp.MAX)     # (the count below is derived from a value every rank agrees on)
    apply_steps(1000 if int(n_max.item()) <= 20000000 else 150)
    torch.cuda.synchronize()
    ta.fill_(0.1); tb.fill_(0.1)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    apply_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms_local = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # cell-kernel time per apply (both launches when the apply is split): CUDA events around every cell-kernel launch,
    # on eager applies right after the timed region (events cannot be read back from inside a captured graph)
    n_k = 20
    dop.op.enable_kernel_timing(True)
    for _ in range(n_k):
        pa, pb = pb, pa
        dop.vmult_ptr(pa, pb)
    torch.cuda.synchronize()
    kernel_ms, kernel_launches = dop.op.kernel_time_ms()
    dop.op.enable_kernel_timing(False)
    t = torch.tensor([ms_local, kernel_ms / n_k], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, k_avg_ms = float(t[0]), float(t[1])

    def make_line(e2e, cg, note):
        ng = dop.n_global
        peak, peak_src = measured_peaks()
        alg_bytes = b_alg(args.degree, args.dim, s) * n
        achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9
        workload = ("bmop: 3D variable-coefficient Laplace apply, FE_Q(%d), " % args.degree
                    + (("the refine_global(%d) cube cut into boxes of %d cells per GPU, " if strong else "one refine_global(%d) cube of %d cells per GPU, ")
                       % (args.refine, dop.mesh.n_cells))
                    + "%s grid of boxes, %d global DoFs (%d per GPU incl. interface replicas), atomic scatter, interface exchange: %s"
                    % ("x".join(map(str, dop.grid)), ng, n,
                       "NVLink P2P stores + device-side barriers" if dop.exchange.symm is not None else "NCCL all_to_all_single"))
        line = {"metric": metric, "value": ng * args.steps / (ms * 1e-3), "unit": "DoFs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload,
                           "l2": "inputs larger than L2"},
                "clocks": clocks,
                "e2e": e2e,
                # per apply and rank: constraint pass + cell kernel(s) + push (or pack) + accumulate; the split apply zeroes
                # dst with cudaMemsetAsync (no zero kernel) and launches the cell kernel twice
                "gpu_launches": args.steps * world * ((dop.op.launches_per_vmult() if not dop.n_iface_groups else dop.op.launches_per_vmult()) + 2),
                "launch_mode": "CUDA graph replay of one apply (cell kernels + pack + all_to_all + accumulate)" if use_graphs else "eager",
                "selfcheck": selfcheck,
                "exchange": ("NVLink P2P stores into the neighbours' symmetric-memory receive buffers + device-side barriers"
                             if dop.exchange.symm is not None else "NCCL all_to_all_single"),
                "overlap": "interface cell groups first (%d of the groups), exchange on a side stream during the interior groups" % dop.n_iface_groups
                           if dop.n_iface_groups else "none",
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             # same kernel and per-GPU workload as the N = 1 line: the committed single-GPU ncu capture applies
                             "traffic": PROFILED_TRAFFIC.get((args.dim, args.degree, args.dtype, args.refine, dop.op.active_variant())),
                             "kernel": "laplace cell kernel (variant %d), per GPU and apply (%d launches), max over ranks"
                                       % (dop.op.active_variant(), kernel_launches // n_k),
                             "kernel_ms": k_avg_ms, "peak_source": peak_src},
                "cpu_baseline": None, "cg_solve": cg}
        if note:
            line["note"] = note
        return line


    # What follows (host-buffer end-to-end steps, the CG solve) has run on 2 GPUs only: a watchdog on every rank makes sure that the
    # apply line measured above is printed even if one of these sections hangs on a larger world (they are then reported as
    # timed out instead of losing the line; all ranks leave).
    section = {"name": "e2e"}
    extra = {"e2e": None, "cg": None, "mg": None}

    base_line = make_line(None, None, None)    # (built now: the watchdog thread only fills in what has finished)

    def on_timeout():
        if rank == 0:
            base_line.update({"e2e": extra["e2e"], "cg_solve": extra["cg"] or {"error": "section '%s' exceeded the watchdog limit" % section["name"]},
                              "mg_solve": extra["mg"],
                              "note": "section '%s' did not finish within %d s: reported without it" % (section["name"], WATCHDOG_S)})
            os.write(json_fd, (json.dumps(base_line) + "\n").encode())
        os._exit(0)

    WATCHDOG_S = int(os.environ.get("MFG_BENCH_WATCHDOG_S", "240"))
    import threading
    watchdog = threading.Timer(WATCHDOG_S, on_timeout)
    watchdog.daemon = True
    watchdog.start()

    # (an exception in one of these sections -- as opposed to a hang, which is the watchdog's -- must not lose the apply line either)
    try:
        e2e_s = e2e_block_s = float("nan")
        n_e2e = 0
        if not getattr(args, "no_e2e", False):
            # end to end: pinned host src -> H2D -> vmult (+exchange) -> D2H, two slots per rank pipelined on three streams (PCIe is
            # full duplex: step k's D2H overlaps step k+1's H2D; the applies stay on the main stream).  Blocking figure next to it.
            hs = [torch.full((n,), 0.1, dtype=tdtype).pin_memory() for _ in range(2)]
            hd = [torch.empty((n,), dtype=tdtype).pin_memory() for _ in range(2)]
            ds = [torch.empty((n,), dtype=tdtype, device="cuda") for _ in range(2)]
            dd = [torch.empty((n,), dtype=tdtype, device="cuda") for _ in range(2)]
            s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
            ev_in = [torch.cuda.Event() for _ in range(2)]; ev_ap = [torch.cuda.Event() for _ in range(2)]; ev_out = [torch.cuda.Event() for _ in range(2)]

            def e2e_blocking():
                ds[0].copy_(hs[0], non_blocking=True)
                dop.vmult_ptr(dd[0].data_ptr(), ds[0].data_ptr())
                hd[0].copy_(dd[0], non_blocking=True)
                torch.cuda.synchronize()

            def e2e_pipelined(steps):
                for k in range(steps):
                    sl = k % 2
                    s_h2d.wait_event(ev_ap[sl])                      # the apply that read this slot's source two steps ago is done
                    with torch.cuda.stream(s_h2d):
                        ds[sl].copy_(hs[sl], non_blocking=True)
                        ev_in[sl].record(s_h2d)
                    main.wait_event(ev_in[sl])
                    main.wait_event(ev_out[sl])                      # this slot's previous result has left the device
                    dop.vmult_graphed(dd[sl].data_ptr(), ds[sl].data_ptr())
                    ev_ap[sl].record(main)
                    s_d2h.wait_event(ev_ap[sl])
                    with torch.cuda.stream(s_d2h):
                        hd[sl].copy_(dd[sl], non_blocking=True)
                        ev_out[sl].record(s_d2h)
                torch.cuda.synchronize()

            e2e_blocking()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                e2e_blocking()
            dist.barrier()
            e2e_block_s = (time.perf_counter() - t0) / 3
            e2e_pipelined(4)   # warm-up: captures the two graphs
            n_e2e = max(args.e2e_steps, 6)
            dist.barrier()
            t0 = time.perf_counter()
            e2e_pipelined(n_e2e)
            dist.barrier()
            e2e_s = (time.perf_counter() - t0) / n_e2e
            te = torch.tensor([e2e_s, e2e_block_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_s, e2e_block_s = float(te[0]), float(te[1])
            del hs, hd, ds, dd
            extra["e2e"] = {"value": dop.n_global / e2e_s, "unit": "DoFs/s", "h2d_bytes_per_step": n * s * world, "d2h_bytes_per_step": n * s * world,
                            "ms_per_step": e2e_s * 1e3, "steps": n_e2e, "pipelined_slots": 2, "blocking_single_call_ms": e2e_block_s * 1e3,
                            "host_buffers": "pinned, rank bound to the GPU's NUMA node (rank 0: node %d)" % numa_node if numa_node is not None
                                            else "pinned, no NUMA binding (topology not visible or a single node)"}


        # CG solve over all GPUs (BASELINE metric "CG time"): b = A u for a vector u whose interface replicas agree
        cg = None
        section["name"] = "cg_solve"
        if not args.no_cg:
            from . import GpuVector as GV
            ta.fill_(1.0)
            dop.vmult_ptr(tb.data_ptr(), ta.data_ptr())      # tb = A 1: bit-identical on the replicas of interface DoFs
            amax = tb.abs().max()
            dist.all_reduce(amax, op=dist.ReduceOp.MAX)
            uvec_t = tb / amax.clamp_min(1e-300) + 1.0       # u = 1 + A1 / max|A1|
            ue = GV.wrap(ctx, uvec_t)
            vb, vx = GV(ctx, n, dtype), GV(ctx, n, dtype)
            dop.vmult(vb, ue)
            vx.fill(0.0)
            bnorm = dop.dot(vb, vb) ** 0.5
            solver_cg_distributed(dop, vx, vb, 0.0, 3)
            vx.fill(0.0)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            its, res = solver_cg_distributed(dop, vx, vb, (1e-12 if args.dtype == "f64" else 1e-5) * bnorm, 20000)
            torch.cuda.synchronize(); dist.barrier()
            cg_s = time.perf_counter() - t0
            vx.add(-1.0, ue)
            err = dop.dot(vx, vx) ** 0.5 / dop.dot(ue, ue) ** 0.5
            cg = {"seconds": cg_s, "iterations": its, "ms_per_iteration": 1e3 * cg_s / max(1, its), "rel_error": err, "n_dofs": dop.n_global,
                  "preconditioner": "jacobi (Chebyshev degree 0)", "tolerance": "1e-12*|b|" if args.dtype == "f64" else "1e-5*|b|",
                  "loop": "distributed.solver_cg_distributed: graph-replayed apply + mfg_cgd_* kernels, scalars all-reduced on the device (NCCL, "
                          "in stream order), one host read of the convergence flag every 8 iterations"}
            extra["cg"] = cg
            # The same system by CG preconditioned with the multigrid V-cycle over the partition (partitioned_mg.py: every level partitioned
            # like the finest, local transfers, Chebyshev(5) smoothers with exchanged diagonals, replicated coarse solve).  First hardware run
            # of this leg is the driver's: any failure is reported in place of the figures, a hang is the watchdog's.
            section["name"] = "mg_solve"
            if not getattr(args, "no_mg", False):
                try:
                    from .partitioned_mg import DistributedLevel, PartitionedMultigrid
                    t0 = time.perf_counter()
                    pm = PartitionedMultigrid(lambda l: DistributedLevel(ctx, rank, world, args.dim, args.degree, l, dtype, strong,
                                                                         dop=dop if l == args.refine else None), 1, args.refine)
                    torch.cuda.synchronize(); dist.barrier()
                    mg_setup_s = time.perf_counter() - t0
                    xf = [GV(ctx, n, dtype)]
                    xf[0].fill(0.0)
                    pm.solve_cg(xf, [vb], 0.0, 1)                  # warm-up: one iteration = one V-cycle
                    xf[0].fill(0.0)
                    pm.coarse_iterations = 0
                    torch.cuda.synchronize(); dist.barrier()
                    t0 = time.perf_counter()
                    its2, res2 = pm.solve_cg(xf, [vb], (1e-10 if args.dtype == "f64" else 1e-5) * bnorm, 100)
                    torch.cuda.synchronize(); dist.barrier()
                    mg_s = time.perf_counter() - t0
                    xf[0].add(-1.0, ue)
                    extra["mg"] = {"seconds": mg_s, "iterations": its2, "rel_error": dop.dot(xf[0], xf[0]) ** 0.5 / dop.dot(ue, ue) ** 0.5,
                                   "n_dofs": dop.n_global, "levels": args.refine, "coarse_cg_iterations": pm.coarse_iterations, "setup_seconds": mg_setup_s,
                                   "tolerance": "1e-10*|b|" if args.dtype == "f64" else "1e-5*|b|",
                                   "preconditioner": "V-cycle over the box partition, levels 1..%d, Chebyshev(5) with the fused update kernel, local transfers, "
                                                     "coarse level solved on one replicated global mesh; host-orchestrated (one host read per dot "
                                                     "product)" % args.refine}
                except Exception as e:
                    extra["mg"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    except Exception as e:
        extra["section_error"] = "section '%s': %s: %s" % (section["name"], type(e).__name__, str(e)[:300])
    watchdog.cancel()
    if rank == 0:
        line = make_line(extra["e2e"], extra["cg"], extra.get("section_error"))
        line["mg_solve"] = extra["mg"]
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    # Teardown: ncclCommDestroy hung on this stack (torch 2.11 / NCCL 2.28) after captured graphs that hold NCCL kernels had
    # run; the graphs are released first and the teardown gets 20 s on a watchdog thread before the ranks leave without it.
    threading.Timer(20.0, lambda: os._exit(0)).start()
    try:
        graphs.clear()
        getattr(dop, "_graphs", {}).clear()
        torch.cuda.synchronize()       # (raises after a sticky device error of a failed section: the line is out, the rank still leaves with 0)
        __import__("sys").stdout.flush()
        __import__("sys").stderr.flush()
        dist.barrier()
        dist.destroy_process_group()
    finally:
        os._exit(0)
