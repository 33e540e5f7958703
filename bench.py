#!/usr/bin/env python
"""bench.py -- headline benchmark: 3D Q4 FP64 variable-coefficient Laplace apply, DoFs/s.

A "step" is one LaplaceOperatorGpu::vmult (dst = A*src: fused zero/constraint pass + cell kernel)
on the uniform cube mesh, the loop of bmop.cu:135-153 (swap(dst,src); vmult(dst,src)).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--refine R] [--degree P] [--dtype f64|f32]
  python bench.py --impl reference ...   # the reference's CPU path (oracle port, all host threads)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def b_alg(p, dim, s):
    """Algorithmic bytes per DoF (SURVEY.md 8d / BASELINE.md 3)."""
    return 2 * s + (s + 4) * ((p + 1) / p) ** dim + s / p ** dim


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the cell kernel from the committed `ncu --set full`
# captures (profiles/r02_slab3_kernel_q4_f64_r6_ncu.txt, r02_stage_*, r01_column_*), keyed by (dim, degree, dtype, refine, kernel variant); None if not profiled
PROFILED_TRAFFIC = {(3, 4, "f64", 6, 50): 817.0e6, (3, 4, "f64", 6, 40): 635.3e6, (3, 4, "f64", 6, 1): 789.3e6}


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def cpu_reference_run(args, steps, warmup):
    """The reference's CPU path (bmop-cpu.cc:138-155) restated: oracle port, OpenMP over colors, all host threads.
    Bounded sample: the same mesh family at a refinement the host finishes in seconds."""
    from oracle.oracle import OracleMesh, lib as olib
    # all host threads this process may use (torchrun sets OMP_NUM_THREADS=1 for its workers)
    olib().orc_set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    r = min(args.refine, args.cpu_refine)
    m = OracleMesh(3, args.degree, r)
    u = np.full(m.n_dofs, 0.1)
    for _ in range(max(1, warmup)):
        u2 = m.vmult(u, fast=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        u2 = m.vmult(u, fast=True)
        u, u2 = u2, u
    dt = time.perf_counter() - t0
    cores = olib().orc_max_threads()
    return {"value": m.n_dofs * steps / dt, "unit": "DoFs/s", "cores": cores, "kind": "port",
            "sample": "%d applies of 3D Q%d FP64 r=%d (%d DoFs); sum-factorised, 8-cell SIMD batches, OpenMP over 8 colors (restatement of deal.II MatrixFree, not its binary)" % (steps, args.degree, r, m.n_dofs)}, dt / steps


def adaptive_main(args, ctx, metric):
    """BASELINE.json configs[3]: apply and CG solve on the reference's pseudo-adaptive mesh (bmop.cu -DADAPTIVE_GRID,
    bmop_common.h:49-105).  The mesh, its DoFs, the hanging-node masks and the constraint list come from the library's host
    substrate (mfg_amesh_*); cells without constraints run on the fast kernel, cells with a mask on the column kernel with the
    interpolation fused into gather / scatter.  Same JSON layout as the headline line (no e2e / cpu_baseline legs)."""
    import torch
    import dealii_cuda_b200 as mf
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    s = 8 if args.dtype == "f64" else 4
    t0 = time.perf_counter()
    am = mf.AdaptiveMesh(args.dim, args.degree).pseudo_adaptive_refinement(args.refine).distribute_dofs()
    setup_host_s = time.perf_counter() - t0
    masks = am.arrays()["constraint_mask"]
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(am)
    n = am.n_dofs
    ta = torch.full((n,), 0.1, dtype=tdtype, device="cuda")
    tb = torch.zeros((n,), dtype=tdtype, device="cuda")
    pa, pb = ta.data_ptr(), tb.data_ptr()

    def apply_steps(k):
        nonlocal pa, pb
        for _ in range(k):
            pa, pb = pb, pa
            op.vmult_ptr(pa, pb)

    apply_steps(max(args.warmup, 3))
    torch.cuda.synchronize()
    ta.fill_(0.1); tb.fill_(0.1)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    t_lead = time.perf_counter()
    while time.perf_counter() - t_lead < 0.3:
        apply_steps(20)
        torch.cuda.synchronize()
    ta.fill_(0.1); tb.fill_(0.1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    apply_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    cg = None
    if not args.no_cg:
        ue = mf.GpuVector.wrap(ctx, torch.rand((n,), dtype=tdtype, device="cuda", generator=torch.Generator("cuda").manual_seed(1)))
        vb_, vx_ = mf.GpuVector(ctx, n, dtype), mf.GpuVector(ctx, n, dtype)
        op.vmult(vb_, ue)
        op.compute_diagonal()
        mf.solver_cg(op, vx_, vb_, 0.0, 3)
        vx_.fill(0.0)
        ctx.synchronize()
        t0 = time.perf_counter()
        its, res = mf.solver_cg(op, vx_, vb_, (1e-12 if args.dtype == "f64" else 1e-5) * vb_.l2_norm(), 20000)
        ctx.synchronize()
        cg_s = time.perf_counter() - t0
        cg = {"seconds": cg_s, "iterations": its, "ms_per_iteration": 1e3 * cg_s / max(1, its), "last_residual": res, "n_dofs": n,
              "preconditioner": "jacobi (Chebyshev degree 0)", "tolerance": "1e-12*|b|" if args.dtype == "f64" else "1e-5*|b|"}
    # the same system solved by CG preconditioned with the multigrid V-cycle (local smoothing on the adaptive hierarchy,
    # poisson_mg.cu:430-552); needs the vertex-balanced mesh of the reference's MG drivers, so it is built separately
    mg_solve = None
    if not args.no_cg:
        try:
            from dealii_cuda_b200.multigrid import AdaptiveMultigrid
            t0 = time.perf_counter()
            am2 = mf.AdaptiveMesh(args.dim, args.degree, limit_level_difference_at_vertices=True).pseudo_adaptive_refinement(args.refine).distribute_dofs()
            amg = AdaptiveMultigrid(ctx, am2, 0, dtype)
            ctx.synchronize()
            mg_setup_s = time.perf_counter() - t0
            n2 = am2.n_dofs
            con = torch.from_numpy(am2.arrays()["constrained"].astype(np.int64)).cuda()
            tu = torch.rand((n2,), dtype=tdtype, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
            tu[con] = 0                                   # right-hand sides are zero on hanging / boundary rows (ConstraintMatrix::condense)
            ue2 = mf.GpuVector.wrap(ctx, tu)
            b2, x2 = mf.GpuVector(ctx, n2, dtype), mf.GpuVector(ctx, n2, dtype)
            amg.op.vmult(b2, ue2)
            amg.solve_cg(x2, b2, 0.0, 1)                  # warm-up
            x2.fill(0.0)
            ctx.synchronize()
            t0 = time.perf_counter()
            its2, res2 = amg.solve_cg(x2, b2, (1e-10 if args.dtype == "f64" else 1e-5) * b2.l2_norm(), 200)
            ctx.synchronize()
            mg_s = time.perf_counter() - t0
            x2.add(-1.0, ue2)
            mg_solve = {"seconds": mg_s, "iterations": its2, "rel_error": x2.l2_norm() / ue2.l2_norm(), "n_dofs": n2, "n_cells": am2.n_cells,
                        "levels": am2.n_levels, "setup_seconds": mg_setup_s, "tolerance": "1e-10*|b|" if args.dtype == "f64" else "1e-5*|b|",
                        "preconditioner": "V-cycle, local smoothing, Chebyshev(5), edge matrices; mesh with limit_level_difference_at_vertices"}
        except Exception as e:  # the apply / CG figures above stand on their own
            mg_solve = {"error": "%s: %s" % (type(e).__name__, e)}
    peak, peak_src = measured_peaks()
    # algorithmic bytes: two vectors + per cell DoF the index (4 B) and the merged weight (s B); the mask is 4 B per cell
    alg_bytes = 2.0 * s * n + float(am.n_cells) * (am.dofs_per_cell * (4 + s) + 4)
    line = {"metric": metric, "value": n * args.steps / (ms * 1e-3), "unit": "DoFs/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "bmop -DADAPTIVE_GRID: %dD hyper_cube(-1,1), pseudo_adaptive_refinement(%d) (bmop_common.h:49-105), FE_Q(%d): %d active cells on "
                                   "%d levels, %d DoFs, %d cells with hanging-node constraints, atomic scatter"
                                   % (args.dim, args.refine, args.degree, am.n_cells, am.n_levels, n, int((masks != 0).sum())),
                       "l2": "inputs larger than L2" if alg_bytes > 126e6 else "L2-resident"},
            "clocks": clocks, "e2e": None, "gpu_launches": args.steps * op.launches_per_vmult(),
            "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / args.steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "whole vmult (fast kernel on plain cells + "
                         "column kernel with fused hanging-node interpolation)", "peak_source": peak_src},
            "cpu_baseline": None, "cg_solve": cg, "mg_solve": mg_solve, "mesh_setup_host_seconds": setup_host_s}
    print(json.dumps(line))
    return 0


def spmv_main(args, ctx, metric):
    """bmop_spm.cu: the assembled CSR matrix of the operator (CUDAWrappers::SparseMatrix::vmult) against the matrix-free apply on the
    same mesh, bmop loop, DoFs/s.  Algorithmic bytes of a CSR product: (s + 4) per entry + 4 per row + 2 s per row."""
    import torch
    import dealii_cuda_b200 as mf
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    s = 8 if args.dtype == "f64" else 4
    mesh = mf.HyperCubeMesh(ctx, args.dim, args.degree, args.refine)
    t0 = time.perf_counter()
    S = mf.SparseMatrixGpu(ctx, dtype)
    S.reinit(mesh)
    assembly_s = time.perf_counter() - t0
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(mesh)
    n = mesh.n_dofs
    va, vb = mf.GpuVector(ctx, n, dtype), mf.GpuVector(ctx, n, dtype)

    def timed(apply):
        va.fill(0.0); vb.fill(0.1)
        for _ in range(max(args.warmup, 3)):
            apply(va, vb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            apply(va, vb)     # (same input every step: the loop is not renormalised; values do not influence the timing)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    ms_spm = timed(S.vmult)
    ms_mf = timed(op.vmult)
    peak, peak_src = measured_peaks()
    nnz = S.n_nonzero_elements()
    bytes_spm = nnz * (s + 4) + n * (4 + 2 * s)
    line = {"metric": metric.replace("Laplace apply", "Laplace apply, assembled CSR matrix"), "value": n / (ms_spm * 1e-3), "unit": "DoFs/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_spm, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "bmop_spm: %dD FE_Q(%d) refine_global(%d), %d DoFs, %d matrix entries (%.1f per row), CSR kernel with one warp per row"
                                   % (args.dim, args.degree, args.refine, n, nnz, nnz / max(1, n)),
                       "l2": "matrix %.0f MB" % (bytes_spm / 1e6)},
            "roofline": {"bound": "hbm", "achieved": bytes_spm / (ms_spm * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bytes_spm / (ms_spm * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "csr_vmult_warp_per_row", "peak_source": peak_src},
            "matrix_free": {"value": n / (ms_mf * 1e-3), "unit": "DoFs/s", "ms_per_step": ms_mf, "speedup_over_assembled": ms_spm / ms_mf},
            "assembly_host_seconds": assembly_s, "matrix_bytes": S.memory_consumption(), "gpu_launches": args.steps}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--refine", type=int, default=6, help="global refinements of the cube per GPU (r=6: 16,974,593 DoFs)")
    ap.add_argument("--degree", type=int, default=4)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--coloring", action="store_true", help="graph-colored scatter instead of FP64 atomics")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-refine", type=int, default=6, help="CPU baseline sample: global refinements (6 = the GPU workload itself)")
    ap.add_argument("--cpu-steps", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cg", action="store_true", help="skip the CG solve-time measurement (second part of BASELINE's metric)")
    ap.add_argument("--no-e2e", action="store_true", help="N > 1 only: skip the host-buffer end-to-end leg (extra lines at sizes whose pinned buffers would not fit)")
    ap.add_argument("--no-mg", action="store_true", help="N > 1 only: skip the multigrid-preconditioned CG over the partition")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = one refine_global(R) cube per GPU (default), strong = the refine_global(R) cube split over the GPUs")
    ap.add_argument("--adaptive", action="store_true",
                    help="BASELINE configs[3] instead of the headline: the reference's pseudo_adaptive_refinement(R) mesh with hanging nodes "
                         "(bmop_common.h:49-105), apply + CG solve, N = 1")
    ap.add_argument("--spmv", action="store_true",
                    help="the reference's competitor row instead of the headline (bmop_spm.cu): the ASSEMBLED sparse matrix of the same operator "
                         "applied with a CSR kernel, next to the matrix-free apply on the same mesh (use --refine <= 4 in 3D: the matrix of r=4 holds 1e8 entries)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = "%dD Q%d %s Laplace apply throughput" % (args.dim, args.degree, "FP64" if args.dtype == "f64" else "FP32")
    s = 8 if args.dtype == "f64" else 4

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, args.cpu_steps))
        cb, sec = cpu_reference_run(args, steps, min(args.warmup, 2))
        line = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "DoFs/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": "bmop-cpu: 3D unit-cube variable-coefficient Laplace apply, FE_Q(%d), uniform mesh; CPU sample %s"
                                       % (args.degree, cb["sample"])},
                "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "DoFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch
    import dealii_cuda_b200 as mf
    if args.gpus > 1 or world > 1:
        from dealii_cuda_b200 import distributed as mfd
        return mfd.bench_main(args, metric)

    torch.cuda.set_device(local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = mf.Context(local_rank, stream)
    dtype = np.float64 if args.dtype == "f64" else np.float32
    tdtype = torch.float64 if args.dtype == "f64" else torch.float32
    if args.adaptive:
        return adaptive_main(args, ctx, metric)
    if args.spmv:
        return spmv_main(args, ctx, metric)
    mesh = mf.HyperCubeMesh(ctx, args.dim, args.degree, args.refine)
    op = mf.LaplaceOperatorGpu(ctx, dtype, use_coloring=args.coloring)
    op.reinit(mesh)
    if args.variant:
        op.set_variant(args.variant)
    n = mesh.n_dofs
    # device memory through torch (plumbing); vectors wrapped as GpuVectors
    ta = torch.full((n,), 0.1, dtype=tdtype, device="cuda")
    tb = torch.zeros((n,), dtype=tdtype, device="cuda")
    va, vb = mf.GpuVector.wrap(ctx, ta), mf.GpuVector.wrap(ctx, tb)
    pa, pb = ta.data_ptr(), tb.data_ptr()

    def apply_steps(k):
        nonlocal pa, pb
        for _ in range(k):
            pa, pb = pb, pa           # dst.swap(src)
            op.vmult_ptr(pa, pb)      # vmult(dst, src)

    # the raw loop is an un-normalised power iteration; restart from 0.1 every 50 steps so FP32 cannot
    # overflow -- values do not influence timing (no data-dependent paths)
    apply_steps(args.warmup)
    torch.cuda.synchronize()
    ta.fill_(0.1); tb.fill_(0.1)
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # clocks are sampled every 20 ms from 0.3 s before the timed region (the GPU keeps applying the operator, so the samples
    # are taken under the same load) until its end: a short timed region (--steps 20 = 5 ms) still gets its samples
    sampler.start()
    t_lead = time.perf_counter()
    while time.perf_counter() - t_lead < 0.3:
        apply_steps(50)
        torch.cuda.synchronize()
    ta.fill_(0.1); tb.fill_(0.1)
    op.enable_kernel_timing(16)   # CUDA events around every 16th cell-kernel launch of the timed region
    torch.cuda.synchronize()
    e0.record()
    apply_steps(args.steps)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    kernel_ms, kernel_launches = op.kernel_time_ms()
    op.enable_kernel_timing(False)
    if kernel_launches < 24 * op.cell_launches_per_vmult():
        # few steps: the cell-kernel time rests on a post-pass of 24 bracketed applies right behind the timed region
        ta.fill_(0.1); tb.fill_(0.1)
        op.enable_kernel_timing(1)
        apply_steps(24)
        torch.cuda.synchronize()
        kernel_ms, kernel_launches = op.kernel_time_ms()
        op.enable_kernel_timing(False)
    value = n * args.steps / (ms * 1e-3)

    # BASELINE.json configs[0] next to the headline: the reference's own CPU-runnable case (3D Q4 r=5, ~2 M DoFs, L2-resident
    # on B200, 100 applications), same loop
    configs0 = None
    if args.dim == 3 and args.refine > 5 and not args.coloring:
        m5 = mf.HyperCubeMesh(ctx, 3, args.degree, 5)
        op5 = mf.LaplaceOperatorGpu(ctx, dtype)
        op5.reinit(m5)
        a5, b5 = mf.GpuVector(ctx, m5.n_dofs, dtype), mf.GpuVector(ctx, m5.n_dofs, dtype)
        op5.bmop(a5, b5, 20, 0.1)
        ms5 = min(op5.bmop(a5, b5, 100, 0.1) for _ in range(3)) / 100
        configs0 = {"workload": "3D Q%d r=5, %d DoFs, 100 applications (L2-resident)" % (args.degree, m5.n_dofs), "value": m5.n_dofs / (ms5 * 1e-3),
                    "unit": "DoFs/s", "ms_per_step": ms5, "roofline_frac": b_alg(args.degree, 3, s) * m5.n_dofs / (ms5 * 1e-3) / 1e9 / measured_peaks()[0]}
        del op5, a5, b5, m5

    # end-to-end through the public host-buffer API: every step copies its input from pinned host memory (H2D),
    # applies, and copies its result back (D2H).  Two slots are pipelined so that step k's D2H overlaps step k+1's
    # H2D (mfg_laplace_vmult_host_async); the blocking single-call latency is reported next to it.
    # (the pinned buffers are allocated while the thread runs on the GPU's NUMA node -- first touch puts them next to its PCIe root --
    # and the affinity is restored afterwards: the CPU baseline below uses all host threads)
    from dealii_cuda_b200.distributed import bind_to_gpu_numa_node
    saved_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_node = bind_to_gpu_numa_node(local_rank)
    saved_threads = torch.get_num_threads()
    torch.set_num_threads(1)       # (no worker threads are born with the narrowed mask: the CPU baseline below must see every core)
    try:
        hs = [torch.full((n,), 0.1, dtype=tdtype).pin_memory() for _ in range(2)]
        hd = [torch.empty((n,), dtype=tdtype).pin_memory() for _ in range(2)]
    finally:
        torch.set_num_threads(saved_threads)
        if numa_node is not None and saved_affinity:
            os.sched_setaffinity(0, saved_affinity)
    op.vmult_host(hd[0].numpy(), hs[0].numpy())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        op.vmult_host(hd[0].numpy(), hs[0].numpy())
    torch.cuda.synchronize()
    e2e_blocking_s = (time.perf_counter() - t0) / args.e2e_steps
    for k in range(2):
        op.vmult_host_async(hd[k].numpy(), hs[k].numpy(), k)
    op.host_sync()
    n_e2e = max(args.e2e_steps, 4)
    t0 = time.perf_counter()
    for k in range(n_e2e):
        op.vmult_host_async(hd[k % 2].numpy(), hs[k % 2].numpy(), k % 2)
    op.host_sync()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    e2e = {"value": n / e2e_s, "unit": "DoFs/s", "h2d_bytes_per_step": n * s, "d2h_bytes_per_step": n * s,
           "ms_per_step": e2e_s * 1e3, "steps": n_e2e, "pipelined_slots": 2, "blocking_single_call_ms": e2e_blocking_s * 1e3,
           "host_buffers": "pinned, allocated on the GPU's NUMA node (node %d)" % numa_node if numa_node is not None else "pinned, no NUMA placement (topology not visible or a single node)"}

    peak, peak_src = measured_peaks()
    alg_bytes = b_alg(args.degree, args.dim, s) * n
    k_avg_ms = kernel_ms / max(1, kernel_launches) * op.cell_launches_per_vmult()
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": PROFILED_TRAFFIC.get((args.dim, args.degree, args.dtype, args.refine, op.active_variant())),
                "kernel": "laplace cell kernel (variant %d)" % op.active_variant(), "kernel_ms": k_avg_ms, "timed_launches": kernel_launches, "peak_source": peak_src,
                "algorithmic_bytes_per_dof": b_alg(args.degree, args.dim, s),
                "whole_vmult_frac": alg_bytes / (ms / args.steps * 1e-3) / 1e9 / peak}
    cpu_baseline = None
    if not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_run(args, args.cpu_steps, 2)

    def make_line(cg, mg_solve):
        line = {"metric": metric, "value": value, "unit": "DoFs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
                "data": "synthetic",
                "config": {"workload": "bmop: %dD unit-cube [-1,1]^%d variable-coefficient Laplace apply, FE_Q(%d), refine_global(%d): "
                                       "%d cells, %d DoFs, %s scatter; bmop.cu loop" % (args.dim, args.dim, args.degree, args.refine,
                                                                                        mesh.n_cells, n, "colored" if args.coloring else "atomic"),
                           "l2": "inputs larger than L2 (index + coefficient + 2 vectors = %.0f MB)" %
                                 ((mesh.n_cells * mesh.dofs_per_cell * (4 + s) + 2 * n * s) / 1e6)},
                "clocks": clocks, "e2e": e2e, "gpu_launches": args.steps * op.launches_per_vmult(), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "cg_solve": cg, "mg_solve": mg_solve, "configs0_r5": configs0}
        return line

    # The solves come last and under a watchdog: the apply line is printed even if one of them hangs (their loops were refactored
    # after the round's last hardware run); an exception is reported in place of the figures.
    import threading
    done = {"cg": None, "mg": None, "section": "cg_solve", "printed": False}
    print_lock = threading.Lock()

    def emit(note=None):
        with print_lock:
            if done["printed"]:
                return
            done["printed"] = True
            line = make_line(done["cg"], done["mg"])
            if note:
                line["note"] = note
            sys.stdout.write(json.dumps(line) + "\n")
            sys.stdout.flush()

    WATCHDOG_S = float(os.environ.get("MFG_BENCH_WATCHDOG_S", "300"))

    def on_timeout():
        emit("section '%s' did not finish within %g s: reported without it" % (done["section"], WATCHDOG_S))
        os._exit(0)

    watchdog = threading.Timer(WATCHDOG_S, on_timeout)
    watchdog.daemon = True
    watchdog.start()

    # CG solve time on the same operator (BASELINE metric "CG time"; poisson.cu:233-260 control flow, Jacobi
    # preconditioner, |r| <= 1e-12 |b|, right-hand side b = A u for a seeded random u)
    cg = None
    if not args.no_cg:
        try:
            ue = mf.GpuVector.wrap(ctx, torch.rand((n,), dtype=tdtype, device="cuda", generator=torch.Generator("cuda").manual_seed(1)))
            vb_, vx_ = mf.GpuVector(ctx, n, dtype), mf.GpuVector(ctx, n, dtype)
            op.vmult(vb_, ue)
            op.compute_diagonal()
            mf.solver_cg(op, vx_, vb_, 0.0, 3)  # warm-up: loads the solver kernels (lazy module loading), like the warm-up applies
            vx_.fill(0.0)
            ctx.synchronize()
            t0 = time.perf_counter()
            its, res = mf.solver_cg(op, vx_, vb_, (1e-12 if args.dtype == "f64" else 1e-5) * vb_.l2_norm(), 10000)
            ctx.synchronize()
            cg_s = time.perf_counter() - t0
            vx_.add(-1.0, ue)
            cg = {"seconds": cg_s, "iterations": its, "ms_per_iteration": 1e3 * cg_s / max(1, its), "rel_error": vx_.l2_norm() / ue.l2_norm(),
                  "n_dofs": n, "preconditioner": "jacobi (Chebyshev degree 0)", "tolerance": "1e-12*|b|" if args.dtype == "f64" else "1e-5*|b|",
                  "kernels_per_iteration": 3 if op.active_variant() == 50 else 5,
                  "loop": "mfg_solver_cg: cell kernel (h = A d and the partial sums of d.h) + cg_residual + cg_advance (also the operator's zero pass)"}
            del ue, vb_, vx_
            done["cg"] = cg
        except Exception as e:
            done["cg"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    done["section"] = "mg_solve"

    # the same kind of system by CG preconditioned with the geometric multigrid V-cycle (poisson_mg.cu:430-552: Chebyshev(5) smoothers
    # with deal.II's eigenvalue estimate, levels 1..r, coarse CG), the library's loop mfg_mg_solve_cg; tools/solve_large.py is the same run
    mg_solve = None
    if not args.no_cg:
        try:
            from dealii_cuda_b200.multigrid import GeometricMultigrid
            t0 = time.perf_counter()
            gm = GeometricMultigrid(ctx, args.dim, args.degree, 1, args.refine, dtype)
            ctx.synchronize()
            mg_setup_s = time.perf_counter() - t0
            gop = gm.ops[args.refine]
            ue = mf.GpuVector.wrap(ctx, torch.rand((n,), dtype=tdtype, device="cuda", generator=torch.Generator("cuda").manual_seed(1)))
            vb_, vx_ = mf.GpuVector(ctx, n, dtype), mf.GpuVector(ctx, n, dtype)
            gop.vmult(vb_, ue)
            gm.solve_cg(vx_, vb_, 0.0, 1)    # warm-up: one V-cycle
            vx_.fill(0.0)
            ctx.synchronize()
            t0 = time.perf_counter()
            its2, res2 = gm.solve_cg(vx_, vb_, (1e-10 if args.dtype == "f64" else 1e-5) * vb_.l2_norm(), 100)
            ctx.synchronize()
            mg_s = time.perf_counter() - t0
            vx_.add(-1.0, ue)
            mg_solve = {"seconds": mg_s, "iterations": its2, "rel_error": vx_.l2_norm() / ue.l2_norm(), "n_dofs": n, "levels": args.refine,
                        "setup_seconds": mg_setup_s, "tolerance": "1e-10*|b|" if args.dtype == "f64" else "1e-5*|b|",
                        "preconditioner": "geometric multigrid V-cycle (mfg_mg_*): Chebyshev(5) smoothers, levels 1..%d, coarse CG" % args.refine}
            del gm, gop, ue, vb_, vx_
        except Exception as e:  # the apply / CG figures stand on their own
            mg_solve = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    done["mg"] = mg_solve
    watchdog.cancel()
    emit()
    return 0


if __name__ == "__main__":
    sys.exit(main())
