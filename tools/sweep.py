"""Degree / dimension / dtype / scatter sweep of the Laplace apply (BASELINE.json configs[1]).
Prints one line per configuration: DoFs/s and fraction of the HBM roofline (algorithmic bytes)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dealii_cuda_b200 as mf  # noqa: E402
from bench import b_alg, measured_peaks  # noqa: E402


def run(ctx, dim, p, r, dtype, coloring, steps, variant=0):
    mesh = mf.HyperCubeMesh(ctx, dim, p, r)
    op = mf.LaplaceOperatorGpu(ctx, dtype, use_coloring=coloring)
    op.reinit(mesh)
    if variant:
        op.set_variant(variant)
    n = mesh.n_dofs
    a, b = mf.GpuVector(ctx, n, dtype), mf.GpuVector(ctx, n, dtype)
    op.bmop(a, b, 5, 0.1)
    best = 1e30
    for _ in range(3):
        best = min(best, op.bmop(a, b, steps, 0.1) / steps)
    s = np.dtype(dtype).itemsize
    peak, _ = measured_peaks()
    gdofs = n / (best * 1e-3) / 1e9
    frac = gdofs * b_alg(p, dim, s) / peak
    return dict(dim=dim, p=p, r=r, dtype=np.dtype(dtype).name, coloring=coloring, variant=op.active_variant(), requested=variant, n_dofs=n, ms=best,
                gdofs=gdofs, roofline_frac=frac)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--custom", default="", help="semicolon-separated dim,p,r,f64|f32,variant cases (A/B runs)")
    args = ap.parse_args()
    ctx = mf.Context(0, torch.cuda.current_stream().cuda_stream)
    cases = []
    if args.custom:
        for c in args.custom.split(";"):
            d, p, r, dt, v = c.split(",")
            cases.append((int(d), int(p), int(r), np.float64 if dt == "f64" else np.float32, False, int(v)))
    elif args.quick:
        cases = [(3, 4, 5, np.float64, False, 0), (3, 4, 6, np.float64, False, 0), (3, 4, 6, np.float64, False, 1), (3, 4, 6, np.float64, False, 40),
                 (3, 4, 6, np.float64, False, 51), (3, 4, 6, np.float64, False, 52), (3, 4, 6, np.float64, False, 53), (3, 4, 6, np.float64, True, 1),
                 (3, 4, 6, np.float32, False, 0), (3, 4, 6, np.float32, False, 51), (3, 4, 6, np.float32, False, 40)]
    else:
        # ~16M+ DoFs per case where memory allows
        r3 = {1: 8, 2: 7, 3: 6, 4: 6, 5: 6, 6: 5, 7: 5, 8: 5}
        r2 = {1: 10, 2: 10, 3: 10, 4: 10, 5: 9, 6: 9, 7: 9, 8: 9}
        for dt in (np.float64, np.float32):
            for p in range(1, 9):
                cases.append((3, p, r3[p], dt, False))
            for p in range(1, 9):
                cases.append((2, p, r2[p], dt, False))
        cases += [(3, 4, 5, np.float64, False), (3, 4, 5, np.float64, True), (3, 4, 6, np.float64, True), (3, 4, 7, np.float64, False)]
    for c in cases:
        try:
            print(json.dumps(run(ctx, *c[:5], steps=args.steps, variant=(c[5] if len(c) > 5 else 0))), flush=True)
        except Exception as e:  # keep sweeping
            print(json.dumps(dict(case=str(c), error=str(e))), flush=True)
