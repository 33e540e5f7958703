"""Offline model of the L1 cost of the gather / scatter of the cell kernels: number of distinct 128-byte lines a warp
instruction touches ("L1 tag requests" in ncu) for different lane <-> (cell, element) assignments, computed from the real
DoF map of a uniform mesh.  Used to choose the gather order of kernels_slab2.cuh (validated against the ncu capture of
the first version: 59.2 tag requests per cell for the x-major slab order).
Usage: python tools/gather_line_model.py [p] [r] [bytes]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import OracleMesh  # noqa: E402  (tooling only)


def lines_per_cell(l2g, n, wb, order, interior_only=True):
    """order: function (group cells array [CW, n,n,n] of dof ids) -> list of instructions, each a list of 32 ids (or -1)."""
    cw = 32 // n
    ncell = l2g.shape[0]
    tot_w = tot_h = ninst = 0
    ng = ncell // cw
    for g in range(0, ng, max(1, ng // 400)):  # sample groups
        cells = l2g[g * cw:(g + 1) * cw].reshape(cw, n, n, n)  # [c][k][j][i]
        for inst in order(cells):
            ids = np.asarray(inst)
            ln = np.where(ids >= 0, ids * wb // 128, -1)
            tot_w += len(set(ln[ln >= 0].tolist()))
            if wb == 8:
                for h in (ln[:16], ln[16:]):
                    tot_h += len(set(h[h >= 0].tolist()))
            else:
                tot_h += len(set(ln[ln >= 0].tolist()))
            ninst += 1
    ngs = len(range(0, ng, max(1, ng // 400)))
    return tot_w / (ngs * cw), tot_h / (ngs * cw), ninst / (ngs * cw)


def order_lex(n):
    def f(cells):
        flat = cells.reshape(-1)
        out = []
        for q in range(0, len(flat), 32):
            ch = flat[q:q + 32].tolist()
            out.append(ch + [-1] * (32 - len(ch)))
        return out
    return f


def order_slab(n, lanemap):
    cw = 32 // n

    def f(cells):
        out = []
        for k in range(n):
            for j in range(n):
                inst = [-1] * 32
                for c in range(cw):
                    for i in range(n):
                        if lanemap == 'x':
                            lane = c + cw * i
                        elif lanemap == 'c':
                            lane = n * c + i
                        else:  # half-split: cells [0, cw/2) in lanes 0..15, the rest in 16..31, x-major inside
                            hc = cw // 2
                            lane = 16 * (c // hc) + (c % hc) + hc * i
                        inst[lane] = int(cells[c, k, j, i])
                out.append(inst)
        return out
    return f


if __name__ == "__main__":
    p = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    wb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    o = OracleMesh(3, p, r)
    n = p + 1
    l2g = np.asarray(o.loc2glob, dtype=np.int64)
    print("3D Q%d r=%d: %d cells, %d DoFs, %d-byte values" % (p, r, l2g.shape[0], o.n_dofs, wb))
    for name, f in (("lexicographic chunks of 32", order_lex(n)), ("slab x-major", order_slab(n, 'x')), ("slab cell-major", order_slab(n, 'c')),
                    ("slab half-split", order_slab(n, 'h'))):
        w, h, ni = lines_per_cell(l2g, n, wb, f)
        print("%-28s instr/cell %5.2f  lines/cell (warp) %6.2f  lines/cell (sum over half warps) %6.2f" % (name, ni, w, h))


def order_sorted_cell(n):
    def f(cells):
        out = []
        for c in range(cells.shape[0]):
            flat = np.sort(cells[c].reshape(-1))
            for q in range(0, len(flat), 32):
                ch = flat[q:q + 32].tolist()
                out.append(ch + [-1] * (32 - len(ch)))
        return out
    return f


def order_sorted_group(n, unique):
    def f(cells):
        flat = np.sort(cells.reshape(-1))
        if unique:
            flat = np.unique(flat)
        out = []
        for q in range(0, len(flat), 32):
            ch = flat[q:q + 32].tolist()
            out.append(ch + [-1] * (32 - len(ch)))
        return out
    return f


if __name__ == "__main__":
    for name, f in (("sorted within cell", order_sorted_cell(n)), ("sorted within group", order_sorted_group(n, False)),
                    ("sorted unique within group", order_sorted_group(n, True))):
        w, h, ni = lines_per_cell(l2g, n, wb, f)
        print("%-28s instr/cell %5.2f  lines/cell (warp) %6.2f  lines/cell (sum over half warps) %6.2f" % (name, ni, w, h))
