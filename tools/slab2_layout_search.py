"""Search bank-conflict-free shared-memory layouts for the slab2 kernel (kernels_slab2.cuh).

A warp group holds CW = 32 // n cells, cell c = cl + HC*ch with HC = CW/2 (CW even) -- the cells with ch = 0 live in
lanes 0..15, the others in lanes 16..31 ("half-split" lane map), so that a 64-bit gather / scatter instruction touches the
DoFs of HC cells per half warp (tools/gather_line_model.py).  Element (c, i, j, k) of the group sits at
    SL*cl + SH*ch + SI*i + SJ*j + SK*k.
Three thread layouts exist: A (lane <-> (c, i), owns j,k), B (lane <-> (c, k), owns i,j), C (lane <-> (c, j), owns i,k).
A warp instruction in layout X touches, for fixed values of the two owned indices, the addresses of all active lanes
(c, x).  It is conflict free when
  * 8-byte elements: inside each half warp all addresses are distinct modulo 16,
  * 4-byte elements: all addresses are distinct modulo 32.
Lane maps inside a half warp: 'x' = x-major (lane16 = cl + HC*x), 'c' = cell-major (lane16 = n*cl + x); for odd CW
(n = 6) the whole warp is treated as one unit (ch = 0, HC = CW, lanes 0..31).

Each buffer is written in one layout and read in another, so it needs two conflict-free index directions:
  AB (A -> B): i,k     BC (B -> C, also the coefficient image): k,j     CA (C -> A): j,i
The address is a mixed-radix number over the digits in some order with optional padding between digits, which makes it
injective by construction; the smallest footprint wins (the coefficient image is also streamed from HBM, so its padding
costs bandwidth).  Prints constexpr tables for kernels_slab2.cuh."""
import itertools
import sys


def lanes_of(n, lanemap):
    """list of (lane, cl, ch, x)"""
    cw = 32 // n
    split = cw % 2 == 0
    hc = cw // 2 if split else cw
    out = []
    for ch in range(2 if split else 1):
        for cl in range(hc):
            for x in range(n):
                l16 = cl + hc * x if lanemap == 'x' else n * cl + x
                out.append((16 * ch + l16 if split else l16, cl, ch, x))
    return out, hc, split


def conflict_free(lanes, wb, SL, SH, S):
    groups = {}
    for lane, cl, ch, x in lanes:
        key = lane // 16 if wb == 8 else 0
        b = (SL * cl + SH * ch + S * x) % (16 if wb == 8 else 32)
        if b in groups.setdefault(key, set()):
            return False
        groups[key].add(b)
    return True


def search(n, wb, lanemap, dirs, maxpad=8, align=1):
    lanes, hc, split = lanes_of(n, lanemap)
    radix = {'l': hc, 'h': 2 if split else 1, 'i': n, 'j': n, 'k': n}
    digits = 'lhijk' if split else 'lijk'
    best = None
    for order in itertools.permutations(digits):
        for pads in itertools.product(range(maxpad + 1), repeat=len(digits) - 1):
            s = {'h': 0}
            cur = 1
            for d, digit in enumerate(order):
                s[digit] = cur
                cur = cur * radix[digit] + (pads[d] if d < len(pads) else 0)
            foot = sum(s[d] * (radix[d] - 1) for d in digits) + 1
            foot = (foot + align - 1) // align * align
            if best is not None and foot >= best[0]:
                continue
            if all(conflict_free(lanes, wb, s['l'], s['h'], s[d]) for d in dirs):
                best = (foot, s['l'], s['h'], s['i'], s['j'], s['k'])
    return best


if __name__ == "__main__":
    ns = [int(a) for a in sys.argv[1:]] or [2, 3, 4, 5, 6]
    for wb in (8, 4):
        for n in ns:
            dense = (32 // n) * n ** 3
            for lanemap in ("c",):
                row = []
                for name, dirs in (("AB", "ik"), ("BC", "kj"), ("CA", "ji")):
                    row.append((name, search(n, wb, lanemap, dirs, align=16 // wb)))
                print("bytes=%d n=%d lanes=%s dense=%d : " % (wb, n, lanemap, dense) +
                      "  ".join("%s{%d,%d,%d,%d,%d} F=%d" % (nm, r[1], r[2], r[3], r[4], r[5], r[0]) if r else "%s none" % nm for nm, r in row),
                      flush=True)
