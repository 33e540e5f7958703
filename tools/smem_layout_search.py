"""Brute-force search of shared-memory strides (row RJ, plane RK, cell SC) that make the three
register-axis layouts of the column kernel bank-conflict free.  Element (i,j,k) of local cell lc lives at
SC*lc + i + RJ*j + RK*k (in units of Number).  Prints a C++ table for kernels_v0.cuh."""
import itertools
import sys


def wavefronts(addrs, wordbytes):
    """addrs: list of element offsets (one per active lane, None = inactive), 32 lanes."""
    if wordbytes == 8:
        tot = 0
        for half in (addrs[:16], addrs[16:]):
            banks = {}
            for a in half:
                if a is None:
                    continue
                banks.setdefault(a % 16, set()).add(a)
            tot += max([len(v) for v in banks.values()], default=0)
        return tot
    banks = {}
    for a in addrs:
        if a is None:
            continue
        banks.setdefault(a % 32, set()).add(a)
    return max([len(v) for v in banks.values()], default=0)


def cost(dim, n, cpb, nthreads, RJ, RK, SC, wordbytes):
    tpc = n * n if dim == 3 else n
    total = 0
    layouts = (0, 1, 2) if dim == 3 else (0, 1)
    for r in layouts:
        for e in range(n):
            for w in range(nthreads // 32):
                addrs = []
                for lane in range(32):
                    tid = 32 * w + lane
                    lc, t = tid // tpc, tid % tpc
                    if lc >= cpb:
                        addrs.append(None)
                        continue
                    if dim == 3:
                        if r == 2:
                            i, j, k = t % n, t // n, e
                        elif r == 1:
                            i, j, k = t % n, e, t // n
                        else:
                            i, j, k = e, t % n, t // n
                    else:
                        k = 0
                        if r == 1:
                            i, j = t, e
                        else:
                            i, j = e, t
                    addrs.append(SC * lc + i + RJ * j + RK * k)
                total += wavefronts(addrs, wordbytes)
    return total


def ideal(dim, n, cpb, nthreads, wordbytes):
    tpc = n * n if dim == 3 else n
    per = 0
    for w in range(nthreads // 32):
        act = sum(1 for lane in range(32) if (32 * w + lane) // tpc < cpb)
        per += -(-act * wordbytes // 128)
    return per * n * (3 if dim == 3 else 2)


CPB3 = {2: 32, 3: 14, 4: 8, 5: 5, 6: 7, 7: 5, 8: 2, 9: 3}
CPB2 = {2: 64, 3: 42, 4: 32, 5: 25, 6: 21, 7: 18, 8: 16, 9: 14}

if __name__ == "__main__":
    for dim in (3, 2):
        for wordbytes in (8, 4):
            for n in range(2, 10):
                cpb = (CPB3 if dim == 3 else CPB2)[n]
                tpc = n * n if dim == 3 else n
                nthreads = ((cpb * tpc + 31) // 32) * 32
                base = cost(dim, n, cpb, nthreads, n, n * n, n ** dim, wordbytes)
                best = None
                rj_range = range(n, n + 9) if dim == 3 else [n]
                for RJ in rj_range:
                    rk0 = n * RJ if dim == 3 else 0
                    for RK in (range(rk0, rk0 + 17) if dim == 3 else range(n, n + 9)):
                        # dim 2: RK plays the role of the row stride RJ
                        if dim == 2:
                            rj, rk = RK, 0
                            sc0 = n * rj
                        else:
                            rj, rk = RJ, RK
                            sc0 = n * rk
                        for SC in range(sc0, sc0 + 33):
                            c = cost(dim, n, cpb, nthreads, rj, rk, SC, wordbytes)
                            key = (c, SC)
                            if best is None or key < best[0]:
                                best = (key, rj, rk, SC)
                print("dim=%d bytes=%d n=%d: unpadded %d -> best %d (ideal %d) RJ=%d RK=%d SC=%d (dense %d)" %
                      (dim, wordbytes, n, base, best[0][0], ideal(dim, n, cpb, nthreads, wordbytes), best[1], best[2], best[3], n ** dim),
                      flush=True)
