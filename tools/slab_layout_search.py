"""Search conflict-free shared-memory strides for the slab kernel's transposes.
Lane L = n*c + a (c = cell in warp group, a = thread in cell).  Address of element (c,i,j,k) = SC*c + i + RJ*j + RK*k.
Layouts: L = lexicographic staging (entry e = 32q + lane), A = S_xy (a=k), B = S_yz (a=i), C = S_xz (a=j).
Prints constexpr tables for kernels_slab.cuh."""
import itertools


def wf(addrs, wordbytes):
    if wordbytes == 8:
        tot = 0
        for half in (addrs[:16], addrs[16:]):
            banks = {}
            for a in half:
                if a is not None:
                    banks.setdefault(a % 16, set()).add(a)
            tot += max([len(v) for v in banks.values()], default=0)
        return tot
    banks = {}
    for a in addrs:
        if a is not None:
            banks.setdefault(a % 32, set()).add(a)
    return max([len(v) for v in banks.values()], default=0)


def pat(kind, n, RJ, RK, SC, wb):
    cw = 32 // n
    lanes = [(L // n, L % n) if L < cw * n else None for L in range(32)]
    tot = 0
    rng = list(itertools.product(range(n), range(n)))
    if kind == 'A':
        for (i, j) in rng:
            tot += wf([None if ca is None else SC * ca[0] + i + RJ * j + RK * ca[1] for ca in lanes], wb)
    elif kind == 'B':
        for (j, k) in rng:
            tot += wf([None if ca is None else SC * ca[0] + ca[1] + RJ * j + RK * k for ca in lanes], wb)
    elif kind == 'C':
        for (i, k) in rng:
            tot += wf([None if ca is None else SC * ca[0] + i + RJ * ca[1] + RK * k for ca in lanes], wb)
    else:
        npc = n ** 3
        tot_e = cw * npc
        for q in range((tot_e + 31) // 32):
            ad = []
            for L in range(32):
                e = 32 * q + L
                if e >= tot_e:
                    ad.append(None)
                    continue
                c, r = divmod(e, npc)
                ad.append(SC * c + (r % n) + RJ * ((r // n) % n) + RK * (r // (n * n)))
            tot += wf(ad, wb)
    return tot


if __name__ == "__main__":
    for wb in (8, 4):
        for n in range(2, 8):
            for pair in (('L', 'B'), ('A', 'B'), ('A', 'C')):
                best = None
                for RJ in range(n, n + 10):
                    for RK in range(n * RJ, n * RJ + 18):
                        for SC in range(n * RK, n * RK + 34):
                            a, b = pat(pair[0], n, RJ, RK, SC, wb), pat(pair[1], n, RJ, RK, SC, wb)
                            key = (a + b, SC)
                            if best is None or key < best[0]:
                                best = (key, RJ, RK, SC, a, b)
                dense = (pat(pair[0], n, n, n * n, n ** 3, wb), pat(pair[1], n, n, n * n, n ** 3, wb))
                print("bytes=%d n=%d pair=%s%s: RJ=%d RK=%d SC=%d wavefronts %d+%d (dense %d+%d)" %
                      (wb, n, pair[0], pair[1], best[1], best[2], best[3], best[4], best[5], dense[0], dense[1]), flush=True)
