"""Plan of the staged cell kernel (dealii_cuda_b200/csrc/stage_plan.cu) checked on the CPU.

The staged kernel replaces the reference's per-entry gather (fee_gpu.cuh:323-338) and atomic scatter (fee_gpu.cuh:346-365)
by staged copies driven by tables that stage_plan.cu derives from the index array.  The builder is plain host code behind
the C ABI (mfg_stage_plan_*), so the tables can be checked here without a GPU: this file replays the kernel's data movement
(copy lists -> staging buffer -> slab reads; face merges -> staged results -> plain stores / red.add) with numpy and
compares it with the plain gather `src[loc2glob]` and scatter-add `dst[loc2glob] += r` on the oracle's meshes.
"""
import ctypes as C

import numpy as np
import pytest

from dealii_cuda_b200 import _capi
from oracle.oracle import OracleMesh  # checker: mesh + DoF numbering

CBIT = np.uint32(0x80000000)
DEAD = 0x8000
H_OWN, H_NHALO, PH = 0, 1, 16


def lane_map(n, cw, hc):
    split = cw % 2 == 0
    out = []
    for l in range(32):
        ch, l16 = (l // 16, l % 16) if split else (0, l)
        if l16 >= (hc if split else cw) * n:
            out.append((-1, 0))
        else:
            out.append(((hc * ch if split else 0) + l16 // n, l16 % n))
    return out


class Plan:
    def __init__(self, degree, dtype, idx, n_plain, n_dofs, merge_dirs=7):
        lib = _capi.lib
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        h = C.c_void_p()
        rc = lib.mfg_stage_plan_build(degree, _capi.F64 if dtype == np.float64 else _capi.F32, n_plain, idx.shape[0], n_dofs,
                                      idx.ctypes.data_as(C.POINTER(C.c_uint32)), merge_dirs, C.byref(h))
        assert rc == 0, lib.mfg_last_error()
        self.wb = 8 if dtype == np.float64 else 4
        info = (C.c_uint32 * 16)()
        assert lib.mfg_stage_plan_info(h, info) == 0
        (self.n_groups, self.n_patterns, self.pstride, n_halo, n_fb, self.cw, self.hc, self.xcap, self.n, self.n_staged, self.lcap, self.ocap,
         self.rd, self.wr, self.cp) = list(info)[:15]
        self.gdesc = np.zeros((self.n_groups, 4), np.uint32)
        self.halo = np.zeros(n_halo, np.uint32)
        self.ptab = np.zeros((self.n_patterns, self.pstride), np.uint16)
        self.fallback = np.zeros(n_fb, np.uint32)
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        assert lib.mfg_stage_plan_get(h, p(self.gdesc, C.c_uint32), p(self.halo, C.c_uint32), p(self.ptab, C.c_uint16), p(self.fallback, C.c_uint32)) == 0
        lib.mfg_stage_plan_destroy(h)


def replay(plan, idx, n_plain, n_dofs, rng):
    """gather and scatter of every group through the plan; returns (gathered [cells][npc], dst, counters)"""
    n, cw, xcap = plan.n, plan.cw, plan.xcap
    ns, npc = n * n, n ** 3
    lanes = lane_map(n, cw, plan.hc)
    src = rng.random(n_dofs) + 1.0
    res = rng.random((idx.shape[0], npc)) + 1.0  # "results" of the cell operator
    gathered = np.full((n_plain, npc), np.nan)
    dst = np.zeros(n_dofs)
    n_plain_st, n_red = np.zeros(n_dofs, int), np.zeros(n_dofs, int)
    fb = set(int(g) for g in plan.fallback)
    for g in range(plan.n_groups):
        cells = [c for c in range(g * cw, g * cw + cw) if c < n_plain]
        own_base, halo_off, z, mm = (int(v) for v in plan.gdesc[g])
        pat = z >> 16
        if pat == 0xffff:
            assert g in fb
            for c in cells:  # slab2 kernel: plain gather, red.add of every unconstrained entry
                e = idx[c]
                u = e & ~CBIT
                ok = (e & CBIT) == 0
                gathered[c] = np.where(ok, src[u], 0.0)
                np.add.at(dst, u[ok], res[c][ok])
                np.add.at(n_red, u[ok], 1)
            continue
        assert g not in fb
        n_halo = z & 0xffff
        t = plan.ptab[pat]
        ns2 = (ns + 1) // 2
        hdr = t[:PH].astype(int)
        pos32 = t[PH:PH + 2 * ns2 * 32].view(np.uint32).reshape(ns2, 32)
        pos = np.empty((2 * ns2, 32), int)
        pos[0::2], pos[1::2] = pos32 & 0xffff, pos32 >> 16
        pos = (pos & 0x7fff) // plan.wb | (pos & DEAD)  # (the tables hold byte offsets)
        own32 = t[PH + 2 * ns2 * 32:PH + 2 * ns2 * 32 + 2 * plan.ocap].view(np.uint32)
        hs = t[PH + 2 * ns2 * 32 + 2 * plan.ocap:PH + 2 * ns2 * 32 + 2 * plan.ocap + (plan.lcap - plan.ocap)].astype(int)
        st = np.concatenate([(own32 & 0xffff).astype(int)[:hdr[H_OWN]], hs]) // plan.wb
        oflag = (own32 >> 16).astype(np.uint8)
        own_total = hdr[H_OWN]
        assert hdr[H_NHALO] == n_halo and own_total + n_halo <= xcap - 1
        used = st[:own_total + n_halo]
        assert len(set(used.tolist())) == len(used) and used.max(initial=0) < xcap - 1, "load-list entries must have distinct slots"
        hl = plan.halo[halo_off:halo_off + n_halo].astype(int)
        # ---- copies into the staging buffer ----
        X = np.full(xcap, np.nan)
        X[xcap - 1] = 0.0
        X[st[:own_total]] = src[own_base:own_base + own_total]
        X[st[own_total:own_total + n_halo]] = src[hl]
        # ---- slab reads ----
        vals = np.zeros((cw, npc))
        for l, (c, i) in enumerate(lanes):
            for s in range(ns):
                v = X[pos[s, l] & 0x7fff]
                if c >= 0 and g * cw + c < n_plain:
                    gathered[g * cw + c, i + n * s] = v
                else:
                    assert pos[s, l] == ((xcap - 1) | DEAD)
        # ---- merges on the results (x, z, y; what was handed over is zeroed) ----
        for c in cells:
            vals[c - g * cw] = res[c]
        v4 = vals.reshape(cw, n, n, n)  # [c][k][j][i]
        for d, step in ((0, 1), (2, 4), (1, 2)):
            snap = v4.copy()
            for c in range(cw):
                if c < 10 and (mm >> (10 * d + c)) & 1:
                    assert c + step < cw
                    if d == 0:
                        v4[c + step, :, :, 0] += snap[c, :, :, n - 1]; v4[c, :, :, n - 1] = 0
                    elif d == 1:
                        v4[c + step, :, 0, :] += snap[c, :, n - 1, :]; v4[c, :, n - 1, :] = 0
                    else:
                        v4[c + step, 0, :, :] += snap[c, n - 1, :, :]; v4[c, n - 1, :, :] = 0
        # ---- staged results ----
        P = np.full(xcap, np.nan)
        written = np.zeros(xcap, int)
        for l, (c, i) in enumerate(lanes):
            for s in range(ns):
                pz = pos[s, l]
                if pz & DEAD:
                    continue
                assert c >= 0
                P[pz] = vals[c, i + n * s]
                written[pz] += 1
        assert written.max(initial=0) <= 1, "two lanes write the same staging slot"
        # ---- write-out ----
        for e in range(own_total):
            f, dd = oflag[e], own_base + e
            if f == 0:
                continue
            v = P[st[e]]
            assert not np.isnan(v), "own DoF without a holder"
            if f == 1:
                dst[dd] = v
                n_plain_st[dd] += 1
            else:
                dst[dd] += v
                n_red[dd] += 1
        for k in range(n_halo):
            v = P[st[own_total + k]]
            assert not np.isnan(v), "halo entry without a holder"
            dst[hl[k]] += v
            n_red[hl[k]] += 1
    return src, res, gathered, dst, n_plain_st, n_red


def make_idx(o):
    l2g = o.loc2glob.astype(np.uint32).copy()
    flag = np.zeros(o.n_dofs, bool)
    flag[o.constrained] = True
    l2g[flag[l2g]] |= CBIT
    return l2g


def check(plan, idx, n_plain, n_dofs, seed=0):
    rng = np.random.default_rng(seed)
    src, res, gathered, dst, n_plain_st, n_red = replay(plan, idx, n_plain, n_dofs, rng)
    e = idx[:n_plain]
    u, ok = e & ~CBIT, (e & CBIT) == 0
    want = np.where(ok, src[u], 0.0)
    assert np.array_equal(gathered, want), "staged gather differs from src[loc2glob]"
    ref = np.zeros(n_dofs)
    np.add.at(ref, u[ok], res[:n_plain][ok])
    assert np.allclose(dst, ref, rtol=1e-13, atol=0), "staged scatter differs from the scatter-add"
    # a DoF written with a plain store is written exactly once and by nobody else (cells behind n_plain included)
    assert n_plain_st.max(initial=0) <= 1
    assert not np.any((n_plain_st == 1) & (n_red > 0))
    all_u, all_ok = idx & ~CBIT, (idx & CBIT) == 0
    touched_late = np.zeros(n_dofs, bool)
    touched_late[all_u[n_plain:][all_ok[n_plain:]]] = True
    assert not np.any((n_plain_st == 1) & touched_late)
    return n_plain_st, n_red


@pytest.mark.parametrize("p,r", [(2, 2), (3, 2), (4, 1), (4, 2), (4, 3), (5, 2)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_plan_reproduces_gather_and_scatter(p, r, dtype):
    o = OracleMesh(3, p, r)
    idx = make_idx(o)
    plan = Plan(p, dtype, idx, idx.shape[0], o.n_dofs)
    assert plan.n_staged + len(plan.fallback) == plan.n_groups
    n_plain_st, n_red = check(plan, idx, idx.shape[0], o.n_dofs)
    if r >= 2:
        assert plan.n_staged > 0.8 * plan.n_groups, "most groups of a uniform mesh must take the staged path"
        assert n_plain_st.sum() > 0.15 * o.n_dofs, "group-interior DoFs must get plain stores"


def test_plan_non_cubic_box_and_partial_last_group():
    o = OracleMesh(3, 4, box=dict(log2_cells=(1, 2, 1), origin=(-1.0, -1.0, -1.0), h=0.5))  # 2 x 4 x 2 cells = 16: groups of 6 -> the last group holds 4 cells
    idx = make_idx(o)
    plan = Plan(4, np.float64, idx, idx.shape[0], o.n_dofs)
    check(plan, idx, idx.shape[0], o.n_dofs)


def test_plan_without_merges_and_with_cells_behind_n_plain():
    """merge_dirs = 0: every shared DoF has several holders (extra halo copies, never a plain store on a shared DoF);
    cells behind n_plain (the hanging-node cells of the operator) only count for the multiplicities"""
    o = OracleMesh(3, 3, 2)
    idx = make_idx(o)
    plan = Plan(3, np.float64, idx, idx.shape[0], o.n_dofs, merge_dirs=0)
    check(plan, idx, idx.shape[0], o.n_dofs)
    n_plain = idx.shape[0] - 13
    plan = Plan(3, np.float64, idx, n_plain, o.n_dofs)
    check(plan, idx, n_plain, o.n_dofs)


def test_plan_shuffled_cells_and_renumbered_dofs():
    """nothing in the plan relies on Morton order or on deal.II's numbering: random cell order and a random DoF
    permutation (no contiguous own ranges left) still reproduce gather and scatter"""
    o = OracleMesh(3, 4, 2)
    idx = make_idx(o)
    rng = np.random.default_rng(5)
    cells = rng.permutation(idx.shape[0])
    check(Plan(4, np.float64, idx[cells], idx.shape[0], o.n_dofs), idx[cells], idx.shape[0], o.n_dofs)
    perm = rng.permutation(o.n_dofs).astype(np.uint32)
    idx2 = (perm[idx & ~CBIT] | (idx & CBIT)).astype(np.uint32)
    check(Plan(4, np.float64, idx2, idx.shape[0], o.n_dofs), idx2, idx.shape[0], o.n_dofs)
