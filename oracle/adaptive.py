"""CPU oracle for adaptively refined meshes with hanging nodes (TEST INFRASTRUCTURE, numpy / pure Python,
small meshes only).  Stands in for what the reference obtains from deal.II on adaptive meshes and restates
the reference's hanging-node scheme:

  * 2:1-balanced quad/octree on hyper_cube(left,right), FE_Q(p) DoFs attached to mesh entities
    (vertices, lines, quads, hexes) and numbered first-touch over the active cells (level-major, Morton order
    inside a level), hierarchic order inside a cell                        [deal.II DoFHandler::distribute_dofs]
  * per active cell the 9-bit constraint mask and the loc2glob rewrite of HangingNodes::setup_constraints
    (matrix_free_gpu/hanging_nodes.cuh:209-454; mask bit layout :38-50)
  * resolve_hanging_nodes (hanging_nodes.cuh:617-778) in numpy, the 1-D weights of setup_constraint_weights
    (:580-598): W[k][i] = phi_i(xi_k / 2)
  * the cell operator (same bilinear form as oracle/mf_oracle.c) and vmult with identity on constrained rows
  * an INDEPENDENT check that does not use masks at all: geometric constraint matrix C (every hanging DoF =
    coarse face/edge polynomial evaluated at its support point) and y = C^T A C u.

PARITY STATUS: unpinned at the deal.II boundary (no deal.II here); the interpolation itself is pinned by the
exact-zero known-answer test of test_hanging_node_interpolation.cu:311-350 (linear polynomial reproduced
exactly for every mask), see tests/test_hanging_nodes.py.
"""
import itertools

import numpy as np

from .oracle import hier_to_lex, shape_1d

CONSTR_TYPE = (1 << 0, 1 << 1, 1 << 2)      # hanging_nodes.cuh:38-40
CONSTR_FACE = (1 << 3, 1 << 4, 1 << 5)      # :43-45
CONSTR_EDGE_XY, CONSTR_EDGE_YZ, CONSTR_EDGE_ZX = 1 << 6, 1 << 7, 1 << 8   # :48-50
# edge bit by the direction the edge runs along (the bit names the two normal directions)
EDGE_BIT_ALONG = (CONSTR_EDGE_YZ, CONSTR_EDGE_ZX, CONSTR_EDGE_XY)


def constraint_weights(p):
    """W[k][i] = phi_i(xi_k / 2): coarse 1-D basis at the support points of the first child (lexicographic)."""
    _, _, xn, _, _ = shape_1d(p)
    n = p + 1
    W = np.zeros((n, n))
    for k in range(n):
        x = xn[k] / 2
        for i in range(n):
            W[k, i] = np.prod([(x - xn[m]) / (xn[i] - xn[m]) for m in range(n) if m != i])
    return W


def resolve_hanging_nodes(values, mask, p, dim, transpose=False):
    """resolve_hanging_nodes_shmem (hanging_nodes.cuh:760-778) on one cell tensor, numpy.
    values: array of shape (n,)*dim indexed [z][y][x] (lexicographic, x fastest)."""
    if mask == 0:
        return values
    n = p + 1
    W = constraint_weights(p)
    v = values.copy()
    for d in range(dim):                                   # sweep along x, then y, (then z)
        others = [a for a in range(dim) if a != d]
        this_type = bool(mask & CONSTR_TYPE[d])
        new = v.copy()
        for idx in itertools.product(range(n), repeat=dim):   # idx = (x, y[, z])
            if dim == 2:
                a = others[0]
                on = (idx[a] == 0) if (mask & CONSTR_TYPE[a]) else (idx[a] == p)
                flag = bool(mask & CONSTR_FACE[a]) and on
            else:
                f1, f2 = (d + 1) % 3, (d + 2) % 3              # hanging_nodes.cuh:633-645
                on1 = (idx[f1] == 0) if (mask & CONSTR_TYPE[f1]) else (idx[f1] == p)
                on2 = (idx[f2] == 0) if (mask & CONSTR_TYPE[f2]) else (idx[f2] == p)
                flag = (bool(mask & CONSTR_FACE[f1]) and on1) or (bool(mask & CONSTR_FACE[f2]) and on2) or \
                       (bool(mask & EDGE_BIT_ALONG[d]) and on1 and on2)
            if not flag:
                continue
            k = idx[d]
            t = 0.0
            for i in range(n):
                src = list(idx); src[d] = i
                if this_type:
                    w = W[i, k] if transpose else W[k, i]
                else:
                    w = W[p - i, p - k] if transpose else W[p - k, p - i]
                t += w * v[tuple(reversed(src))]
            new[tuple(reversed(idx))] = t
        v = new
    return v


class AdaptiveMesh:
    """2:1-balanced tree mesh with FE_Q(p) DoFs and the reference's hanging-node data."""

    def __init__(self, dim, p, base_refine, refine_steps, left=-1.0, right=1.0, balance="vertex", cells=None):
        """refine_steps: list of callables f(center ndarray, half-size) -> bool (refine this active cell).
        balance: "vertex" keeps neighbours across faces, edges AND vertices within one level; "dealii" is deal.II's rule
        (faces, in 3D also edges).  cells: an explicit list of active cells (level, x, y[, z]) IN THE ORDER TO USE (e.g. the
        creation order of a deal.II-like triangulation); base_refine / refine_steps are then ignored."""
        self.dim, self.p, self.n, self.left, self.right = dim, p, p + 1, left, right
        self.max_nonzero = dim if balance == "vertex" else (2 if dim == 3 else 1)
        if cells is not None:
            self.cells = [tuple(int(v) for v in c[:1 + dim]) for c in cells]
            assert len(set(self.cells)) == len(self.cells)
        else:
            cells = {(0,) + (0,) * dim}
            for _ in range(base_refine):
                cells = set(ch for c in cells for ch in self._children(c))
            for crit in refine_steps:
                flagged = [c for c in cells if crit(*self._center(c))]
                for c in flagged:
                    cells = self._refine(cells, c)
            self.cells = sorted(cells, key=lambda c: (c[0], self._morton(c)))   # active cells, level-major
        self.n_cells = len(self.cells)
        self.lmax = max(c[0] for c in self.cells)
        self._build_lookup()
        self._number_dofs()
        self._hanging_setup()
        self._geometry()

    # ---- tree -----------------------------------------------------------------------------------------
    def _children(self, c):
        l, xs = c[0], c[1:]
        return [(l + 1,) + tuple(2 * x + b for x, b in zip(xs, bits)) for bits in itertools.product((0, 1), repeat=self.dim)]

    def _morton(self, c):
        key = 0
        for b in range(c[0]):
            for d in range(self.dim):
                key |= ((c[1 + d] >> b) & 1) << (self.dim * b + d)
        return key

    def _center(self, c):
        h = (self.right - self.left) / (1 << c[0])
        return np.array([self.left + h * (x + 0.5) for x in c[1:]]), h / 2

    def _find_active(self, cells, level, xs):
        """active cell containing the level-`level` cell position xs (or None if outside / finer cells there)"""
        if any(x < 0 or x >= (1 << level) for x in xs):
            return None
        for l in range(level, -1, -1):
            c = (l,) + tuple(x >> (level - l) for x in xs)
            if c in cells:
                return c
        return "finer"

    def _refine(self, cells, c):
        if c not in cells:
            return cells
        # 2:1 balance across faces, edges and vertices: every neighbour must be at least as fine as c
        for delta in itertools.product((-1, 0, 1), repeat=self.dim):
            if not any(delta) or sum(1 for d in delta if d) > self.max_nonzero:
                continue
            nb = self._find_active(cells, c[0], tuple(x + d for x, d in zip(c[1:], delta)))
            while nb not in (None, "finer") and nb[0] < c[0]:
                cells = self._refine(cells, nb)
                nb = self._find_active(cells, c[0], tuple(x + d for x, d in zip(c[1:], delta)))
        cells = set(cells)
        cells.remove(c)
        cells.update(self._children(c))
        return cells

    def _build_lookup(self):
        self.cellset = set(self.cells)
        self.cell_index = {c: i for i, c in enumerate(self.cells)}

    def _size(self, c):
        return 1 << (self.lmax - c[0])           # edge length in finest-level units

    def _origin(self, c):
        s = self._size(c)
        return tuple(x * s for x in c[1:])

    # ---- DoFs -----------------------------------------------------------------------------------------
    def _entity_of(self, c, li):
        """entity key and position of local lexicographic dof li inside that entity"""
        dim, p, n = self.dim, self.p, self.n
        idx = [(li // n ** d) % n for d in range(dim)]
        s, o = self._size(c), self._origin(c)
        free = [d for d in range(dim) if 0 < idx[d] < p]          # directions the entity extends along
        base = tuple(o[d] + (s if idx[d] == p else 0) for d in range(dim))
        base = tuple(base[d] if d not in free else o[d] for d in range(dim))
        key = (base, tuple(free), s if free else 0)
        pos = tuple(idx[d] - 1 for d in free)
        return key, pos

    def _number_dofs(self):
        dim, p, n = self.dim, self.p, self.n
        npc = n ** dim
        h2l = hier_to_lex(dim, p)
        number = {}
        nxt = 0
        self.l2g_own = np.zeros((self.n_cells, npc), dtype=np.uint32)
        for ci, c in enumerate(self.cells):
            for li in h2l:                                      # hierarchic order inside the cell
                key = self._entity_of(c, int(li))
                if key not in number:
                    number[key] = nxt
                    nxt += 1
                self.l2g_own[ci, li] = number[key]
        self.n_dofs = nxt
        self.dof_key = number
        # support point of every DoF (finest-level lattice units scaled by p for exact integer arithmetic is not
        # possible with GLL nodes -> floating point in units of the finest cell)
        _, _, xn, _, _ = shape_1d(p)
        self.support = np.zeros((self.n_dofs, dim))
        for ci, c in enumerate(self.cells):
            s, o = self._size(c), self._origin(c)
            for li in range(npc):
                idx = [(li // n ** d) % n for d in range(dim)]
                self.support[self.l2g_own[ci, li]] = [o[d] + s * xn[idx[d]] for d in range(dim)]

    # ---- hanging nodes: masks + loc2glob rewrite (hanging_nodes.cuh:209-454) -----------------------------
    def _hanging_setup(self):
        dim, p, n = self.dim, self.p, self.n
        npc = n ** dim
        self.l2g = self.l2g_own.copy()
        self.mask = np.zeros(self.n_cells, dtype=np.uint32)
        lat = lambda idx: sum(idx[d] * n ** d for d in range(dim))
        hanging = set()
        for ci, c in enumerate(self.cells):
            l, xs = c[0], c[1:]
            mask = 0
            for d in range(dim):
                for side in (0, 1):
                    nbx = list(xs); nbx[d] += -1 if side == 0 else 1
                    nb = self._find_active(self.cellset, l, tuple(nbx))
                    if nb in (None, "finer") or nb[0] >= l:
                        continue
                    # only the outer face of a child can see a coarser neighbour
                    assert (xs[d] & 1) == side
                    mask |= CONSTR_FACE[d]
                    nbi = self.cell_index[nb]
                    others = [a for a in range(dim) if a != d]
                    for t in itertools.product(range(n), repeat=dim - 1):
                        mine = [0] * dim; theirs = [0] * dim
                        mine[d] = 0 if side == 0 else p
                        theirs[d] = p if side == 0 else 0         # the neighbour's opposite face
                        for a, ta in zip(others, t):
                            mine[a] = ta; theirs[a] = ta
                        hanging.add(int(self.l2g_own[ci, lat(mine)]))
                        self.l2g[ci, lat(mine)] = self.l2g_own[nbi, lat(theirs)]
            if dim == 3:
                for along in range(3):                           # edges running along `along`
                    a1, a2 = (along + 1) % 3, (along + 2) % 3
                    if mask & (CONSTR_FACE[a1] | CONSTR_FACE[a2]):
                        continue                                 # already part of a constrained face (:371)
                    s1, s2 = xs[a1] & 1, xs[a2] & 1              # the outer edge of this child
                    coarse = None
                    for d1, d2 in ((-1, -1), (-1, 0), (0, -1)):
                        nbx = list(xs)
                        nbx[a1] += (d1 if s1 == 0 else -d1)
                        nbx[a2] += (d2 if s2 == 0 else -d2)
                        nb = self._find_active(self.cellset, l, tuple(nbx))
                        if nb not in (None, "finer") and nb[0] < l:
                            coarse = nb
                            break
                    if coarse is None:
                        continue
                    mask |= EDGE_BIT_ALONG[along]
                    nbi = self.cell_index[coarse]
                    # the coarse cell's edge that contains ours: compare positions in finest units
                    so, ss = self._origin(coarse), self._size(coarse)
                    mo, ms = self._origin(c), self._size(c)
                    for t in range(n):
                        mine = [0, 0, 0]; theirs = [0, 0, 0]
                        mine[along] = t; theirs[along] = t
                        for a, sd in ((a1, s1), (a2, s2)):
                            mine[a] = 0 if sd == 0 else p
                            pos = mo[a] + (0 if sd == 0 else ms)
                            assert pos in (so[a], so[a] + ss)
                            theirs[a] = 0 if pos == so[a] else p
                        hanging.add(int(self.l2g_own[ci, lat(mine)]))
                        self.l2g[ci, lat(mine)] = self.l2g_own[nbi, lat(theirs)]
            if mask:
                for a in range(dim):
                    if (xs[a] & 1) == 0:
                        mask |= CONSTR_TYPE[a]
            self.mask[ci] = mask
        # a "hanging" candidate that is still referenced through some cell's rewritten map is a real coarse DoF
        referenced = set(np.unique(self.l2g).tolist())
        self.hanging = np.array(sorted(h for h in hanging if h not in referenced), dtype=np.uint32)
        # Dirichlet boundary
        S = 1 << self.lmax
        onb = np.any((np.abs(self.support) < 1e-12) | (np.abs(self.support - S) < 1e-12), axis=1)
        self.boundary = np.nonzero(onb)[0].astype(np.uint32)
        # ConstraintHandlerGpu list: every constrained DoF, ascending (constraint_handler_gpu.cu:77-83)
        self.constrained = np.unique(np.concatenate([self.hanging, self.boundary])).astype(np.uint32)
        self.is_constrained = np.zeros(self.n_dofs, dtype=bool)
        self.is_constrained[self.constrained] = True

    # ---- geometry / operator ---------------------------------------------------------------------------
    def _geometry(self):
        dim, p, n = self.dim, self.p, self.n
        self.sv, self.sg, self.xn, self.xq, self.wq = shape_1d(p)
        H = (self.right - self.left)
        self.h = np.array([H / (1 << c[0]) for c in self.cells])
        self.inv_jac = 1.0 / self.h
        npc = n ** dim
        idx = np.array([[(q // n ** d) % n for d in range(dim)] for q in range(npc)])
        self.coef = np.zeros((self.n_cells, npc))
        for ci, c in enumerate(self.cells):
            x = [self.left + self.h[ci] * (c[1 + d] + self.xq[idx[:, d]]) for d in range(dim)]
            self.coef[ci] = 1.0 / (0.05 + 2.0 * sum(xx * xx for xx in x))
        # dense reference-cell gradient tables B[d][i][q] (without 1/h) and weights
        B = np.zeros((dim, npc, npc)); w = np.ones(npc)
        for d in range(dim):
            t = np.ones((npc, npc))
            for e in range(dim):
                M = self.sg if e == d else self.sv
                t *= M[idx[:, e][:, None], idx[:, e][None, :]]
            B[d] = t
        for e in range(dim):
            w *= self.wq[idx[:, e]]
        self.B, self.wref = B, w

    def cell_matrix(self, ci):
        h, dim = self.h[ci], self.dim
        cw = self.coef[ci] * self.wref * h ** dim / h ** 2
        return sum((self.B[d] * cw[None, :]) @ self.B[d].T for d in range(dim))

    def vmult(self, src):
        """matrix-free semantics of the reference with MATRIX_FREE_HANGING_NODES: gather through the rewritten map,
        interpolate, cell operator, transposed interpolation, scatter; identity on constrained rows."""
        dim, n = self.dim, self.n
        dst = np.zeros(self.n_dofs)
        shape = (n,) * dim
        for ci in range(self.n_cells):
            row = self.l2g[ci]
            u = np.where(self.is_constrained[row], 0.0, src[row]).reshape(shape)
            u = resolve_hanging_nodes(u, int(self.mask[ci]), self.p, dim, transpose=False)
            v = (self.cell_matrix(ci) @ u.ravel()).reshape(shape)
            v = resolve_hanging_nodes(v, int(self.mask[ci]), self.p, dim, transpose=True).ravel()
            ok = ~self.is_constrained[row]
            np.add.at(dst, row[ok], v[ok])
        dst[self.constrained] += src[self.constrained]
        return dst

    def inverse_diagonal(self):
        """compute_diagonal with hanging nodes (laplace_operator_gpu.h:355-421): local diagonals pass through the
        transposed interpolation of distribute_local_to_global, constrained entries -> 1, then inverted."""
        dim, n = self.dim, self.n
        diag = np.zeros(self.n_dofs)
        for ci in range(self.n_cells):
            row = self.l2g[ci]
            d = np.diag(self.cell_matrix(ci)).copy().reshape((n,) * dim)
            d = resolve_hanging_nodes(d, int(self.mask[ci]), self.p, dim, transpose=True).ravel()
            ok = ~self.is_constrained[row]
            np.add.at(diag, row[ok], d[ok])
        diag[self.constrained] = 1.0
        return 1.0 / diag

    # ---- independent check: geometric constraint matrix ----------------------------------------------------
    def constraint_matrix(self):
        """C (n_dofs x n_dofs): identity on unconstrained DoFs, zero rows for Dirichlet DoFs, and for every hanging
        DoF the coarse neighbour's shape functions evaluated at its support point (no masks, no sweeps)."""
        dim, p, n = self.dim, self.p, self.n
        npc = n ** dim
        C = np.eye(self.n_dofs)
        C[self.boundary] = 0.0
        done = set()
        hanging_set = set(self.hanging.tolist())
        for ci, c in enumerate(self.cells):
            for li in range(npc):
                g = int(self.l2g_own[ci, li])
                if g in done or g not in hanging_set:
                    continue
                # find a coarser active cell whose closure contains the support point
                x = self.support[g]
                for cj, cc in enumerate(self.cells):
                    if cc[0] >= c[0]:
                        continue
                    o, s = np.array(self._origin(cc)), self._size(cc)
                    if np.all(x >= o - 1e-12) and np.all(x <= o + s + 1e-12):
                        xi = (x - o) / s
                        row = np.zeros(self.n_dofs)
                        for lj in range(npc):
                            jj = [(lj // n ** d) % n for d in range(dim)]
                            val = 1.0
                            for d in range(dim):
                                val *= np.prod([(xi[d] - self.xn[m]) / (self.xn[jj[d]] - self.xn[m]) for m in range(n) if m != jj[d]])
                            if abs(val) > 1e-14:
                                row[self.l2g_own[cj, lj]] += val
                        C[g] = row
                        done.add(g)
                        break
        for _ in range(4):                                   # resolve chains (hanging DoF depending on hanging DoF)
            C2 = C.copy()
            for g in self.hanging:
                C2[g] = C[g] @ C
            if np.allclose(C2, C, atol=1e-15):
                break
            C = C2
        C[self.boundary] = 0.0
        return C

    def assembled_vmult(self, src):
        A = np.zeros((self.n_dofs, self.n_dofs))
        for ci in range(self.n_cells):
            row = self.l2g_own[ci]
            A[np.ix_(row, row)] += self.cell_matrix(ci)
        C = self.constraint_matrix()
        u = src.copy(); u[self.constrained] = 0.0
        y = C.T @ (A @ (C @ u))
        y[self.constrained] = src[self.constrained]
        return y
