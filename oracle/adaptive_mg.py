"""CPU oracle for geometric multigrid with local smoothing on adaptively refined meshes (TEST INFRASTRUCTURE, numpy, small
meshes only).  Restates, from the mesh tree alone, what the reference obtains from deal.II in poisson_mg.cu / bmop_mg.cu on
adaptive grids and what its own classes do with it:

  * level meshes = ALL cells of a level (active or refined), level DoFs by first touch over them  [DoFHandler::distribute_mg_dofs]
  * MGConstrainedDoFs: boundary indices and refinement-edge indices per level (faces of level cells whose neighbour is not
    refined to that level; MGTools::extract_inner_interface_dofs)
  * level operators with the constraint handler of constraint_handler_gpu.cu:99-123 (constrained = boundary + edge,
    identity rows), vmult_interface_down / vmult_interface_up (laplace_operator_gpu.h:306-352)
  * MGTransferMatrixFreeGpu: prolongation = the coarse FE function evaluated at the fine support points (built
    GEOMETRICALLY here, no index blocks, no weights), boundary columns of the coarse level zero (.cu:592-654);
    copy_to_mg / copy_from_mg through index pairs that skip refinement-edge DoFs (.cu:688-757, MGTransfer::fill_copy_indices)
  * Multigrid::level_v_step with edge_out = edge_in = the interface operators (poisson_mg.cu:365-375), Chebyshev
    smoothing (deal.II PreconditionChebyshev restated), exact coarse solve, PreconditionMG::vmult.

PARITY STATUS: unpinned at the deal.II boundary (no deal.II here).  What pins it: the V-cycle built this way is a symmetric
positive definite preconditioner whose CG iteration counts stay bounded under refinement (tests/test_adaptive_multigrid.py)."""
import itertools

import numpy as np

from .adaptive import AdaptiveMesh
from .oracle import shape_1d


class LevelMesh:
    """all cells of one level in storage order, level DoFs, constraint sets, dense operators"""

    def __init__(self, dim, p, level, coords, left, right):
        self.dim, self.p, self.n, self.level = dim, p, p + 1, level
        self.coords = [tuple(int(v) for v in c[:dim]) for c in coords]
        self.am = AdaptiveMesh(dim, p, 0, [], left, right, cells=[(level,) + c for c in self.coords])
        assert self.am.mask.max() == 0
        self.n_dofs, self.l2g = self.am.n_dofs, self.am.l2g_own
        self.boundary = self.am.boundary
        n, pos = self.n, set(self.coords)
        lat = lambda idx: sum(idx[d] * n ** d for d in range(dim))
        edge = set()
        for ci, c in enumerate(self.coords):
            for d in range(dim):
                for side in (0, 1):
                    nb = list(c); nb[d] += -1 if side == 0 else 1
                    if nb[d] < 0 or nb[d] >= (1 << level) or tuple(nb) in pos:
                        continue          # domain boundary, or the neighbour is refined to this level as well
                    for t in itertools.product(range(n), repeat=dim - 1):
                        idx = [0] * dim
                        idx[d] = 0 if side == 0 else p
                        for a, ta in zip([a for a in range(dim) if a != d], t):
                            idx[a] = ta
                        edge.add(int(self.l2g[ci, lat(idx)]))
        self.edge = np.array(sorted(edge), dtype=np.uint32)
        self.constrained = np.unique(np.concatenate([self.edge, self.boundary])).astype(np.uint32)
        A = np.zeros((self.n_dofs, self.n_dofs))
        for ci in range(len(self.coords)):
            row = self.l2g[ci]
            A[np.ix_(row, row)] += self.am.cell_matrix(ci)
        self.A_raw = A
        Ac = A.copy()
        Ac[self.constrained, :] = 0.0; Ac[:, self.constrained] = 0.0
        Ac[self.constrained, self.constrained] = 1.0
        self.A = Ac

    def vmult(self, x):
        return self.A @ x

    def inverse_diagonal(self):
        return 1.0 / np.diag(self.A)

    def vmult_interface_down(self, x):
        x0 = x.copy(); x0[self.constrained] = 0.0
        t = self.A_raw @ x0
        out = np.zeros_like(x); out[self.edge] = t[self.edge]
        return out

    def vmult_interface_up(self, x):
        xe = np.zeros_like(x); xe[self.edge] = x[self.edge]
        out = self.A_raw @ xe
        out[self.constrained] = 0.0
        return out


def lagrange(xn, j, x):
    return np.prod([(x - xn[m]) / (xn[j] - xn[m]) for m in range(len(xn)) if m != j])


def geometric_prolongation(coarse, fine):
    """P (n_fine x n_coarse): the coarse-level FE function at every fine-level support point; coarse boundary columns zero"""
    dim, p, n = fine.dim, fine.p, fine.n
    _, _, xn, _, _ = shape_1d(p)
    cpos = {c: i for i, c in enumerate(coarse.coords)}
    P = np.zeros((fine.n_dofs, coarse.n_dofs))
    done = np.zeros(fine.n_dofs, dtype=bool)
    for ci, c in enumerate(fine.coords):
        parent = cpos[tuple(v >> 1 for v in c)]
        for li in range(n ** dim):
            g = int(fine.l2g[ci, li])
            if done[g]:
                continue
            done[g] = True
            idx = [(li // n ** d) % n for d in range(dim)]
            xi = [((c[d] & 1) + xn[idx[d]]) / 2.0 for d in range(dim)]       # reference coordinates in the parent
            for lj in range(n ** dim):
                jj = [(lj // n ** d) % n for d in range(dim)]
                val = 1.0
                for d in range(dim):
                    val *= lagrange(xn, jj[d], xi[d])
                if abs(val) > 1e-15:
                    P[g, coarse.l2g[parent, lj]] = val
    assert done.all()
    P[:, coarse.boundary] = 0.0
    return P


class AdaptiveMultigridOracle:
    """level_cells: {level: [(x, y[, z], has_children), ...]} in storage order; active: oracle AdaptiveMesh on the active cells
    (same order as the library's)"""

    def __init__(self, dim, p, level_cells, active, left=-1.0, right=1.0, min_level=0, smoother_degree=5, smoothing_range=15.0, n_eig=15):
        self.dim, self.p, self.active, self.min_level = dim, p, active, min_level
        self.max_level = max(level_cells)
        self.levels = {l: LevelMesh(dim, p, l, [c[:dim] for c in level_cells[l]], left, right) for l in range(min_level, self.max_level + 1)}
        self.P = {l: geometric_prolongation(self.levels[l - 1], self.levels[l]) for l in range(min_level + 1, self.max_level + 1)}
        # copy indices: active cells of a level, DoFs that are not on its refinement edge
        act_index = {c: i for i, c in enumerate(active.cells)}
        self.copy = {}
        for l, lm in self.levels.items():
            is_edge = np.zeros(lm.n_dofs, dtype=bool); is_edge[lm.edge] = True
            pairs = set()
            for ci, c in enumerate(lm.coords):
                if level_cells[l][ci][-1]:
                    continue
                ai = act_index[(l,) + c]
                for li in range(lm.n ** dim):
                    lv = int(lm.l2g[ci, li])
                    if not is_edge[lv]:
                        pairs.add((int(active.l2g_own[ai, li]), lv))
            pairs = sorted(pairs)
            self.copy[l] = (np.array([a for a, _ in pairs], dtype=np.int64), np.array([b for _, b in pairs], dtype=np.int64))
        self.smoothers = {l: chebyshev(self.levels[l], smoother_degree, smoothing_range, n_eig) for l in range(min_level + 1, self.max_level + 1)}

    def copy_to_mg(self, src):
        out = {}
        for l, lm in self.levels.items():
            v = np.zeros(lm.n_dofs)
            g, lv = self.copy[l]
            v[lv] = src[g]
            out[l] = v
        return out

    def copy_from_mg(self, sol):
        dst = np.zeros(self.active.n_dofs)
        for l in sorted(self.levels):
            g, lv = self.copy[l]
            dst[g] = sol[l][lv]
        return dst

    def level_v_step(self, level, sol, defect):
        lm = self.levels[level]
        if level == self.min_level:
            sol[level] = np.linalg.solve(lm.A, defect[level])
            return
        _, _, _, apply = self.smoothers[level]
        sol[level] = apply(defect[level])                       # pre-smoothing from zero
        t = lm.vmult(sol[level])
        t += lm.vmult_interface_down(sol[level])                # edge_out->vmult_add
        t = defect[level] - t
        defect[level - 1] = defect[level - 1] + self.P[level].T @ t   # restrict_and_add
        self.level_v_step(level - 1, sol, defect)
        sol[level] = sol[level] + self.P[level] @ sol[level - 1]      # prolongate + add
        defect[level] = defect[level] - lm.vmult_interface_up(sol[level])   # edge_in->Tvmult
        sol[level] = apply(defect[level], sol[level])           # post-smoothing

    def vmult(self, src):
        """PreconditionMG::vmult"""
        defect = self.copy_to_mg(src)
        sol = {l: np.zeros(lm.n_dofs) for l, lm in self.levels.items()}
        self.level_v_step(self.max_level, sol, defect)
        return self.copy_from_mg(sol)

    def matrix(self):
        """the V-cycle as a dense matrix on the active DoFs (for symmetry / definiteness checks)"""
        n = self.active.n_dofs
        M = np.zeros((n, n))
        for j in range(n):
            e = np.zeros(n); e[j] = 1.0
            M[:, j] = self.vmult(e)
        return M


def chebyshev(o, degree, smoothing_range, n_eig):
    """deal.II PreconditionChebyshev restated on an operator object with n_dofs / vmult / inverse_diagonal (SURVEY Appendix A.9;
    the same statement as tests/test_gpu_multigrid.py::numpy_chebyshev): returns (lambda_max, theta, delta, apply(b, x0=None))"""
    n = o.n_dofs
    dinv = o.inverse_diagonal()
    rhs = np.full(n, 1.0 / np.sqrt(n)); rhs[0] = 0.0
    g = -rhs; h = dinv * g; d = -h; gh = g @ h
    al, be = [], []
    for _ in range(n_eig):
        if np.sqrt(g @ g) <= 1e-2:
            break
        h = o.vmult(d)
        alpha = gh / (d @ h); al.append(alpha)
        g = g + alpha * h
        h = dinv * g
        ghn = g @ h; beta = ghn / gh; gh = ghn; be.append(beta)
        d = beta * d - h
    k = len(al)
    if k == 0:
        lmax = 1.0
    else:
        T = np.zeros((k, k))
        for j in range(k):
            T[j, j] = 1.0 / al[j] + (be[j - 1] / al[j - 1] if j else 0.0)
            if j + 1 < k:
                T[j, j + 1] = T[j + 1, j] = np.sqrt(be[j]) / al[j]
        lmax = np.linalg.eigvalsh(T)[-1]
    b_, a_ = 1.2 * lmax, lmax / smoothing_range
    theta, delta = 0.5 * (b_ + a_), 0.5 * (b_ - a_)

    def apply(b, x0=None):
        rhok, sigma = delta / theta, theta / delta
        if x0 is None:
            dvec = dinv * b / theta
            x = dvec.copy()
        else:
            dvec = dinv * (b - o.vmult(x0)) / theta
            x = x0 + dvec
        for _ in range(degree):
            r = b - o.vmult(x)
            rhokp = 1.0 / (2.0 * sigma - rhok)
            dvec = rhokp * rhok * dvec + 2.0 * rhokp / delta * (dinv * r)
            rhok = rhokp
            x = x + dvec
        return x
    return lmax, theta, delta, apply


def cg_preconditioned(vmult, precond, b, tol, max_iter=200):
    """SolverCG control flow (poisson.cu:233-260 / SURVEY Appendix A.9): returns (x, iterations, residual history)"""
    x = np.zeros_like(b)
    g = -b.copy()
    hist = [np.linalg.norm(g)]
    if hist[0] <= tol:
        return x, 0, hist
    h = precond(g); d = -h; gh = g @ h
    for it in range(1, max_iter + 1):
        h = vmult(d)
        alpha = gh / (d @ h)
        x = x + alpha * d
        g = g + alpha * h
        hist.append(np.linalg.norm(g))
        if hist[-1] <= tol:
            return x, it, hist
        h = precond(g)
        ghn = g @ h; beta = ghn / gh; gh = ghn
        d = beta * d - h
    return x, max_iter, hist
