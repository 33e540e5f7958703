"""Hanging-node support (BASELINE config 4).
CPU: the reference's exact-zero known-answer test (test_hanging_node_interpolation.cu:311-350: a linear polynomial
given in the COARSE parametrisation on the constrained faces is reproduced exactly) for every face / type / edge mask,
2D and 3D, p = 1..4, on the oracle's restatement of resolve_hanging_nodes; the mask-driven operator against an
independent geometric-constraint operator C^T A C.  GPU: the CUDA path (column kernel with fused interpolation) against
the oracle on adaptive meshes."""
import itertools

import numpy as np
import pytest

from oracle.adaptive import (CONSTR_FACE, CONSTR_TYPE, EDGE_BIT_ALONG, AdaptiveMesh, constraint_weights, resolve_hanging_nodes)
from oracle.oracle import shape_1d, sm64


def all_masks(dim):
    for types in range(1 << dim):
        for faces in range(1 << dim):
            edges = [0]
            if dim == 3:
                free = [d for d in range(3) if not (faces >> ((d + 1) % 3)) & 1 and not (faces >> ((d + 2) % 3)) & 1]
                edges = [sum(EDGE_BIT_ALONG[d] for d in sub) for k in range(len(free) + 1) for sub in itertools.combinations(free, k)]
            for e in edges:
                m = types | (faces << 3) | e
                if (m >> 3) == 0:
                    continue
                yield m


def coarse_filled_cell(mask, p, dim, poly):
    """cell tensor of `poly` at the fine support points, with the constrained faces / edges holding the values of the
    coarse parametrisation (setup_values, test_hanging_node_interpolation.cu:85-207)"""
    n = p + 1
    _, _, xn, _, _ = shape_1d(p)
    vals = np.zeros((n,) * dim)
    for idx in itertools.product(range(n), repeat=dim):      # idx = (x, y[, z])
        outer = [(idx[a] == 0) if (mask & CONSTR_TYPE[a]) else (idx[a] == p) for a in range(dim)]
        coarse_axes = set()
        for d in range(dim):
            if (mask & CONSTR_FACE[d]) and outer[d]:
                coarse_axes |= {a for a in range(dim) if a != d}
        if dim == 3:
            for d in range(3):
                a1, a2 = (d + 1) % 3, (d + 2) % 3
                if (mask & EDGE_BIT_ALONG[d]) and outer[a1] and outer[a2]:
                    coarse_axes.add(d)
        x = []
        for a in range(dim):
            if a in coarse_axes:
                x.append(2 * xn[idx[a]] if (mask & CONSTR_TYPE[a]) else 2 * xn[idx[a]] - 1)
            else:
                x.append(xn[idx[a]])
        vals[tuple(reversed(idx))] = poly(x)
    return vals


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_known_answer_linear_polynomial_all_masks(dim, p):
    n = p + 1
    _, _, xn, _, _ = shape_1d(p)
    poly = lambda x: x[0] + 2 * x[1] + (3 * x[2] if dim == 3 else 0)
    exact = np.zeros((n,) * dim)
    for idx in itertools.product(range(n), repeat=dim):
        exact[tuple(reversed(idx))] = poly([xn[i] for i in idx])
    count = 0
    for mask in all_masks(dim):
        got = resolve_hanging_nodes(coarse_filled_cell(mask, p, dim, poly), mask, p, dim)
        assert np.abs(got - exact).max() <= 1e-14, (dim, p, mask)
        count += 1
    assert count >= (12 if dim == 2 else 100)


def test_weights_partition_of_unity_and_library_table():
    import dealii_cuda_b200 as mf
    for p in range(1, 9):
        W = constraint_weights(p)
        assert np.allclose(W.sum(axis=1), 1.0, atol=1e-14)
        assert W[0, 0] == 1.0 and abs(W[p].sum() - 1) < 1e-14
        assert np.allclose(mf.hanging_node_weights(p), W, atol=1e-15)   # host-only entry point of the C ABI


def test_transpose_is_adjoint():
    rng = np.random.default_rng(0)
    for dim, p in [(2, 3), (3, 2)]:
        for mask in list(all_masks(dim))[::7]:
            a, b = rng.random(((p + 1),) * dim), rng.random(((p + 1),) * dim)
            lhs = np.vdot(resolve_hanging_nodes(a, mask, p, dim), b)
            rhs = np.vdot(a, resolve_hanging_nodes(b, mask, p, dim, transpose=True))
            assert abs(lhs - rhs) <= 1e-13 * abs(lhs)


ADAPTIVE_CASES = [
    (2, 1, 2, "disk"), (2, 2, 2, "disk"), (2, 4, 2, "offset"), (2, 3, 3, "disk"),
    (3, 1, 1, "corner"), (3, 2, 1, "corner"), (3, 2, 2, "ball"), (3, 3, 1, "corner"), (3, 4, 1, "corner"),
    (3, 2, 1, "lshape"), (3, 4, 1, "lshape"),   # fine cells that touch a coarse cell along an edge only
]


def make_mesh(dim, p, base, kind):
    crit = {"disk": lambda c, h: np.linalg.norm(c) < 0.5, "offset": lambda c, h: np.linalg.norm(c - 0.3) < 0.6,
            "corner": lambda c, h: bool(np.all(c < 0)), "ball": lambda c, h: np.linalg.norm(c) < 0.6,
            "lshape": lambda c, h: not (c[0] > 0 and c[1] > 0)}[kind]
    steps = [crit] if kind != "corner" else [crit, lambda c, h: bool(np.all(c < -0.5))]
    return AdaptiveMesh(dim, p, base, steps)


@pytest.mark.parametrize("dim,p,base,kind", ADAPTIVE_CASES)
def test_mask_operator_equals_geometric_constraints(dim, p, base, kind):
    m = make_mesh(dim, p, base, kind)
    assert len(m.hanging) > 0 and m.mask.max() > 0
    u = sm64(1, m.n_dofs)
    a, b = m.vmult(u), m.assembled_vmult(u)
    assert np.linalg.norm(a - b) <= 1e-13 * np.linalg.norm(a)
    assert np.array_equal(a[m.constrained], u[m.constrained])


def test_edge_only_masks_occur():
    m = make_mesh(3, 2, 1, "lshape")
    assert any(int(x) & (64 | 128 | 256) for x in m.mask), "no edge-only constraint in the test mesh"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("dim,p,base,kind", ADAPTIVE_CASES + [(2, 8, 1, "disk"), (3, 5, 1, "corner")])
def test_gpu_hanging_node_apply(ctx, dim, p, base, kind, dtype):
    import dealii_cuda_b200 as mf
    m = make_mesh(dim, p, base, kind)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    arrays = dict(dim=dim, degree=p, n_dofs=m.n_dofs, loc2glob=m.l2g, inv_jac=m.inv_jac, constraint_mask=m.mask)
    data = mf.MatrixFreeGpu(ctx, dtype)
    data.reinit(arrays)
    ch = mf.ConstraintHandlerGpu(ctx, dtype)
    ch.reinit(m.constrained, m.n_dofs)
    op = mf.LaplaceOperatorGpu(ctx, dtype)
    op.reinit(data, ch, coefficient=m.coef)
    u = sm64(3, m.n_dofs).astype(dtype)
    src, dst = mf.GpuVector.from_numpy(ctx, u), mf.GpuVector(ctx, m.n_dofs, dtype)
    op.vmult(dst, src)
    got, want = dst.toVector().astype(np.float64), m.vmult(u.astype(np.float64))
    assert np.linalg.norm(got - want) <= tol * np.linalg.norm(want)
    assert np.array_equal(got[m.constrained], u[m.constrained].astype(np.float64))
    if dtype == np.float64:
        op.compute_diagonal()
        gd, wd = op.get_diagonal_inverse().toVector(), m.inverse_diagonal()
        assert np.linalg.norm(gd - wd) <= 1e-12 * np.linalg.norm(wd)
        # every kernel variant gives the same operator on the unconstrained cells
        for variant in ([1, 50, 40] if dim == 3 and 2 <= p <= 5 else [1]):
            op.set_variant(variant)
            op.vmult(dst, src)
            assert np.linalg.norm(dst.toVector() - want) <= tol * np.linalg.norm(want)


@pytest.mark.gpu
@pytest.mark.parametrize("dim,p,base,kind", [(2, 3, 2, "disk"), (3, 3, 1, "corner"), (3, 4, 1, "corner")])
def test_gpu_cg_solve_on_adaptive_mesh(ctx, dim, p, base, kind):
    """BASELINE.json configs[3]: apply AND solve on an adaptively refined mesh with hanging-node constraints.  mfg_solver_cg
    on the GPU operator against the same CG recurrence (poisson.cu:233-260 control flow) run in numpy on the oracle's operator:
    same iteration count, iterates equal to rounding; the solution satisfies the oracle's operator equation."""
    import dealii_cuda_b200 as mf
    m = make_mesh(dim, p, base, kind)
    arrays = dict(dim=dim, degree=p, n_dofs=m.n_dofs, loc2glob=m.l2g, inv_jac=m.inv_jac, constraint_mask=m.mask)
    data = mf.MatrixFreeGpu(ctx, np.float64)
    data.reinit(arrays)
    ch = mf.ConstraintHandlerGpu(ctx, np.float64)
    ch.reinit(m.constrained, m.n_dofs)
    op = mf.LaplaceOperatorGpu(ctx, np.float64)
    op.reinit(data, ch, coefficient=m.coef)
    ue = sm64(17, m.n_dofs)
    b = m.vmult(ue)
    tol = 1e-10 * np.linalg.norm(b)
    # numpy restatement of SolverCG with the Jacobi preconditioner on the oracle operator
    minv = m.inverse_diagonal()
    x = np.zeros(m.n_dofs); g = -b.copy(); h = minv * g; d = -h; gh = g @ h
    it_ref = 0
    hist_ref = [np.linalg.norm(g)]
    while hist_ref[-1] > tol and it_ref < 2000:
        it_ref += 1
        Ad = m.vmult(d)
        alpha = gh / (d @ Ad)
        x += alpha * d; g += alpha * Ad
        hist_ref.append(np.linalg.norm(g))
        if hist_ref[-1] <= tol:
            break
        h = minv * g
        beta = (g @ h) / gh
        gh = g @ h
        d = beta * d - h
    vb, vx = mf.GpuVector.from_numpy(ctx, b), mf.GpuVector(ctx, m.n_dofs)
    it, res, hist = mf.solver_cg(op, vx, vb, tol, 2000, use_jacobi=True, history=True)
    assert abs(it - it_ref) <= 1, (it, it_ref)
    # CG amplifies rounding differences as orthogonality is lost: tight on the first half of the iterations, loose after
    k = min(it, it_ref)
    assert np.allclose(hist[:k // 2 + 1], hist_ref[:k // 2 + 1], rtol=1e-6)
    assert np.allclose(hist[:k + 1], hist_ref[:k + 1], rtol=5e-2)
    got = vx.toVector()
    assert np.linalg.norm(m.vmult(got) - b) <= 2 * tol
    assert np.linalg.norm(got - ue) <= 1e-6 * np.linalg.norm(ue)


def _generic_path_resolve(values, mask, p, dim, transpose):
    """line-by-line transliteration of FEEvaluationGpu::resolve_hanging_nodes (include/dealii_cuda_b200/fee_gpu.cuh): one thread per
    local DoF t, sweep per direction, stride / base index arithmetic and weight-table indexing as in the device code"""
    n = p + 1
    W = constraint_weights(p).ravel()                 # W[k*n+i]
    vals = values.copy().ravel()                      # lexicographic, x fastest
    for d in range(dim):
        new = vals.copy()
        for t in range(n ** dim):
            idx = [t % n, (t // n) % n, (t // (n * n)) if dim == 3 else 0]
            if dim == 2:
                a = 1 - d
                on = (idx[a] == 0) if (mask & (1 << a)) else (idx[a] == p)
                flag = bool(mask & (8 << a)) and on
            else:
                f1, f2 = (d + 1) % 3, (d + 2) % 3
                on1 = (idx[f1] == 0) if (mask & (1 << f1)) else (idx[f1] == p)
                on2 = (idx[f2] == 0) if (mask & (1 << f2)) else (idx[f2] == p)
                edge_bit = (1 << 7) if d == 0 else (1 << 8) if d == 1 else (1 << 6)
                flag = (bool(mask & (8 << f1)) and on1) or (bool(mask & (8 << f2)) and on2) or (bool(mask & edge_bit) and on1 and on2)
            if not flag:
                continue
            stride = 1 if d == 0 else n if d == 1 else n * n
            k = idx[d]
            base = t - k * stride
            lower = (mask & (1 << d)) != 0
            acc = 0.0
            for i in range(n):
                if lower:
                    w = W[i * n + k] if transpose else W[k * n + i]
                else:
                    w = W[(p - i) * n + (p - k)] if transpose else W[(p - k) * n + (p - i)]
                acc += w * vals[base + i * stride]
            new[t] = acc
        vals = new
    return vals


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 3), (3, 1), (3, 2)])
def test_generic_path_resolve_index_arithmetic(dim, p):
    """the index arithmetic of the generic path's device function, transliterated, equals resolve_hanging_nodes for every mask"""
    rng = np.random.default_rng(0)
    n = p + 1
    for mask in range(1 << (6 if dim == 2 else 9)):
        if dim == 2 and (mask & 0b100100):
            continue
        v = rng.random((n,) * dim)
        for tr in (False, True):
            assert np.array_equal(resolve_hanging_nodes(v, mask, p, dim, transpose=tr).ravel(), _generic_path_resolve(v, mask, p, dim, tr))
