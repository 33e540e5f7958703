"""GPU tests of GpuVector BLAS-1 and ConstraintHandlerGpu through the C ABI, following
test_gpu_vec.cpp:29-111 (N=5 and N=10000, srand48(0)-style random vectors) against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-14), (np.float32, 1e-6)])
@pytest.mark.parametrize("n", [0, 1, 5, 10000, 1024000 + 3])
def test_blas1(ctx, n, dtype, tol):
    import dealii_cuda_b200 as mf
    rng = np.random.default_rng(0)
    a, b, c = (rng.random(n).astype(dtype) for _ in range(3))
    va, vb, vc = (mf.GpuVector.from_numpy(ctx, x) for x in (a, b, c))
    assert va.size() == n
    ref = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    assert abs(va.dot(vb) - ref) <= tol * max(1.0, abs(ref))
    assert abs(va.l2_norm() - np.linalg.norm(a.astype(np.float64))) <= tol * max(1.0, np.sqrt(n))
    va.sadd(2.0, 3.0, vb); a = (2.0 * a + 3.0 * b).astype(dtype)
    assert np.allclose(va.toVector(), a, rtol=10 * tol, atol=0)
    va *= 0.5; a = (a * dtype(0.5)).astype(dtype)
    assert np.allclose(va.toVector(), a, rtol=10 * tol, atol=0)
    r = va.add_and_dot(-1.5, vb, vc); a = (a - 1.5 * b).astype(dtype)
    ref = float(np.dot(a.astype(np.float64), c.astype(np.float64)))
    assert np.allclose(va.toVector(), a, rtol=10 * tol, atol=10 * tol)
    assert abs(r - ref) <= 10 * tol * max(1.0, abs(ref))
    vb.equ(4.0, vc); assert np.allclose(vb.toVector(), 4 * c, rtol=10 * tol)
    vb.scale(vc); assert np.allclose(vb.toVector(), 4 * c * c, rtol=10 * tol)
    if n:
        vc.fill(2.0); vb /= vc; assert np.allclose(vb.toVector(), 2 * c * c, rtol=10 * tol)
        vc.invert(); assert np.allclose(vc.toVector(), 0.5)
    assert va.all_zero() == (n == 0 or not np.any(a))
    z = mf.GpuVector(ctx, n, dtype)
    assert z.all_zero()  # GpuVector(n) zero-fills (gpu_vec.cu:33-37)
    if n:
        z.fill(0.1); assert not z.all_zero()
    # deep copy, type conversion, swap
    w = mf.GpuVector(ctx, 0, np.float64 if dtype == np.float32 else np.float32)
    w.assign(va)
    assert w.size() == n and np.allclose(w.toVector(), a, rtol=1e-5, atol=1e-6)
    x, y = mf.GpuVector.from_numpy(ctx, a), mf.GpuVector.from_numpy(ctx, b)
    px, py = x.getData(), y.getData()
    x.swap(y)
    assert (x.getData(), y.getData()) == (py, px) or n == 0
    assert np.array_equal(x.toVector(), b) and np.array_equal(y.toVector(), a)


def test_dot_is_deterministic(ctx):
    import dealii_cuda_b200 as mf
    rng = np.random.default_rng(1)
    a = mf.GpuVector.from_numpy(ctx, rng.standard_normal(3_000_001))
    vals = {a.dot(a) for _ in range(5)}
    assert len(vals) == 1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_constraint_handler(ctx, dtype):
    """constraint_handler_gpu.cu:127-200 semantics."""
    import dealii_cuda_b200 as mf
    n = 1000
    rng = np.random.default_rng(2)
    idx = np.sort(rng.choice(n, 137, replace=False)).astype(np.uint32)
    edge = idx[::3].copy()
    ch = mf.ConstraintHandlerGpu(ctx, dtype)
    ch.reinit(idx, n, edge_indices=edge)
    assert ch.n_constrained() == idx.size
    a, b = rng.random(n).astype(dtype), rng.random(n).astype(dtype)
    va, vb = mf.GpuVector.from_numpy(ctx, a), mf.GpuVector.from_numpy(ctx, b)
    ch.save_constrained_values(va, vb)          # tmp_dst = a[c]; tmp_src = b[c]; b[c] = 0
    z = vb.toVector(); assert np.all(z[idx] == 0); mask = np.ones(n, bool); mask[idx] = False
    assert np.array_equal(z[mask], b[mask]) and np.array_equal(va.toVector(), a)
    ch.load_and_add_constrained_values(va, vb)  # a[c] = tmp_dst + tmp_src ; b[c] = tmp_src
    exp = a.copy(); exp[idx] = a[idx] + b[idx]
    assert np.array_equal(va.toVector(), exp) and np.array_equal(vb.toVector(), b)
    ch.save_constrained_values(vb)
    assert np.all(vb.toVector()[idx] == 0)
    ch.load_constrained_values(vb)
    assert np.array_equal(vb.toVector(), b)
    ch.set_constrained_values(vb, 7.0)
    exp = b.copy(); exp[idx] = 7.0
    assert np.array_equal(vb.toVector(), exp)
    d = mf.GpuVector(ctx, n, dtype)
    ch.copy_edge_values(d, va)
    exp = np.zeros(n, dtype); exp[edge] = va.toVector()[edge]
    assert np.array_equal(d.toVector(), exp)
    # empty constraint set: all operations are no-ops
    ch0 = mf.ConstraintHandlerGpu(ctx, dtype)
    ch0.reinit(np.zeros(0, np.uint32), n)
    ch0.save_constrained_values(va, vb); ch0.load_and_add_constrained_values(va, vb); ch0.set_constrained_values(va, 1.0)
