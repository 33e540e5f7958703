"""The reference's Poisson driver on the drop-in (examples/poisson.cu: precompiled LaplaceOperatorGpu + user-written
right-hand-side-with-lifting and error functors on the generic FEEvaluationGpu path + mfg_solver_cg).
Known-answer test of SURVEY 8c (8): against the analytic solution of poisson_common.cc the L2 error falls like h^(p+1),
i.e. the squared error by 2^(-2(p+1)) per global refinement."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dim,p,rmin,rmax", [(2, 2, 3, 6), (2, 4, 2, 5), (3, 2, 2, 4), (3, 4, 2, 4)])
def test_l2_error_converges_at_the_optimal_rate(dim, p, rmin, rmax):
    exe = os.path.join(ROOT, "examples", "_build", "poisson")
    assert os.path.exists(exe), "examples/_build/poisson is missing: run __graft_entry__.build()"
    out = subprocess.run([exe, str(dim), str(p), str(rmin), str(rmax)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert len(rows) == rmax - rmin + 1
    errs = [float(r[5]) for r in rows]
    its = [int(r[4]) for r in rows]
    for a, b in zip(errs[:-1], errs[1:]):
        assert 0.6 * 2 ** (p + 1) <= a / b <= 1.6 * 2 ** (p + 1), (errs, "expected a factor 2^(p+1) per refinement")
    # Jacobi-preconditioned CG: iterations roughly double per refinement (condition number ~ h^-2)
    for a, b in zip(its[:-1], its[1:]):
        assert 1.5 <= b / a <= 2.6, its
