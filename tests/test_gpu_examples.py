"""The C++ facade (include/dealii_cuda_b200/*.h) through the compiled examples: the reference's bmop driver and the
MGTransferMatrixFreeGpu facade class (plain g++ host code calling libmfgpu.so)."""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BMOP = os.path.join(ROOT, "examples", "_build", "bmop")


def test_bmop_driver_runs():
    assert os.path.exists(BMOP), "examples/_build/bmop is missing: run __graft_entry__.build()"
    out = subprocess.run([BMOP, "3", "2"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [l.split() for l in out.stdout.strip().splitlines()]
    assert [int(r[2]) for r in rows] == [4913, 35937]        # (4 * 2^r + 1)^3 DoFs for Q4, r = 2, 3
    assert all(float(r[3]) > 0 for r in rows)


def test_mg_transfer_facade_is_adjoint():
    out = subprocess.run([BMOP, "1", "0", "mg"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    a, b = map(float, re.findall(r"= ([-0-9.e+]+)", out.stdout))
    assert abs(a - b) <= 1e-10 * abs(a) and a > 0
