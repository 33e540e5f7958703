"""Runs the GPU tests of tests/late_gpu/ -- written after the round's GPU budget was spent, never run on hardware -- one test
function per CHILD pytest process.  A failure, a hang (timeout) or a crash of the process in there is one failed test here and
cannot mask or take down the hardware-verified suite in front of it (this file sorts last)."""
import ast
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LATE = os.path.join(ROOT, "tests", "late_gpu", "test_late_gpu_additions.py")


def late_test_names():
    tree = ast.parse(open(LATE).read())
    return [n.name for n in tree.body if isinstance(n, ast.FunctionDef) and n.name.startswith("test_")]


def run_child(args, timeout):
    env = dict(os.environ, MFG_RUN_LATE_GPU="1")
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider"] + args, cwd=ROOT, env=env, capture_output=True, text=True,
                          timeout=timeout)


def test_late_gpu_tests_are_collectable():
    """(CPU) the late file imports and collects: syntax, fixtures and parametrisations are in order"""
    r = run_child([LATE, "--collect-only"], 600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert len(late_test_names()) >= 8 and "test_adaptive_multigrid_vcycle_and_cg" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", late_test_names())
def test_late(name):
    r = run_child([LATE + "::" + name, "-x", "-m", "gpu"], 1500)
    assert r.returncode == 0, "late GPU test %s (first run on hardware) failed:\n%s" % (name, (r.stdout + r.stderr)[-4000:])
