"""Runs the GPU tests of tests/late_gpu/ -- written after the round's GPU budget was spent, never run on hardware -- in a CHILD
pytest process (all of them, no -x; a new child for the remaining ones if the process dies) and reports every test function as one
test here.  A failure, a hang (timeout) or a crash in there shows up as failed tests of this file and cannot mask or take down the
hardware-verified suite in front of it (this file sorts last)."""
import ast
import json
import os
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LATE = os.path.join(ROOT, "tests", "late_gpu", "test_late_gpu_additions.py")
CHILD_TIMEOUT_S = 720      # all late tests together run for three to five minutes (the torchrun ones only where there are GPUs for them)
MAX_CHILDREN = 3           # one, plus one per crash / hang
_results = None


def late_test_names(path=None):
    tree = ast.parse(open(path or LATE).read())
    return [n.name for n in tree.body if isinstance(n, ast.FunctionDef) and n.name.startswith("test_")]


def run_child(args, timeout, log=None):
    env = dict(os.environ, MFG_RUN_LATE_GPU="1")
    if log:
        env["MFG_LATE_LOG"] = log
    # (--rootdir: tests/conftest.py with the `ctx` fixture and the gpu marker must be found whatever file is named)
    return subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", ROOT] + args, cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=timeout)


def late_results(late=None, extra_args=("-m", "gpu")):
    """{test function: (ok, message)}: outcomes of every test id, read from the log the child writes as it goes"""
    global _results
    if _results is not None and late is None:
        return _results
    late = late or LATE
    outcome = {}                     # test id -> ("passed" | "failed" | "skipped" | "crashed", message)
    with tempfile.TemporaryDirectory() as tmp:
        for attempt in range(MAX_CHILDREN):
            log = os.path.join(tmp, "late%d.jsonl" % attempt)
            args = [late] + list(extra_args)
            for tid in outcome:
                args += ["--deselect", tid]
            try:
                r = run_child(args, CHILD_TIMEOUT_S, log)
                tail, died = (r.stdout + r.stderr)[-2500:], r.returncode not in (0, 1, 5)
            except subprocess.TimeoutExpired:
                tail, died = "child pytest process exceeded %d s" % CHILD_TIMEOUT_S, True
            started = None
            if os.path.exists(log):
                for line in open(log):
                    try:
                        e = json.loads(line)
                    except ValueError:
                        continue
                    if e["when"] == "start":
                        started = e["id"]
                    elif e["id"] not in outcome or e["outcome"] == "failed":
                        outcome[e["id"]] = (e["outcome"], e.get("msg", ""))
                        if e["id"] == started:
                            started = None
            if not died:
                break
            if started is None:
                break                # died outside a test (collection, start-up): nothing more to learn from another child
            outcome[started] = ("crashed", "the child pytest process died or hung in this test\n" + tail)
    results = {}
    for fn in late_test_names(late):
        mine = {tid: o for tid, o in outcome.items() if tid.split("::")[-1].split("[")[0] == fn}
        bad = ["%s: %s\n%s" % (tid, o[0], o[1]) for tid, o in mine.items() if o[0] in ("failed", "crashed")]
        if not mine:
            results[fn] = (False, "no result: the child process did not get to this test")
        elif bad:
            results[fn] = (False, "\n".join(bad))
        elif all(o[0] == "skipped" for o in mine.values()):
            # (a test that needs more GPUs than the box has is rightly skipped; any other skip means the child saw no device)
            results[fn] = (all("needs" in o[1] and "GPUs" in o[1] for o in mine.values()), "skipped in the child process (no CUDA device there?)")
        else:
            results[fn] = (True, "")
    if late == LATE:
        _results = results
    return results


def test_late_gpu_tests_are_collectable():
    """(CPU) the late file imports and collects: syntax, fixtures and parametrisations are in order"""
    r = run_child([LATE, "--collect-only"], 600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert len(late_test_names()) >= 8 and "test_adaptive_multigrid_vcycle_and_cg" in r.stdout


def test_wrapper_survives_failures_and_crashes(tmp_path):
    """(CPU) the machinery itself: a passing, a partly failing, a crashing test and one behind the crash"""
    sim = tmp_path / "test_sim.py"
    sim.write_text("import os, signal, pytest\n"
                   "@pytest.mark.parametrize('k', [1, 2])\ndef test_a_passes(k):\n    assert k > 0\n"
                   "@pytest.mark.parametrize('k', [1, 2])\ndef test_b_fails_once(k):\n    assert k == 1, 'k was %d' % k\n"
                   "def test_c_crashes():\n    os.kill(os.getpid(), signal.SIGSEGV)\n"
                   "def test_d_runs_in_a_second_child():\n    pass\n"
                   "def test_e_needs_gpus():\n    pytest.skip('needs 8 GPUs')\n"
                   "def test_f_skipped_otherwise():\n    pytest.skip('no device')\n")
    (tmp_path / "conftest.py").write_text(open(os.path.join(ROOT, "tests", "late_gpu", "conftest.py")).read())
    r = late_results(str(sim), extra_args=())
    assert r["test_a_passes"][0] and r["test_d_runs_in_a_second_child"][0]
    assert not r["test_b_fails_once"][0] and "k was 2" in r["test_b_fails_once"][1]
    assert not r["test_c_crashes"][0] and "died or hung" in r["test_c_crashes"][1]
    assert r["test_e_needs_gpus"][0] and not r["test_f_skipped_otherwise"][0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", late_test_names())
def test_late(name):
    res = late_results()
    ok, msg = res[name]
    # (`pytest -x` stops at the first failure: the message carries the outcome of every late test, so that nothing is lost)
    table = "\n".join("  %s  %s" % ("PASS" if res[n][0] else "FAIL", n) for n in late_test_names())
    assert ok, "late GPU test %s (first run on hardware) failed:\n%s\noutcome of all late GPU tests:\n%s" % (name, msg, table)
