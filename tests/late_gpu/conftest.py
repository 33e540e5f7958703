"""Reports every test of this directory to the file named by MFG_LATE_LOG as soon as it starts / ends (one JSON object per line,
flushed), so that tests/test_z_late_gpu_additions.py knows what ran even if the process dies in the middle."""
import json
import os


def _log(obj):
    path = os.environ.get("MFG_LATE_LOG")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(obj) + "\n")
            f.flush()
            os.fsync(f.fileno())


def pytest_runtest_logstart(nodeid, location):
    _log({"id": nodeid, "when": "start"})


def pytest_runtest_logreport(report):
    if report.when == "call" or report.outcome != "passed":   # the call result, or a setup / teardown that failed or skipped
        _log({"id": report.nodeid, "when": report.when, "outcome": report.outcome,
              "msg": str(report.longrepr)[-3000:] if report.outcome in ("failed", "skipped") else ""})
