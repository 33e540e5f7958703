import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# tests/late_gpu/ holds GPU tests that have not run on hardware yet; tests/test_z_late_gpu_additions.py runs them in child processes
collect_ignore = [] if os.environ.get("MFG_RUN_LATE_GPU") else ["late_gpu"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


# MFG_EMULATION=1 (set by tests/test_emulated_library.py for its child process): `dealii_cuda_b200` on sys.path is a copy of the
# binding next to libmfgpu_emu.so, the library's own sources compiled for the CPU (tests/emu/): GPU tests of paths that need no
# fast kernel then run without a device
EMULATION = os.environ.get("MFG_EMULATION") == "1"


def _has_cuda():
    if EMULATION:
        return True
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ctx():
    import dealii_cuda_b200 as mf
    if not EMULATION:
        import torch
        torch.cuda.init()
    return mf.Context(0, None)


@pytest.fixture(scope="session")
def emu(tmp_path_factory):
    """libmfgpu_emu.so (the library's own sources compiled for the CPU, tests/emu/build_emu_lib.py), the examples built against it and
    a copy of the Python binding next to it: {"so", "examples", "pkg"}"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu_lib
    out = str(tmp_path_factory.mktemp("mfg_emu"))
    so = build_emu_lib.build(out)
    return {"so": so, "examples": build_emu_lib.build_examples(out, so), "pkg": build_emu_lib.build_package(out, so)}
