"""CPU tests of the drop-in boundary: libmfgpu.so loads without a GPU and exports every
symbol include/mfgpu.h declares; the Python binding declares the same set; and the
product fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mfg_build", os.path.join(ROOT, "dealii_cuda_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return b.build()


def header_functions():
    txt = open(os.path.join(ROOT, "include", "mfgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mfg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    so = _ensure_built()
    L = ctypes.CDLL(so)
    names = header_functions()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_python_binding_covers_header():
    _ensure_built()
    from dealii_cuda_b200 import _capi
    assert sorted(_capi.DECLARED_SYMBOLS) == header_functions()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _ensure_built()
    import dealii_cuda_b200 as mf
    with pytest.raises(mf.MfgError) as e:
        mf.Context(0)
    assert "no usable CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under dealii_cuda_b200/ may import, load or link it"""
    pkg = os.path.join(ROOT, "dealii_cuda_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)|liboracle|oracle/_build|mf_oracle|orc_[a-z_]+\(", re.M)
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert not pat.search(txt), (dirpath, f)
