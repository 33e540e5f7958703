"""TEST INFRASTRUCTURE: python run_with_cpu_shim.py <script> [args...] -- runs a torch-side script (bench.py, the torchrun workers of the
GPU tests) on the CPU: torch.cuda shimmed (torch_cpu_shim.py), dealii_cuda_b200 bound to libmfgpu_emu.so (the package copy must be in
front on PYTHONPATH; it is imported here first, so that a script that puts the repository in front of sys.path still gets it)."""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch_cpu_shim  # noqa: E402

torch_cpu_shim.install()
import dealii_cuda_b200  # noqa: E402,F401

assert "emu" in dealii_cuda_b200._capi.LIB_PATH or os.path.islink(dealii_cuda_b200._capi.LIB_PATH), dealii_cuda_b200._capi.LIB_PATH
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(script, run_name="__main__")
