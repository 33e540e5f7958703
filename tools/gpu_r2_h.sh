#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_solver.py tests/test_gpu_poisson.py -m gpu -x -q 2>&1 | tail -12
timeout 600 python tools/solve_large.py --mg-only --mg7 2>&1 | tail -12
