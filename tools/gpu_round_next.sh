#!/bin/bash
# First GPU-box pass of the next round: everything that was built in round 2 after the GPU budget was spent.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round_next.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
# BASELINE configs[3]: pseudo-adaptive mesh with hanging nodes, Q4 and Q3, apply + Jacobi-CG + MG-CG
python bench.py --adaptive --refine 6 --steps 200 --warmup 5 > gpurun_out/bench_adaptive_q4.json 2> gpurun_out/bench_adaptive_q4.err
python bench.py --adaptive --refine 6 --degree 3 --steps 200 --warmup 5 > gpurun_out/bench_adaptive_q3.json 2> gpurun_out/bench_adaptive_q3.err
# the assembled-matrix competitor row (bmop_spm.cu)
python bench.py --spmv --refine 4 --steps 50 --warmup 3 > gpurun_out/bench_spmv_q4_r4.json 2> gpurun_out/bench_spmv.err
python bench.py --spmv --refine 5 --degree 2 --steps 50 --warmup 3 > gpurun_out/bench_spmv_q2_r5.json 2>> gpurun_out/bench_spmv.err
examples/_build/bmop_ball 4 2 > gpurun_out/bmop_ball.txt 2>&1
examples/_build/bmop_adaptive 6 5 > gpurun_out/bmop_adaptive.txt 2>&1
examples/_build/bmop_adaptive 6 6 mg >> gpurun_out/bmop_adaptive.txt 2>&1
cat gpurun_out/bench_adaptive_q4.json gpurun_out/bench_spmv_q4_r4.json gpurun_out/bmop_adaptive.txt
# multigrid over the box partition: all boxes on one GPU (C++ facade), then one box per rank in the N > 1 bench lines (mg_solve)
examples/_build/partitioned_mg 8 4 > gpurun_out/partitioned_mg.txt 2>&1
examples/_build/partitioned_mg 8 6 strong >> gpurun_out/partitioned_mg.txt 2>&1
cat gpurun_out/partitioned_mg.txt
