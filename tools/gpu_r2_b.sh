#!/bin/bash
# round 2, pass B: slab3 kernel correctness + A/B timing + ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_apply.py -m gpu -x -q -k "slab3 or grouped or variant_rejected or staged or full_size_r6_against" 2>&1 | tail -15 > gpurun_out/b_pytest.log
cat gpurun_out/b_pytest.log
C="3,4,6,f64,9;3,4,6,f64,50;3,4,6,f32,9;3,4,6,f32,50;3,4,5,f64,9;3,4,5,f64,50;3,3,6,f64,9;3,3,6,f64,50;3,5,5,f64,9;3,5,5,f64,50;3,2,7,f64,2;3,2,7,f64,50;3,1,8,f64,9;3,1,8,f64,50;3,4,7,f64,50"
timeout 600 python tools/sweep.py --steps 50 --custom "$C" > gpurun_out/b_sweep.jsonl 2> gpurun_out/b_sweep.err
cut -c1-200 gpurun_out/b_sweep.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_cell_slab3 -s 3 -c 1 -o gpurun_out/prof_slab3_b -f python tools/sweep.py --steps 5 --custom "3,4,6,f64,50" > gpurun_out/b_ncu.log 2>&1
tail -2 gpurun_out/b_ncu.log
