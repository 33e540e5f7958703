#!/bin/bash
# round 2, pass A: staged kernel correctness + A/B timing against slab2 + ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_apply.py -m gpu -x -q -k "staged or variant_rejected or full_size_r6_against" 2>&1 | tail -15 > gpurun_out/a_pytest.log
cat gpurun_out/a_pytest.log
C="3,4,6,f64,9;3,4,6,f64,40;3,4,6,f32,40;3,4,5,f64,40;3,3,6,f64,40;3,5,5,f64,40;3,2,7,f64,40"
timeout 600 python tools/sweep.py --steps 50 --custom "$C" > gpurun_out/a_sweep.jsonl 2> gpurun_out/a_sweep.err
MFG_STAGE_SYNC=0 timeout 600 python tools/sweep.py --steps 50 --custom "$C" > gpurun_out/a_sweep_nosync.jsonl 2>> gpurun_out/a_sweep.err
cut -c1-200 gpurun_out/a_sweep.jsonl; echo nosync; cut -c1-200 gpurun_out/a_sweep_nosync.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:laplace_cell_stage -s 3 -c 1 -o gpurun_out/prof_stage_a -f python tools/sweep.py --steps 5 --custom "3,4,6,f64,40" > gpurun_out/a_ncu.log 2>&1
tail -2 gpurun_out/a_ncu.log
