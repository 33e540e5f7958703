#!/bin/bash
# One GPU-box pass (round 2): tests, bench, reference arm, sweeps, launch list, full ncu captures of the cell kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 30 --warmup 2 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
python tools/sweep.py --steps 50 > gpurun_out/sweep_full.jsonl 2> gpurun_out/sweep.err
python tools/sweep.py --steps 50 --quick > gpurun_out/sweep_variants.jsonl 2>> gpurun_out/sweep.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:laplace_cell_slab3 -s 5 -c 1 -o gpurun_out/prof_slab3 -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg >> gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:laplace_cell_stage -s 3 -c 1 -o gpurun_out/prof_stage -f python tools/sweep.py --steps 5 --custom "3,4,6,f64,40" >> gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none -k regex:laplace_cell_slab3 -s 3 -c 1 -o gpurun_out/prof_slab3_f32 -f python tools/sweep.py --steps 5 --custom "3,4,6,f32,0" >> gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
