#!/bin/bash
# One GPU-box pass: tests, bench, reference arm, launch list, full ncu capture of the cell kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 30 --warmup 2 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:laplace_cell -s 5 -c 1 -o gpurun_out/prof_default -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-cg >> gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json
