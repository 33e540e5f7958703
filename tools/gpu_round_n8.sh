#!/bin/bash
# 8-GPU pass: N-rank parity tests (torchrun) and the scaling lines of BASELINE.json configs[4]
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/n8_pytest.log; cat gpurun_out/n8_pytest.log
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc $?"; tail -c 600 gpurun_out/$name.json | head -c 600; echo; }
run n8_bench_weak_r6 8 --steps 500 --warmup 10
run n8_bench_weak_r7_1Bdofs 8 --refine 7 --steps 50 --warmup 5 --no-cg --no-e2e
run n8_bench_strong_r7 8 --refine 7 --scaling strong --steps 300 --warmup 10 --no-e2e
run n4_bench_weak_r6 4 --steps 500 --warmup 10 --no-cg
